#!/bin/bash
# late round-2 checks, ZGEMM tile selection (re-run with the final rule: tests, tile timings, hops, eigensolver, N=32 step).  Output: gpurun_out/late4_*.log
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 ) > gpurun_out/late4_tests.log
timeout 300 python tools/dev_zgemm.py > gpurun_out/late4_zgemm.log 2>&1
timeout 120 python tools/dev_hop.py > gpurun_out/late4_hop.log 2>&1
timeout 120 python tools/dev_heig.py 1024 2048 > gpurun_out/late4_heig.log 2>&1
TEVO_BENCH_BUDGET_S=100 timeout 200 python bench.py --mode sequential --nsites 32 --steps 1 --warmup 1 --no-e2e --no-cpu --no-parallel-mode 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ms_per_step', round(d['ms_per_step'],1), {k: round(v,1) for k,v in d['phases_ms'].items()}, d['centre_moves'])" > gpurun_out/late4_bench.log 2>&1
TEVO_BENCH_BUDGET_S=100 timeout 200 python bench.py --mode parallel --nsites 48 --steps 1 --warmup 1 --no-e2e --no-cpu 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('parallel N=48 ms_per_step', round(d['ms_per_step'],1))" >> gpurun_out/late4_bench.log 2>&1
timeout 200 python tools/dev_tdvp.py 16 1024 > gpurun_out/late4_tdvp.log 2>&1
echo done
