#!/bin/bash
# late round-2 checks: divide & conquer (leaf rotations, parallel deflation search / compaction, sliced ranking) and
# parallel sums in the truncation kernels.  Output: gpurun_out/late7_*.log
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/late7_tests.log
timeout 120 python tools/dev_heig.py 512 1024 2048 > gpurun_out/late7_heig.log 2>&1
TEVO_BENCH_BUDGET_S=100 timeout 200 python bench.py --mode sequential --nsites 32 --steps 1 --warmup 1 --no-e2e --no-cpu --no-parallel-mode 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ms_per_step', round(d['ms_per_step'],1), {k: round(v,1) for k,v in d['phases_ms'].items()}, d['centre_moves'])" > gpurun_out/late7_bench.log 2>&1
echo done
