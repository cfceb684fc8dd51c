#!/bin/bash
# late round-2 checks: three-multiplication complex products in the small-tile ZGEMM kernels.  Output: gpurun_out/late8_*.log
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/late8_tests.log
( timeout 300 python -m pytest tests/test_gpu_parity_at_scale.py -m gpu -x -q -s 2>&1 | grep -E "bond|passed|failed|rel dev" | tail -12 ) > gpurun_out/late8_parity.log
for m in 1 0; do
  echo "== TEVO_ZGEMM_3M=$m" >> gpurun_out/late8_zgemm.log
  TEVO_ZGEMM_3M=$m timeout 100 python tools/dev_zgemm.py >> gpurun_out/late8_zgemm.log 2>&1
done
timeout 120 python tools/dev_heig.py 1024 2048 > gpurun_out/late8_heig.log 2>&1
timeout 120 python tools/dev_hop.py > gpurun_out/late8_hop.log 2>&1
TEVO_BENCH_BUDGET_S=100 timeout 200 python bench.py --mode sequential --nsites 32 --steps 1 --warmup 1 --no-e2e --no-cpu --no-parallel-mode 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ms_per_step', round(d['ms_per_step'],1), {k: round(v,1) for k,v in d['phases_ms'].items()}, d['centre_moves'])" > gpurun_out/late8_bench.log 2>&1
echo done
