#!/bin/bash
# late round-2 checks: pipelined Cholesky chain of the centre moves (TEVO_CHOL_PIPELINE) and the compact-WY fork
# (TEVO_WY_OVERLAP), tests first, then A/B timings.  Output: gpurun_out/late_*.log
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 ) > gpurun_out/late_tests.log
for p in 0 1; do
  echo "== TEVO_CHOL_PIPELINE=$p" >> gpurun_out/late_hop.log
  TEVO_CHOL_PIPELINE=$p timeout 120 python tools/dev_hop.py >> gpurun_out/late_hop.log 2>&1
done
for w in 0 1; do
  echo "== TEVO_WY_OVERLAP=$w" >> gpurun_out/late_bench.log
  TEVO_WY_OVERLAP=$w TEVO_BENCH_BUDGET_S=100 timeout 200 python bench.py --mode sequential --nsites 32 --steps 1 --warmup 1 --no-e2e --no-cpu --no-parallel-mode 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ms_per_step', round(d['ms_per_step'],1), {k: round(v,1) for k,v in d['phases_ms'].items()}, d['centre_moves'])" >> gpurun_out/late_bench.log 2>&1
done
echo done
