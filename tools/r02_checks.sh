#!/bin/bash
# round-2 A/B checks on one B200 (about 6 minutes):  gpurun --timeout 900 -- 'bash tools/r02_checks.sh > gpurun_out/r02_checks.log 2>&1'
cd "$(dirname "$0")/.."
echo "== full GPU suite with the cluster tail as default"
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
echo "== DMMA ZGEMM device-resident throughput: run-time signs (default) vs compile-time signs"
timeout 60 python tools/dev_zgemm.py
TEVO_ZGEMM_CTSIGN=1 timeout 60 python tools/dev_zgemm.py
echo "== operand staging by the TMA engine (cp.async.bulk + mbarrier): tests, then throughput"
TEVO_ZGEMM_BULK=1 timeout 200 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_tebd.py -m gpu -x -q 2>&1 | tail -3
TEVO_ZGEMM_BULK=1 timeout 60 python tools/dev_zgemm.py
echo "== layer-parallel mode, lanes 3 (default) / 4 / 6: N=48 chi=1024, one TEBD4 step each"
for l in 3 4 6; do
  TEVO_LANES=$l TEVO_BENCH_BUDGET_S=100 timeout 200 python bench.py --mode parallel --nsites 48 --steps 1 --warmup 1 --no-e2e --no-cpu 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('lanes', $l, 'ms_per_step', round(d['ms_per_step'],1))"
done
