#!/bin/bash
# First GPU call of the next round (about 2 minutes on one B200):
#   gpurun --timeout 400 -- 'bash tools/next_round_checks.sh > gpurun_out/next_round_checks.log 2>&1'
# 1. validates the experimental cluster-resident tail of the tridiagonalisation (csrc/tail.cu, never run on hardware)
#    with the existing eigensolver tests, single CTA first, then clusters of 8 and 16;
# 2. times the eigensolver with and without it;
# 3. A/B-times the two kernels that were switched at the very end of round 1 without an isolated measurement
#    (register-tiled Cholesky block, 128-wide back-transformation).
cd "$(dirname "$0")/.."
for c in 1 8 16; do
  echo "== TEVO_CLUSTER_TAIL=$c: eigensolver tests"
  TEVO_CLUSTER_TAIL=$c timeout 120 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "heig or split" 2>&1 | tail -3
done
for c in 0 1 8 16; do
  echo "== TEVO_CLUSTER_TAIL=$c: dev_heig 1024 2048"
  TEVO_CLUSTER_TAIL=$c timeout 60 python tools/dev_heig.py 1024 2048 2>&1 | tail -4
done
echo "== centre moves: register-tiled Cholesky block (default) vs shared-memory version"
timeout 60 python tools/dev_hop.py 2>&1 | tail -1
TEVO_POTRF_SMEM=1 timeout 60 python tools/dev_hop.py 2>&1 | tail -1
echo "== back-transformation: pairs of panels (default) vs single panels"
timeout 60 python tools/dev_heig.py 2048 2>&1 | tail -2
TEVO_BACK64=1 timeout 60 python tools/dev_heig.py 2048 2>&1 | tail -2
echo "== ZGEMM with compile-time conjugation signs (experimental) vs default: tests, then timing"
TEVO_ZGEMM_CTSIGN=1 timeout 120 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "zgemm or split or heig" 2>&1 | tail -3
timeout 120 python tools/dev_time.py 1024 24 2>&1 | tail -1      # per-phase device times of TEBD4 layers at chi=1024
TEVO_ZGEMM_CTSIGN=1 timeout 120 python tools/dev_time.py 1024 24 2>&1 | tail -1
echo "== cluster barrier micro-benchmark"
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/cluster_bar_bench proto/cluster_bar_bench.cu && timeout 20 /tmp/cluster_bar_bench
