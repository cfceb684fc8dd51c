#!/bin/bash
# round-2 ncu captures of the remaining kernel families (one gpurun call, one GPU, ~3 minutes):
#   divide & conquer + CholeskyQR block kernels (centre moves at chi = 1024) and the TDVP kernels (Lanczos BLAS-1,
#   mid_contract) at chi = 512
cd "$(dirname "$0")/.."
timeout 100 python tools/dev_hop.py > gpurun_out/plain_hop.log 2>&1 &&
timeout 250 ncu --set full --clock-control none --import-source on -k regex:"potrf_block|secular_kernel|merge_prepare|leaf_kernel|dgemm_merge" -s 40 -c 8 -o gpurun_out/r02_hop_kernels python tools/dev_hop.py > gpurun_out/ncu_hop.log 2>&1
echo "rc=$?"
timeout 100 python tools/dev_tdvp.py 12 512 > gpurun_out/plain_tdvp.log 2>&1 &&
timeout 250 ncu --set full --clock-control none --import-source on -k regex:"multi_dotc|multi_axpy|lincomb|mid_contract|gate_apply|expect2" -s 200 -c 6 -o gpurun_out/r02_tdvp_kernels python tools/dev_tdvp.py 12 512 > gpurun_out/ncu_tdvp.log 2>&1
echo "rc=$?"
