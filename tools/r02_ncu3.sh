#!/bin/bash
# round-2 (late) ncu captures of the kernels changed at the end of the round (one gpurun call, one GPU):
#   DMMA ZGEMM with the 64x32 tile at the theta shape and with the 32x32 tile at the K=128 Hermitian rank-2k shape,
#   the panel version of the CholeskyQR diagonal-block kernel, and a launch list of a chi=1024 step
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 60 python tools/dev_zgemm.py > gpurun_out/plain_z3.log 2>&1 &&
timeout 150 ncu --set full --clock-control none --import-source on -k regex:zgemm_dmma -s 5 -c 1 -o gpurun_out/r02_zgemm_theta_t64x32 python tools/dev_zgemm.py > gpurun_out/ncu_z31.log 2>&1
echo "rc=$?"
timeout 150 ncu --set full --clock-control none --import-source on -k regex:zgemm_dmma -s 74 -c 1 -o gpurun_out/r02_zgemm_k128_t32x32 python tools/dev_zgemm.py > gpurun_out/ncu_z32.log 2>&1
echo "rc=$?"
timeout 100 python tools/dev_hop.py > gpurun_out/plain_hop3.log 2>&1 &&
timeout 200 ncu --set full --clock-control none --import-source on -k regex:potrf_block -s 40 -c 2 -o gpurun_out/r02_potrf_panel_final python tools/dev_hop.py > gpurun_out/ncu_hop3.log 2>&1
echo "rc=$?"
timeout 60 python tools/dev_heig.py 2048 > gpurun_out/plain_h3.log 2>&1 &&
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"leaf_kernel|merge_prepare|rot_apply|gather_kernel|secular|zhat|secvec|dgemm_merge|finalize|scatter|tnorm|unscale" --csv --log-file gpurun_out/r02_late_ncu_stedc_launches.csv python tools/dev_heig.py 2048 > gpurun_out/ncu_h3.log 2>&1
echo "rc=$?"
timeout 100 python bench.py --nsites 24 --chi 1024 --steps 1 --warmup 1 --no-e2e --no-cpu --no-parallel-mode > gpurun_out/plain_b3.log 2>&1 &&
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 30000 -c 2200 --csv --log-file gpurun_out/r02_late_ncu_launches_chi1024_n24.csv python bench.py --nsites 24 --chi 1024 --steps 1 --warmup 1 --no-e2e --no-cpu --no-parallel-mode > gpurun_out/ncu_b3.log 2>&1
echo "rc=$?"
