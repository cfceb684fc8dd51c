#!/bin/bash
# round-2 ncu full captures (one gpurun call, one GPU, ~4 minutes):
#   gpurun --timeout 700 -- 'bash tools/r02_ncu.sh > gpurun_out/r02_ncu.log 2>&1'
# (the launch list profiles/r02_ncu_launches_* came from:  ncu --metrics gpu__time_duration.sum --clock-control none
#  -s 30000 -c 3000 --csv python bench.py --nsites 24 --chi 1024 --steps 1 --warmup 1 --no-e2e --no-cpu --no-parallel-mode)
cd "$(dirname "$0")/.."
echo "== full capture of the dominant kernel (tridiag_panel_kernel, n=2048, first two panels of a factorisation)"
timeout 120 python tools/dev_heig.py 2048 > gpurun_out/plain_h.log 2>&1 &&
timeout 250 ncu --set full --clock-control none --import-source on -k regex:tridiag_panel -s 52 -c 2 -o gpurun_out/r02_panel_prof python tools/dev_heig.py 2048 > gpurun_out/ncu_h.log 2>&1
echo "rc=$?"
echo "== full capture of the DMMA ZGEMM at the theta shape (2048x2048x1024) and at the K=128 rank-2k shape"
timeout 60 python tools/dev_zgemm.py > gpurun_out/plain_z.log 2>&1 &&
timeout 150 ncu --set full --clock-control none --import-source on -k regex:zgemm_dmma -s 5 -c 1 -o gpurun_out/r02_zgemm_theta python tools/dev_zgemm.py > gpurun_out/ncu_z1.log 2>&1
echo "rc=$?"
timeout 150 ncu --set full --clock-control none --import-source on -k regex:zgemm_dmma -s 74 -c 1 -o gpurun_out/r02_zgemm_k128 python tools/dev_zgemm.py > gpurun_out/ncu_z2.log 2>&1
echo "rc=$?"
