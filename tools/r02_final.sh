#!/bin/bash
# end of round 2: GPU suite on the final library, the driver's bench command, then the C3 and C4 lines
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 ) > gpurun_out/final2_tests.log
timeout 100 python tools/dev_zgemm.py > gpurun_out/final2_zgemm.log 2>&1
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/final2_bench.json 2> gpurun_out/final2_bench.err
timeout 200 python bench.py --nsites 100 --chi 256 --steps 3 --warmup 3 --budget 100 > gpurun_out/final2_c3.json 2> gpurun_out/final2_c3.err
timeout 200 python bench.py --workload tdvp --nsites 64 --chi 1024 --steps 1 --warmup 1 --no-cpu > gpurun_out/final2_c4.json 2> gpurun_out/final2_c4.err
echo done
