#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for t in 0 1 2; do
  echo "== TEVO_ZGEMM_TILE=$t" >> gpurun_out/late3_zgemm.log
  TEVO_ZGEMM_TILE=$t timeout 200 python tools/dev_zgemm.py >> gpurun_out/late3_zgemm.log 2>&1
done
echo done
