#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
L=goofy_b200/libgoofy_b200.so
timeout 900 tools/shapebench --shapes strip1k,strip,tex8192,batch4x8192 --json gpurun_out/shape_d.json \
  r01=build/ab/libgoofy_r01.so \
  new=$L \
  new_noq=$L:GOOFY_B200_WAVE_QUANT=0 \
  dualrows=$L:GOOFY_B200_DUAL_ASYNC=0 \
  oneshot=$L:path=3 \
  rows=$L:path=1 \
  rows_r2=$L:path=1:GOOFY_B200_ROWS_PER_CTA=2 \
  rows_r3=$L:path=1:GOOFY_B200_ROWS_PER_CTA=3 \
  rows_r6=$L:path=1:GOOFY_B200_ROWS_PER_CTA=6 \
  async=$L:path=4 \
  > gpurun_out/shape_d.txt 2>&1; echo "shapebench rc=$?"; cat gpurun_out/shape_d.txt
