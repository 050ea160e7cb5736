#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
L=goofy_b200/libgoofy_b200.so
timeout 900 tools/shapebench --shapes strip,tex8192,batch1024p,batch4x8192 --json gpurun_out/shape_c.json \
  r01=build/ab/libgoofy_r01.so \
  new=$L \
  new_nopf=$L:GOOFY_B200_L2PF=0 \
  oneshot=$L:path=3 \
  rows=$L:path=1 \
  tma_rb4s2=$L:path=2:GOOFY_B200_TMA_STAGES=2 \
  tma_rb4s3g4=$L:path=2:GOOFY_B200_TMA_GRID_MULT=4 \
  > gpurun_out/shape_c.txt 2>&1; echo "shapebench rc=$?"; cat gpurun_out/shape_c.txt
