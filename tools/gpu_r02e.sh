#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
L=goofy_b200/libgoofy_b200.so
timeout 900 tools/shapebench --shapes strip1k,strip,tex8192,batch4x8192 --json gpurun_out/shape_e.json \
  r01=build/ab/libgoofy_r01.so \
  new=$L \
  new_pfn=$L:GOOFY_B200_PF_NEXT=1 \
  rows_pfn=$L:path=1:GOOFY_B200_PF_NEXT=1 \
  rows_r2_pfn=$L:path=1:GOOFY_B200_PF_NEXT=1:GOOFY_B200_ROWS_PER_CTA=2 \
  rows_r3_pfn=$L:path=1:GOOFY_B200_PF_NEXT=1:GOOFY_B200_ROWS_PER_CTA=3 \
  rows_r6_pfn=$L:path=1:GOOFY_B200_PF_NEXT=1:GOOFY_B200_ROWS_PER_CTA=6 \
  > gpurun_out/shape_e.txt 2>&1; echo "shapebench rc=$?"; cat gpurun_out/shape_e.txt
