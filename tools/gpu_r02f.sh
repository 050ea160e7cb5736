#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu.log
L=goofy_b200/libgoofy_b200.so
timeout 900 tools/shapebench --shapes strip1k,strip,tex8192,batch1024p,batch4x8192 --json gpurun_out/shape_f.json \
  r01=build/ab/libgoofy_r01.so \
  new=$L \
  new_nopfn=$L:GOOFY_B200_PF_NEXT=0 \
  > gpurun_out/shape_f.txt 2>&1; echo "shapebench rc=$?"; cat gpurun_out/shape_f.txt
tools/hostlat 768 512 400 > gpurun_out/hostlat_768.txt 2>&1; cat gpurun_out/hostlat_768.txt
tools/hostlat 2048 2048 100 > gpurun_out/hostlat_2048.txt 2>&1; cat gpurun_out/hostlat_2048.txt
