#!/bin/bash
# round 2, session A: parity after the TMA rewrite + computed control table, then the short-launch / load-layer A/B
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv,noheader
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
L=goofy_b200/libgoofy_b200.so
timeout 900 tools/shapebench --iters 200 --rounds 3 --json gpurun_out/shape_a.json \
  r01=build/ab/libgoofy_r01.so \
  new=$L \
  new_nopf=$L:GOOFY_B200_L2PF=0 \
  new128=build/ab/libgoofy_new128.so \
  rows=$L:path=1 \
  async=$L:path=4 \
  tma_s3=$L:path=2 \
  tma_s2=$L:path=2:GOOFY_B200_TMA_STAGES=2 \
  tma_s4=$L:path=2:GOOFY_B200_TMA_STAGES=4 \
  tma_s6=$L:path=2:GOOFY_B200_TMA_STAGES=6 \
  tma_r2s2=$L:path=2:GOOFY_B200_TMA_ROWS=2:GOOFY_B200_TMA_STAGES=2 \
  tma_r2s3=$L:path=2:GOOFY_B200_TMA_ROWS=2:GOOFY_B200_TMA_STAGES=3 \
  tma_s3g2=$L:path=2:GOOFY_B200_TMA_GRID_MULT=2 \
  tma_s3g4=$L:path=2:GOOFY_B200_TMA_GRID_MULT=4 \
  r01_tma=build/ab/libgoofy_r01.so:path=2 \
  > gpurun_out/shape_a.txt 2>&1; echo "shapebench rc=$?"; cat gpurun_out/shape_a.txt
# ncu: the strip shape, AUTO kernel and TMA kernel (full set, 2 launches each after warm-up)
for v in "auto=$L" "tma=$L:path=2"; do
  n=${v%%=*}
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:encode_ -s 6 -c 2 -f -o gpurun_out/r02a_strip_$n \
     tools/shapebench --iters 4 --rounds 1 --shapes strip --modes dxt1 "$v" > gpurun_out/ncu_strip_$n.log 2>&1; echo "ncu $n rc=$?"
done
