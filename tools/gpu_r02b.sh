#!/bin/bash
# round 2, session B: producer/consumer TMA ring (tile heights, ring depths) against the plain-load layers
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
L=goofy_b200/libgoofy_b200.so
timeout 900 tools/shapebench --json gpurun_out/shape_b.json \
  r01=build/ab/libgoofy_r01.so \
  new=$L \
  new_nopf=$L:GOOFY_B200_L2PF=0 \
  r01_tma=build/ab/libgoofy_r01.so:path=2 \
  tma_rb4s3=$L:path=2 \
  tma_rb4s2=$L:path=2:GOOFY_B200_TMA_STAGES=2 \
  tma_rb4s4=$L:path=2:GOOFY_B200_TMA_STAGES=4 \
  tma_rb2s3=$L:path=2:GOOFY_B200_TMA_ROWS=2 \
  tma_rb2s4=$L:path=2:GOOFY_B200_TMA_ROWS=2:GOOFY_B200_TMA_STAGES=4 \
  tma_rb2s6=$L:path=2:GOOFY_B200_TMA_ROWS=2:GOOFY_B200_TMA_STAGES=6 \
  tma_rb1s4=$L:path=2:GOOFY_B200_TMA_ROWS=1:GOOFY_B200_TMA_STAGES=4 \
  tma_rb4s3g2=$L:path=2:GOOFY_B200_TMA_GRID_MULT=2 \
  tma_rb4s3g4=$L:path=2:GOOFY_B200_TMA_GRID_MULT=4 \
  > gpurun_out/shape_b.txt 2>&1; echo "shapebench rc=$?"; cat gpurun_out/shape_b.txt
for v in "tma=$L:path=2"; do
  n=${v%%=*}
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:encode_ -s 6 -c 2 -f -o gpurun_out/r02b_strip_$n \
     tools/shapebench --iters 4 --rounds 1 --shapes strip --modes dxt1 "$v" > gpurun_out/ncu_strip_$n.log 2>&1; echo "ncu $n rc=$?"
done
