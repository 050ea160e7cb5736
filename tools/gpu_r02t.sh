#!/bin/bash
# round 2, session T: measured choice between packing and plain DMA (HybridChoice): N=1 host calls, then the bench at N ranks
N=${1:-1}
mkdir -p gpurun_out
if [ "$N" = 1 ]; then
  timeout 300 python -m pytest tests/test_gpu_rgb24.py -x -q 2>&1 | tail -3
  for m in 1 0 1 0; do echo "=== GOOFY_B200_HOST_RGB=$m"; GOOFY_B200_HOST_RGB=$m tools/hostlat 8192 8192 24 2>&1 | grep -E "lib pinned"; GOOFY_B200_HOST_RGB=$m tools/hostlat 4096 4096 40 2>&1 | grep -E "lib pinned"; done
fi
bash tools/gpu_r02s.sh $N
