#!/bin/bash
# round 2, session R: hybrid pinned path: calling thread packs too (sync) vs only feeds the link (async); strip size
mkdir -p gpurun_out
for rep in 1 2; do
for as in 1 0; do
for kb in 0 4096 16384; do
  echo "=== GOOFY_B200_HYBRID_ASYNC=$as GOOFY_B200_HYBRID_STRIP_KB=$kb (rep $rep)"
  GOOFY_B200_HYBRID_ASYNC=$as GOOFY_B200_HYBRID_STRIP_KB=$kb tools/hostlat 8192 8192 12 2>&1 | grep -E "lib pinned"
  GOOFY_B200_HYBRID_ASYNC=$as GOOFY_B200_HYBRID_STRIP_KB=$kb tools/hostlat 4096 4096 30 2>&1 | grep -E "lib pinned"
done; done
echo "=== GOOFY_B200_HOST_RGB=0 (rep $rep)"
GOOFY_B200_HOST_RGB=0 tools/hostlat 8192 8192 12 2>&1 | grep -E "lib pinned"
GOOFY_B200_HOST_RGB=0 tools/hostlat 4096 4096 30 2>&1 | grep -E "lib pinned"
done
echo "=== GOOFY_B200_HOST_THREADS=9 async"; GOOFY_B200_HOST_THREADS=9 tools/hostlat 8192 8192 12 2>&1 | grep -E "lib pinned"
echo "=== GOOFY_B200_HOST_THREADS=12 async"; GOOFY_B200_HOST_THREADS=12 tools/hostlat 8192 8192 12 2>&1 | grep -E "lib pinned"
echo "=== GOOFY_B200_HOST_RGB=2 async"; GOOFY_B200_HOST_RGB=2 tools/hostlat 8192 8192 12 2>&1 | grep -E "lib pinned"
timeout 300 python -m pytest tests/test_gpu_rgb24.py -x -q 2>&1 | tail -3
