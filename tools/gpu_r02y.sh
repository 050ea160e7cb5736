#!/bin/bash
# round 2, session Y: write-combined staging strips (cudaHostAllocWriteCombined) for the input side of the host path
for rep in 1 2; do for wc in 0 1; do
  echo "=== GOOFY_B200_STAGE_WC=$wc (rep $rep)"
  GOOFY_B200_STAGE_WC=$wc tools/hostlat 8192 8192 16 2>&1 | grep -E "lib p"
  GOOFY_B200_STAGE_WC=$wc tools/hostlat 2048 2048 80 2>&1 | grep -E "lib pageable"
  GOOFY_B200_STAGE_WC=$wc tools/hostlat 768 512 300 2>&1 | grep -E "lib pageable|same"
done; done
echo "=== WC + regular stores"; GOOFY_B200_STAGE_WC=1 GOOFY_B200_HOST_RGB=2 tools/hostlat 8192 8192 12 2>&1 | grep -E "lib pinned"
