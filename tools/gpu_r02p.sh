#!/bin/bash
# round 2, session P: pack prefetch / threads in the hybrid pinned path; NT stores and load flavour in the pageable path
mkdir -p gpurun_out
for cfg in "8 0" "8 512" "8 1024" "12 0" "12 1024" "16 0" "16 1024" "6 1024"; do
  set -- $cfg
  echo "=== hybrid: GOOFY_B200_HOST_THREADS=$1 GOOFY_B200_PACK_PREFETCH=$2 8192^2"
  GOOFY_B200_HOST_THREADS=$1 GOOFY_B200_PACK_PREFETCH=$2 tools/hostlat 8192 8192 12 2>&1 | grep -E "lib"
done
for rep in 1 2 3; do for nt in 0 1; do
  echo "=== pageable: GOOFY_B200_PACK_NT=$nt (rep $rep) 8192^2 / 4096^2"
  GOOFY_B200_PACK_NT=$nt tools/hostlat 8192 8192 8 2>&1 | grep -E "lib pageable"
  GOOFY_B200_PACK_NT=$nt tools/hostlat 4096 4096 20 2>&1 | grep -E "lib pageable"
done; done
for sz in "768 512 300" "2048 2048 80" "4096 4096 20"; do
  set -- $sz
  for coop in 0 1; do
    echo "=== pageable: GOOFY_B200_RGB24_COOP=$coop GOOFY_B200_PACK_NT=0 $1 x $2"
    GOOFY_B200_RGB24_COOP=$coop GOOFY_B200_PACK_NT=0 tools/hostlat $1 $2 $3 2>&1 | grep -E "lib pageable|same"
  done
done
echo "=== 4096^2 pinned hybrid (strip = 1/16 of the image)"
tools/hostlat 4096 4096 30 2>&1 | grep -E "lib pinned"
GOOFY_B200_HOST_RGB=0 tools/hostlat 4096 4096 30 2>&1 | grep -E "lib pinned"
