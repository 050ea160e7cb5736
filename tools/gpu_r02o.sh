#!/bin/bash
# round 2, session O: rgb24 tests again; hybrid pinned path with non-temporal pack stores and strip sizes; 32-bit vs cooperative loads
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_rgb24.py -x -q -s 2>&1 | tail -15 | tee gpurun_out/o_pytest_rgb24.log
for coop in 1 0; do
echo "=== GOOFY_B200_RGB24_COOP=$coop"
GOOFY_B200_RGB24_COOP=$coop timeout 300 python tools/bench_next_rows.py --steps 30 2>/dev/null | python -c "
import json,sys
d=json.load(sys.stdin)
for k,v in d['results'].items():
    if 'rgb24' in k: print(f\"{k:32s} {v['gb_per_s']:8.0f} GB/s {v['mp_per_s']/1e6:6.3f} TP/s\")"
done
for cfg in "0 1 4096" "1 1 4096" "1 0 4096" "1 1 8192" "1 1 16384" "2 1 4096" "2 1 16384" "2 0 16384"; do
  set -- $cfg
  echo "=== GOOFY_B200_HOST_RGB=$1 GOOFY_B200_PACK_NT=$2 GOOFY_B200_HYBRID_STRIP_KB=$3 8192^2"
  GOOFY_B200_HOST_RGB=$1 GOOFY_B200_PACK_NT=$2 GOOFY_B200_HYBRID_STRIP_KB=$3 tools/hostlat 8192 8192 12 2>&1 | grep -E "lib|same"
done
for sz in "768 512 300" "2048 2048 80"; do
  set -- $sz
  for nt in 0 1; do
    echo "=== GOOFY_B200_PACK_NT=$nt  $1 x $2"
    GOOFY_B200_PACK_NT=$nt tools/hostlat $1 $2 $3 2>&1 | grep -E "lib|same"
  done
done
