#!/bin/bash
# round 2, session N: packed-RGB kernels and the alpha-stripping host path (tests, device throughput, host-call A/B)
mkdir -p gpurun_out
nproc > gpurun_out/n_nproc.txt
timeout 900 python -m pytest tests/test_gpu_rgb24.py -x -q -s 2>&1 | tail -15 | tee gpurun_out/n_pytest_rgb24.log
timeout 300 python tools/bench_next_rows.py --steps 30 > gpurun_out/n_next_rows.json 2> gpurun_out/n_next_rows.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/n_next_rows.json'))
for k,v in d['results'].items():
    if 'rgb24' in k or k in ('encode_dxt1','encode_etc1'): print(f"{k:32s} {v['gb_per_s']:8.0f} GB/s {v['mp_per_s']/1e6:6.3f} TP/s")
PY
for cfg in "0 8" "1 8" "2 8" "1 16" "2 16" "1 4"; do
  set -- $cfg
  echo "=== GOOFY_B200_HOST_RGB=$1 GOOFY_B200_HOST_THREADS=$2  8192^2"
  GOOFY_B200_HOST_RGB=$1 GOOFY_B200_HOST_THREADS=$2 tools/hostlat 8192 8192 12 2>&1 | grep -E "lib|same"
done
for sz in "768 512 300" "2048 2048 80" "4096 4096 30"; do
  set -- $sz
  for m in 0 1; do
    echo "=== GOOFY_B200_HOST_RGB=$m  $1 x $2"
    GOOFY_B200_HOST_RGB=$m tools/hostlat $1 $2 $3 2>&1 | grep -E "lib|same"
  done
done
