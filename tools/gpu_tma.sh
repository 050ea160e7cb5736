#!/bin/bash
# TMA tile kernel: sweep one environment knob.  usage: gpu_tma.sh VAR v1 v2 ...
mkdir -p gpurun_out
VAR=$1; shift
for v in "$@"; do
env $VAR=$v timeout 600 python bench.py --no-cpu-baseline --no-e2e --load-path tma --steps 100 > gpurun_out/bench_tma.json 2> gpurun_out/bench_tma.err || tail -3 gpurun_out/bench_tma.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_tma.json'))
print('tma $VAR=$v: DXT1 %.0f GB/s | ETC1 %.0f GB/s | dual %.0f GB/s | per-tex DXT1 %.0f' % (d['roofline']['achieved'], d['other_codec']['achieved_gbs_per_gpu'], d['dual_output']['achieved_gbs_per_gpu'], d['per_texture_launch']['achieved_gbs_per_gpu']))
PY
done
