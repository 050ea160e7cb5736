#!/bin/bash
mkdir -p gpurun_out
for lib in "$@"; do
GOOFY_B200_LIB=$PWD/$lib timeout 600 python bench.py --no-cpu-baseline --no-e2e --load-path tma --steps 100 > gpurun_out/bench_v.json 2> gpurun_out/bench_v.err || tail -3 gpurun_out/bench_v.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_v.json'))
print('$lib: DXT1 %.0f GB/s | ETC1 %.0f GB/s | dual %.0f GB/s | per-tex DXT1 %.0f' % (d['roofline']['achieved'], d['other_codec']['achieved_gbs_per_gpu'], d['dual_output']['achieved_gbs_per_gpu'], d['per_texture_launch']['achieved_gbs_per_gpu']))
PY
done
GOOFY_B200_LIB=$PWD/goofy_b200/libvariant_tmab2.so timeout 600 python -m pytest tests -m gpu -x -q -k tma 2>&1 | tail -2
