#!/bin/bash
mkdir -p gpurun_out
for lib in goofy_b200/libvariant_*.so goofy_b200/libgoofy_b200.so; do
GOOFY_B200_LIB=$PWD/$lib timeout 600 python bench.py --no-cpu-baseline --no-e2e --steps 30 > gpurun_out/bench_v.json 2> gpurun_out/bench_v.err || tail -3 gpurun_out/bench_v.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_v.json'))
print('$lib: DXT1 %.0f MP/s %.0f GB/s frac %.3f | ETC1 %.0f MP/s %.0f GB/s | dual %.0f MP/s %.0f GB/s' % (d['value'], d['roofline']['achieved'], d['roofline']['frac'], d['other_codec']['value'], d['other_codec']['achieved_gbs_per_gpu'], d['dual_output']['value'], d['dual_output']['achieved_gbs_per_gpu']))
PY
done
