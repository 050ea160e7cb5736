#!/bin/bash
mkdir -p gpurun_out
for m in 1 2 3 4 7 14; do
export GOOFY_B200_ROWS_GY_MULT=$m
timeout 600 python bench.py --no-cpu-baseline --no-e2e --load-path direct --steps 30 > gpurun_out/bench_pf.json 2> gpurun_out/bench_pf.err || tail -3 gpurun_out/bench_pf.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_pf.json'))
print('rows gy x$m: DXT1 %.0f MP/s %.0f GB/s frac %.3f | ETC1 %.0f MP/s %.0f GB/s | dual %.0f MP/s %.0f GB/s' % (d['value'], d['roofline']['achieved'], d['roofline']['frac'], d['other_codec']['value'], d['other_codec']['achieved_gbs_per_gpu'], d['dual_output']['value'], d['dual_output']['achieved_gbs_per_gpu']))
PY
done
