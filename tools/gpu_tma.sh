#!/bin/bash
mkdir -p gpurun_out
for st in 2 3 4 6; do
GOOFY_B200_TMA_STAGES=$st timeout 600 python bench.py --no-cpu-baseline --no-e2e --load-path tma --steps 30 > gpurun_out/bench_tma$st.json 2> gpurun_out/bench_tma$st.err; echo "bench tma stages=$st rc=$?"
python - <<PY
import json
d=json.load(open('gpurun_out/bench_tma$st.json'))
print('stages $st: DXT1 %.0f MP/s %.0f GB/s frac %.3f | ETC1 %.0f MP/s %.0f GB/s | dual %.0f MP/s %.0f GB/s' % (d['value'], d['roofline']['achieved'], d['roofline']['frac'], d['other_codec']['value'], d['other_codec']['achieved_gbs_per_gpu'], d['dual_output']['value'], d['dual_output']['achieved_gbs_per_gpu']))
PY
done
GOOFY_B200_TMA_STAGES=2 timeout 600 python -m pytest tests -m gpu -x -q -k tma 2>&1 | tail -2
