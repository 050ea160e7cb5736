#!/bin/bash
# quick iteration: parity tests + bench (both codecs) + optional tools
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -2 gpurun_out/bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench.json'))
print('DXT1 %.0f MP/s %.0f GB/s frac %.3f | per-tex %.0f GB/s | ETC1 %.0f MP/s %.0f GB/s | dual %.0f MP/s %.0f GB/s | e2e %.0f | clocks %s' % (d['value'], d['roofline']['achieved'], d['roofline']['frac'], d['per_texture_launch']['achieved_gbs_per_gpu'], d['other_codec']['value'], d['other_codec']['achieved_gbs_per_gpu'], d['dual_output']['value'], d['dual_output']['achieved_gbs_per_gpu'], d.get('e2e',{}).get('value',0), d['clocks']))
PY
for t in "$@"; do timeout 300 $t > gpurun_out/$(basename ${t%% *}).txt 2>&1; echo "$t rc=$?"; tail -3 gpurun_out/$(basename ${t%% *}).txt; done
