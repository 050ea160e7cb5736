#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
for lp in "$@"; do
timeout 600 python bench.py --no-cpu-baseline --no-e2e --load-path $lp --steps 30 > gpurun_out/bench_$lp.json 2> gpurun_out/bench_$lp.err || tail -3 gpurun_out/bench_$lp.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_$lp.json'))
print('$lp: DXT1 %.0f MP/s %.0f GB/s frac %.3f | ETC1 %.0f MP/s %.0f GB/s | dual %.0f MP/s %.0f GB/s' % (d['value'], d['roofline']['achieved'], d['roofline']['frac'], d['other_codec']['value'], d['other_codec']['achieved_gbs_per_gpu'], d['dual_output']['value'], d['dual_output']['achieved_gbs_per_gpu']))
PY
done
