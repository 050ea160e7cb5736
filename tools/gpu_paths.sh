#!/bin/bash
# parity tests + bench for each requested load layer: gpu_paths.sh auto direct async ...
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
for round in 1 2; do
for lp in "$@"; do
timeout 600 python bench.py --no-cpu-baseline --no-e2e --load-path $lp --steps 100 > gpurun_out/bench_$lp.json 2> gpurun_out/bench_$lp.err || tail -3 gpurun_out/bench_$lp.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_$lp.json'))
print('$lp: DXT1 %.0f GB/s | ETC1 %.0f GB/s | dual %.0f GB/s | per-tex DXT1 %.0f' % (d['roofline']['achieved'], d['other_codec']['achieved_gbs_per_gpu'], d['dual_output']['achieved_gbs_per_gpu'], d['per_texture_launch']['achieved_gbs_per_gpu']))
PY
done; done
