#!/bin/bash
# A/B of an environment switch inside one session: usage gpu_env_ab.sh [--load-path P] VAR val1 val2 ...
mkdir -p gpurun_out
LP=auto
if [ "$1" = "--load-path" ]; then LP=$2; shift 2; fi
VAR=$1; shift
for round in 1 2; do
for v in "$@"; do
env $VAR=$v timeout 600 python bench.py --no-cpu-baseline --no-e2e --steps 200 --load-path $LP > gpurun_out/bench_v.json 2> gpurun_out/bench_v.err || tail -3 gpurun_out/bench_v.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_v.json'))
print('$LP $VAR=$v: DXT1 %.0f GB/s | ETC1 %.0f GB/s | dual %.0f GB/s | per-tex DXT1 %.0f | clk %s' % (d['roofline']['achieved'], d['other_codec']['achieved_gbs_per_gpu'], d['dual_output']['achieved_gbs_per_gpu'], d['per_texture_launch']['achieved_gbs_per_gpu'], d['clocks'].get('sm_mhz')))
PY
done; done
