#!/bin/bash
# round 2, session AH: roofline.traffic re-measured by bench.py itself (ncu child after the timed legs); and not under ncu
mkdir -p gpurun_out
SECONDS=0
python bench.py --no-configs > gpurun_out/ah_bench.json 2> gpurun_out/ah_bench.err; echo "bench.py --no-configs: $SECONDS s"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/ah_bench.json').read().strip().splitlines()[-1])
print('traffic', d['roofline']['traffic'], '|', d['roofline']['traffic_source'][:200]); print('value', round(d['value']), 'e2e', round(d['e2e']['value']))
PY
SECONDS=0
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/ah_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-configs > gpurun_out/ah_bench_under_ncu.json 2> gpurun_out/ah_under_ncu.err; echo "under ncu rc=$? $SECONDS s"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/ah_bench_under_ncu.json').read().strip().splitlines()[-1])
print('under ncu: traffic', d['roofline']['traffic'], '|', d['roofline']['traffic_source'][-120:])
PY
env | grep -i -E "inject|nsight|profiler" | head -3
