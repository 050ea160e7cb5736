#!/bin/bash
# round 2, session Q: full GPU test suite, bench line, hybrid pinned path after the raw-tail rule
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/q_pytest_gpu.log
for sz in "4096 4096 30" "8192 8192 12" "16384 8192 6"; do
  set -- $sz
  for m in 0 1; do
    echo "=== GOOFY_B200_HOST_RGB=$m  $1 x $2"
    GOOFY_B200_HOST_RGB=$m tools/hostlat $1 $2 $3 2>&1 | grep -E "lib|same"
  done
done
timeout 600 python bench.py > gpurun_out/q_bench_n1.json 2> gpurun_out/q_bench_n1.err; tail -3 gpurun_out/q_bench_n1.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/q_bench_n1.json').read().strip().splitlines()[-1])
print('value', d['value'], 'frac', d['roofline']['frac'])
print('rgb24', d['rgb24_input'])
e=d['e2e']; print({k:v for k,v in e.items() if k not in ('pcie_bound_note','api')})
print('cpu', d.get('cpu_baseline',{}).get('value'), d.get('cpu_baseline',{}).get('cores'))
PY
