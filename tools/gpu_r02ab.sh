#!/bin/bash
# round 2, session AB: is the end-to-end leg of bench.py slower than the same call from tools/hostlat on the same box?
show() { python - "$1" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); e=d['e2e']
print('  e2e ms', round(e['ms_per_step'],3), 'packing/plain', e['calls_with_packing'], e['calls_plain_dma'], '| rgba-only ms', round(e['rgba_dma_only']['ms_per_step'],3), '| probe ms', round(e['pcie_bound_ms'],3), '| rgb24 host ms', round(e['rgb24_host_call']['ms_per_step'],3), '| dual ms', round(e['dual_output_host_call']['ms_per_step'],3), '| pageable ms', round(e['pageable_buffers']['ms_per_step'],3))
PY
}
for rep in 1 2; do
echo "=== hostlat AUTO / OFF"; tools/hostlat 8192 8192 24 | grep "lib pinned"; GOOFY_B200_HOST_RGB=0 tools/hostlat 8192 8192 24 | grep "lib pinned"
echo "=== bench.py --no-configs"; python bench.py --no-configs --no-cpu-baseline > gpurun_out/ab1.json 2>/dev/null; show gpurun_out/ab1.json
echo "=== bench.py (full)"; python bench.py --no-cpu-baseline > gpurun_out/ab2.json 2>/dev/null; show gpurun_out/ab2.json
done
