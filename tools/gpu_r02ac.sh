#!/bin/bash
# round 2, session AC: L2 warm-up before griddepcontrol.wait in the relaxed-shape kernels
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "relaxed" 2>&1 | tail -2
for pf in 1 0 1 0; do
echo "=== GOOFY_B200_L2PF=$pf"
GOOFY_B200_L2PF=$pf timeout 300 python tools/bench_next_rows.py --steps 50 2>/dev/null | python -c "
import json,sys
d=json.load(sys.stdin)['results']
print(' '.join(f\"{k.replace('encode_','')}={v['gb_per_s']:.0f}\" for k,v in d.items() if 'relaxed' in k or k in ('encode_dxt1','encode_etc1')))"
done
