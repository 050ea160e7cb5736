#!/bin/bash
# round 2, session AA: programmatic dependent launch in the secondary kernels (float-reference, relaxed, ragged batch, decode, SSE)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
for pdl in 1 0 1 0; do
echo "=== GOOFY_B200_PDL=$pdl"
GOOFY_B200_PDL=$pdl timeout 300 python tools/bench_next_rows.py --steps 50 2>/dev/null | tee gpurun_out/aa_next_rows_pdl$pdl.json | python -c "
import json,sys
d=json.load(sys.stdin)['results']
print(' '.join(f\"{k.replace('encode_','')}={v['gb_per_s']:.0f}\" for k,v in d.items() if 'rgb24' not in k))"
done
