#!/bin/bash
# round 2, session X: relaxed shapes through the strict kernels + edge kernel: parity, throughput (split vs all-in-one)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "relaxed" 2>&1 | tail -4
for split in 1 0; do
echo "=== GOOFY_B200_RELAXED_SPLIT=$split"
GOOFY_B200_RELAXED_SPLIT=$split timeout 300 python tools/bench_next_rows.py --steps 50 2>/dev/null | python -c "
import json,sys
d=json.load(sys.stdin)['results']
for k,v in d.items():
    if 'relaxed' in k or k in ('encode_dxt1','encode_etc1','encode_dxt1_floatref','encode_etc1_floatref'): print(f\"{k:28s} {v['gb_per_s']:7.0f} GB/s  {v['frac_of_measured_peak']:.3f} of measured peak\")"
done
Src/goofy_bench --size 8192 --textures 4 --iters 20 --rgb24
