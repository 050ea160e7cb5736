#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
timeout 600 python tools/bench_next_rows.py > gpurun_out/next_rows.json 2> gpurun_out/next_rows.err; echo "next rows rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/next_rows.json'))
for k,v in d['results'].items(): print(f"{k:28s} {v['gb_per_s']:7.0f} GB/s  frac {v['frac_of_measured_peak']:.3f}")
PY
timeout 600 Src/goofy_bench --images oracle/_ref/test-data --csv gpurun_out/images.csv > gpurun_out/images.txt 2> gpurun_out/images.err; echo "harness images rc=$?"; tail -1 gpurun_out/images.txt; head -3 gpurun_out/images.err
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"encode_|decode_|block_sse" -f -o gpurun_out/r02_all python tools/profile_all_target.py > gpurun_out/ncu_all.log 2>&1; echo "ncu all rc=$?"; tail -2 gpurun_out/ncu_all.log
