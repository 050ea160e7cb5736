#!/bin/bash
mkdir -p gpurun_out
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"encode_|decode_|block_sse" -f -o gpurun_out/r02_all python tools/profile_all_target.py > gpurun_out/ncu_all.log 2>&1; echo "ncu all rc=$?"; grep -v "==PROF==" gpurun_out/ncu_all.log | tail -3
# the headline launches: launch list of the bench command + full capture of the three AUTO kernels (round-1 recipe)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-configs > gpurun_out/bench_under_ncu.log 2>&1; echo "launchlist rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:encode_ -s 9 -c 3 -f -o gpurun_out/prof \
    python tools/profile_target.py > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
for lib in goofy_b200/libgoofy_b200.so build/ab/libgoofy_relaxed6.so; do
GOOFY_B200_LIB=$PWD/$lib timeout 300 python tools/bench_next_rows.py 2>/dev/null | python -c "
import json,sys
d=json.load(sys.stdin)['results']
print('$lib', {k: round(v['gb_per_s']) for k,v in d.items() if 'relaxed' in k or 'floatref' in k})"
done
