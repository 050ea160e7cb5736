#!/bin/bash
mkdir -p gpurun_out
python tools/strip_launch_probe.py 2>&1 | tee gpurun_out/strip_launch_probe.txt
tools/shapebench --shapes strip --modes dxt1 new=goofy_b200/libgoofy_b200.so 2>&1 | tail -3
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_n1.err
python -c "
import json; d=json.load(open('gpurun_out/bench_n1.json')); print(json.dumps(d['clocks'])); print(d['configs']['strip16384']['ms_per_step_by_rank'], d['value'])"
