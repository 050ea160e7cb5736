#!/bin/bash
# ncu session: launch list of the bench command + one full capture of each encoder kernel.
mkdir -p gpurun_out
timeout 300 tools/pipebench > gpurun_out/pipebench.txt 2>&1; echo "pipebench rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_under_ncu.log 2>&1; echo "launchlist rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:encode_direct -s 9 -c 3 -f -o gpurun_out/prof \
    python tools/profile_target.py > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
ls -la gpurun_out
