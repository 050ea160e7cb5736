#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 Src/goofy_bench --images oracle/_ref/test-data --csv gpurun_out/images.csv > gpurun_out/images.txt 2> gpurun_out/images.err; echo "harness images rc=$?"; tail -1 gpurun_out/images.txt
bash tools/gpu_cfg.sh 1
