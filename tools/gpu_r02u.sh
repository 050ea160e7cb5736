#!/bin/bash
# round 2, session U: why the hybrid host call is slower inside bench.py's process than from tools/hostlat
mkdir -p gpurun_out
echo "=== hostlat (C++ caller)"; tools/hostlat 8192 8192 24 2>&1 | grep -E "lib pinned"
echo "=== python caller, torch pinned buffers (tools/gpu_hostpath.py)"; python tools/gpu_hostpath.py 8192 8192 2>&1 | grep pinned
echo "=== same, OMP_NUM_THREADS=1"; OMP_NUM_THREADS=1 python tools/gpu_hostpath.py 8192 8192 2>&1 | grep pinned
echo "=== bench.py"; bash tools/gpu_r02s.sh 1 | cut -c1-700
echo "=== bench.py OMP_NUM_THREADS=1"; OMP_NUM_THREADS=1 bash tools/gpu_r02s.sh 1 | cut -c1-700
echo "=== hostlat again"; tools/hostlat 8192 8192 24 2>&1 | grep -E "lib pinned"
