#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
for kb in 0 4096; do
  echo "=== GOOFY_B200_ZEROCOPY_MAX_KB=$kb"
  GOOFY_B200_ZEROCOPY_MAX_KB=$kb tools/hostlat 768 512 400 2>&1 | grep -E "lib|bare|same"
  GOOFY_B200_ZEROCOPY_MAX_KB=$kb tools/hostlat 1024 1024 300 2>&1 | grep -E "lib|bare|same"
done
for kb in 0 32768; do
  echo "=== GOOFY_B200_ZEROCOPY_MAX_KB=$kb"
  GOOFY_B200_ZEROCOPY_MAX_KB=$kb tools/hostlat 2048 2048 100 2>&1 | grep -E "lib|bare|same"
done
GOOFY_B200_TRACE_HOST=500 tools/hostlat 768 512 300 2>&1 | grep trace
GOOFY_B200_TRACE_HOST=300 tools/hostlat 768 512 300 2>&1 | grep trace
