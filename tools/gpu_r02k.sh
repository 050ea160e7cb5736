#!/bin/bash
mkdir -p gpurun_out
tools/hostlat 768 512 400 2>&1 | grep -E "lib|bare|same"
tools/hostlat 1024 1024 300 2>&1 | grep -E "lib|bare|same"
tools/hostlat 2048 2048 100 2>&1 | grep -E "lib|bare|same"
for kb in 32768 131072; do
  echo "=== 4096^2 GOOFY_B200_ZEROCOPY_MAX_KB=$kb"
  GOOFY_B200_ZEROCOPY_MAX_KB=$kb tools/hostlat 4096 4096 30 2>&1 | grep -E "lib|bare|same"
done
echo "=== 8192^2 GOOFY_B200_ZEROCOPY_MAX_KB=1048576"
GOOFY_B200_ZEROCOPY_MAX_KB=1048576 tools/hostlat 8192 8192 10 2>&1 | grep -E "lib|bare|same"
tools/hostlat 8192 8192 10 2>&1 | grep -E "lib|bare|same"
