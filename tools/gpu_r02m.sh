#!/bin/bash
for rep in 1 2; do
for z in 0 1; do
  echo "=== GOOFY_B200_ZEROCOPY_PAGEABLE=$z"
  GOOFY_B200_ZEROCOPY_PAGEABLE=$z tools/hostlat 8192 8192 10 2>&1 | grep -E "lib pageable|memcpy"
  GOOFY_B200_ZEROCOPY_PAGEABLE=$z python bench.py --steps 20 --warmup 5 --no-configs --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.load(sys.stdin); print('bench pageable', d['e2e']['pageable_buffers'], 'pinned ms', d['e2e']['ms_per_step'])"
done; done
