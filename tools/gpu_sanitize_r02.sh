#!/bin/bash
# compute-sanitizer (round 2): the round-1 selection of the parity tests plus the packed-RGB kernels and the alpha-stripping
# host path (zero-copy reads of pinned staging strips, the hybrid scheduler's pack ring), and the rebuilt TMA ring.
mkdir -p gpurun_out
SEL='golden_fixtures or synthetic_families or relaxed_shapes or ragged_shapes or padded_stride or uniform_batch or ragged_batch or tma_tile_path_shapes or tma_tile_path_padded or decoder or floatref or return_codes or random_shapes'
SEL2='rgb24 or host_path_float_reference'
for tool in memcheck racecheck synccheck; do
  timeout 1500 compute-sanitizer --tool $tool --report-api-errors no --error-exitcode 99 --log-file gpurun_out/sanitizer_r02_$tool.log \
      python -m pytest tests/test_gpu_parity.py tests/test_gpu_rgb24.py -m gpu -x -q -k "$SEL or $SEL2" > gpurun_out/sanitizer_r02_${tool}_pytest.log 2>&1
  echo "$tool rc=$? : $(tail -1 gpurun_out/sanitizer_r02_${tool}_pytest.log) | $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/sanitizer_r02_$tool.log | tail -1)"
done
