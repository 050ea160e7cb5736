#!/bin/bash
# compute-sanitizer over the small parity tests (memcheck: out-of-bounds / misaligned; racecheck: shared-memory hazards
# in the TMA ring and the control-table staging; synccheck: barrier misuse).
mkdir -p gpurun_out
SEL='golden_fixtures or synthetic_families or relaxed_shapes or ragged_shapes or padded_stride or uniform_batch or ragged_batch or tma_tile_path_shapes or tma_tile_path_padded or decoder or floatref or return_codes or random_shapes'
for tool in memcheck racecheck synccheck; do
  timeout 1500 compute-sanitizer --tool $tool --report-api-errors no --error-exitcode 99 --log-file gpurun_out/sanitizer_$tool.log \
      python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "$SEL" > gpurun_out/sanitizer_${tool}_pytest.log 2>&1
  echo "$tool rc=$? : $(tail -1 gpurun_out/sanitizer_${tool}_pytest.log) | $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/sanitizer_$tool.log | tail -1)"
done
