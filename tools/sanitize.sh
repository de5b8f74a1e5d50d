#!/bin/bash
# compute-sanitizer over the bandwidth-bound kernels (CCL / union-find, glue, tile cut, crops): memcheck, racecheck
# and synccheck on the parity tests that drive them.  Run on the GPU box: gpurun -- tools/sanitize.sh
# Logs: gpurun_out/r02_sanitize_{memcheck,racecheck,synccheck}.log (copy the summaries to profiles/).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TESTS="tests/test_gpu_seg.py::test_ccl_random_masks_vs_cv2 tests/test_gpu_seg.py::test_ccl_label_stats_fused_vs_cv2 \
tests/test_gpu_seg.py::test_ccl_labels_match_opencv_golden tests/test_gpu_seg.py::test_glue_random_values_and_threshold \
tests/test_gpu_seg.py::test_glue_other_overlaps_and_paste_table tests/test_gpu_seg.py::test_tile_extract_batch_vs_oracle \
tests/test_gpu_seg.py::test_partition_matches_reference_golden"
for tool in memcheck racecheck synccheck; do
  log=gpurun_out/r02_sanitize_${tool}.log
  timeout 900 /usr/local/cuda/bin/compute-sanitizer --tool $tool --error-exitcode 9 --print-limit 20 \
      python -m pytest $TESTS -x -q -p no:cacheprovider > $log 2>&1
  echo "[sanitize] $tool exit=$? : $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|passed|failed' $log | tr '\n' ' ')"
done
