#!/bin/bash
# Runs on the GPU box: rebuilds the library with -DSD_BOUNDS_CHECK into a scratch copy (every index of the CCL kernels is
# checked against its array, a violation prints the site and traps) and runs the CCL / partition parity tests under it.
# Stands in for compute-sanitizer memcheck, which is closed on this pool (profiles/r02_sanitize.md).
set -u
rm -rf /tmp/sdbounds && cp -r . /tmp/sdbounds && cd /tmp/sdbounds
export SD_EXTRA_NVCC_FLAGS=-DSD_BOUNDS_CHECK        # exported: the build digest covers the flags
python -m stroke_derenderer_b200.build --force > /tmp/sdbounds_build.log 2>&1 || { echo "build failed"; tail -5 /tmp/sdbounds_build.log; exit 1; }
echo "flags: $(grep -c -- -DSD_BOUNDS_CHECK /tmp/sdbounds_build.log) nvcc commands carry -DSD_BOUNDS_CHECK; check strings in the library: $(strings stroke_derenderer_b200/libsd_b200.so | grep -c 'CW_CHECK failed')"
python -m pytest tests/test_gpu_seg.py -q -x -p no:cacheprovider -k "ccl or partition or full_size or islands" 2>&1 | tail -15
echo "CW_CHECK sites compiled in: $(grep -c CW_CHECK stroke_derenderer_b200/csrc/ccl_warp.cuh)"
