#!/bin/bash
# Builds libsd_b200.so from the sources of a git ref (default HEAD) into tools/ab/libsd_b200_<tag>.so, for in-box A/B runs
# with tools/lib_ab.sh.  usage: tools/build_lib_at.sh <tag> [ref]
set -e
tag=$1; ref=${2:-HEAD}
root=$(cd "$(dirname "$0")/.." && pwd)
tmp=$(mktemp -d)
git -C "$root" archive "$ref" stroke_derenderer_b200/csrc include | tar -x -C "$tmp"
cd "$tmp/stroke_derenderer_b200/csrc"
for f in sd_api seg_kernels unet; do
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -fno-strict-aliasing $SD_EXTRA_NVCC_FLAGS -c $f.cu -o "$tmp/$f.o" &
done
wait
mkdir -p "$root/tools/ab"
nvcc -shared -o "$root/tools/ab/libsd_b200_$tag.so" "$tmp/sd_api.o" "$tmp/seg_kernels.o" "$tmp/unet.o" -gencode arch=compute_100a,code=sm_100a -cudart static
rm -rf "$tmp"
ls -la "$root/tools/ab/libsd_b200_$tag.so"
