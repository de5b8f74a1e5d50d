#!/bin/bash
# Runs on the GPU box: ncu --set full capture of the kernels matching <regex> in one profile pass, exported as
# details text + per-SASS-instruction CSV (for tools/ncu_source_lines.py).
#   tools/ncu_kernel.sh <tag> <kernel regex> <what: unet|seg|dense> [extra profile_pass args]
set -u
tag=$1; rx=$2; what=$3; shift 3
out=gpurun_out; mkdir -p $out
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:$rx -f -o /tmp/$tag \
    python tools/profile_pass.py --what $what "$@" > $out/${tag}_ncu.log 2>&1
ncu -i /tmp/$tag.ncu-rep --page raw --csv > $out/${tag}_raw.csv
ncu -i /tmp/$tag.ncu-rep --page source --csv --print-source sass > $out/${tag}_sass.csv 2>&1
ncu -i /tmp/$tag.ncu-rep --page details > $out/${tag}_details.txt
grep -E "Duration|Registers Per|Achieved Occ|Executed Ipc Active|Issued Instructions  |DRAM Throughput|No Eligible" $out/${tag}_details.txt
