#!/bin/bash
# Runs on the GPU box (under gpurun): ncu captures of one warmed pass.
#   tools/gpu_profile.sh <tag> <what: unet|seg|dense> [extra profile_pass args]
# Writes gpurun_out/<tag>_<what>.{json (plain run, CUDA-event times), raw.csv (ncu --set full, per launch),
# launches.csv}.  The plain run goes first and must exit 0 (B200_PROFILING.md).
set -u
tag=$1; what=$2; shift 2
out=gpurun_out
mkdir -p $out
python tools/profile_pass.py --what $what "$@" > $out/${tag}_${what}.json 2> $out/${tag}_${what}.err || { echo "plain run failed"; tail -5 $out/${tag}_${what}.err; exit 1; }
ncu --profile-from-start off --set full --clock-control none --import-source on -f -o /tmp/${tag}_${what} \
    python tools/profile_pass.py --what $what "$@" > $out/${tag}_${what}_ncu.log 2>&1
echo "ncu rc=$?"
ncu -i /tmp/${tag}_${what}.ncu-rep --page raw --csv > $out/${tag}_${what}_raw.csv 2>> $out/${tag}_${what}_ncu.log
ls -la /tmp/${tag}_${what}.ncu-rep
sz=$(stat -c %s /tmp/${tag}_${what}.ncu-rep)
if [ "$sz" -lt 30000000 ]; then cp /tmp/${tag}_${what}.ncu-rep $out/; fi
