#!/bin/bash
# Runs on the GPU box: the ncu evidence of a round.  tools/round_profile.sh <tag, e.g. r02> [tiles per UNet pass, default 256]
#   1. plain bench (must exit 0), then the launch list of ONE resident step of the same command
#   2. ncu --set full of one UNet pass and one segmentation pass (text-like and dense)
set -u
tag=${1:-r01}; out=gpurun_out; mkdir -p $out
python bench.py --steps 2 --warmup 3 > $out/${tag}_bench.json 2> $out/${tag}_bench.err || { echo "bench failed"; tail -5 $out/${tag}_bench.err; exit 1; }
SD_BENCH_PROFILE=1 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
   --log-file $out/${tag}_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-api > $out/${tag}_bench_under_ncu.log 2>&1
python tools/summarize_launches.py $out/${tag}_launches.csv $out/${tag}_launches_summary.md | head -30
tiles=${2:-256}
for what in unet seg dense; do
  extra=""; [ $what = unet ] && extra="--tiles $tiles --lines 32"
  python tools/profile_pass.py --what $what $extra > $out/${tag}_${what}_pass_events.json 2> $out/${tag}_${what}.err || { echo "plain $what failed"; continue; }
  ncu --profile-from-start off --set full --clock-control none --import-source on -f -o /tmp/${tag}_${what} \
      python tools/profile_pass.py --what $what $extra > $out/${tag}_${what}_ncu.log 2>&1
  ncu -i /tmp/${tag}_${what}.ncu-rep --page raw --csv > $out/${tag}_${what}_raw.csv 2>> $out/${tag}_${what}_ncu.log
done
ls -la $out | grep ${tag}_
