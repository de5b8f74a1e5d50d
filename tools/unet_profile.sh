#!/bin/bash
# Runs on the GPU box: CUDA-event times and the ncu --set full capture of one UNet pass, then the bench line.
set -u
out=gpurun_out; tag=r01
python tools/profile_pass.py --what unet > $out/${tag}_unet_pass_events.json 2> $out/${tag}_unet.err || { echo "plain unet failed"; exit 1; }
ncu --profile-from-start off --set full --clock-control none --import-source on -f -o /tmp/${tag}_unet \
    python tools/profile_pass.py --what unet > $out/${tag}_unet_ncu.log 2>&1
ncu -i /tmp/${tag}_unet.ncu-rep --page raw --csv > $out/${tag}_unet_raw.csv 2>> $out/${tag}_unet_ncu.log
python bench.py --steps 3 --warmup 3 > $out/${tag}_bench.json 2> $out/${tag}_bench.err; tail -1 $out/${tag}_bench.err
SD_BENCH_PROFILE=1 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
   --log-file $out/${tag}_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $out/${tag}_bench_under_ncu.log 2>&1
