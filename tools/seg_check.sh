#!/bin/bash
# Runs on the GPU box (under gpurun): parity tests of the bandwidth-bound stages, then CUDA-event times and an
# ncu launch list (cold-cache, serialised: compare shares) of one text-like and one dense pass.
#   tools/seg_check.sh <tag>
set -u
tag=${1:-seg}
out=gpurun_out
mkdir -p $out
python -m pytest tests/test_gpu_seg.py -x -q > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $out/${tag}_pytest.log
for what in seg dense; do
  python tools/profile_pass.py --what $what > $out/${tag}_${what}.json 2> $out/${tag}_${what}.err || { echo "plain $what failed"; tail -5 $out/${tag}_${what}.err; continue; }
  cat $out/${tag}_${what}.json
  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
      --log-file $out/${tag}_${what}_launches.csv python tools/profile_pass.py --what $what > /dev/null 2>&1
  python tools/summarize_launches.py $out/${tag}_${what}_launches.csv
done
