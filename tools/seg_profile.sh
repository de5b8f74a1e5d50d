set -u
out=gpurun_out; tag=r01
for what in seg dense; do
  python tools/profile_pass.py --what $what > $out/${tag}_${what}_pass_events.json 2> $out/${tag}_${what}.err || { echo "plain $what failed"; continue; }
  cat $out/${tag}_${what}_pass_events.json | cut -c1-300
  ncu --profile-from-start off --set full --clock-control none --import-source on -f -o /tmp/${tag}_${what} \
      python tools/profile_pass.py --what $what > $out/${tag}_${what}_ncu.log 2>&1
  ncu -i /tmp/${tag}_${what}.ncu-rep --page raw --csv > $out/${tag}_${what}_raw.csv 2>> $out/${tag}_${what}_ncu.log
done
python bench.py --steps 3 --warmup 3 > $out/${tag}_bench.json 2> $out/${tag}_bench.err; tail -1 $out/${tag}_bench.err
