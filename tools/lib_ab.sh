#!/bin/bash
# Runs on the GPU box: A/B of two builds of libsd_b200.so in ONE box (boxes of the pool differ by +-3 %, more than most
# kernel changes).  tools/lib_ab.sh old.so new.so [rounds]   -- alternates old / new UNet passes (burst, sustained, per layer)
old=$1; new=$2; rounds=${3:-2}; out=gpurun_out; mkdir -p $out
dst=stroke_derenderer_b200/libsd_b200.so
cp $dst /tmp/lib_keep.so
for r in $(seq 1 $rounds); do
  for which in old new; do
    [ $which = old ] && cp $old $dst || cp $new $dst
    python tools/profile_pass.py --what unet --tiles 256 --lines 32 --sustain-reps 30 > $out/ab_${which}_$r.json 2> $out/ab_${which}_$r.err || tail -3 $out/ab_${which}_$r.err
  done
done
cp /tmp/lib_keep.so $dst
python - <<PY
import json
R = $rounds
acc = {}
for which in ("old", "new"):
    runs = [json.load(open("$out/ab_%s_%d.json" % (which, r))) for r in range(1, R + 1)]
    acc[which] = runs
    print(which, "burst ms", [round(d["unet_ms"], 3) for d in runs], "sustained ms", [round(d["unet_sustained_ms"], 3) for d in runs])
names = list(acc["old"][0]["unet_layers_ms"])
print("%-24s %8s %8s %7s" % ("layer (ms, mean)", "old", "new", "new/old"))
for n in names:
    o = sum(d["unet_layers_ms"][n] for d in acc["old"]) / R
    w = sum(d["unet_layers_ms"][n] for d in acc["new"]) / R
    print("%-24s %8.4f %8.4f %7.3f" % (n, o, w, w / o))
PY
