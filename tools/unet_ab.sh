#!/bin/bash
# Runs on the GPU box: per-layer CUDA-event times of one UNet pass under different env switches.
#   tools/unet_ab.sh "<ENV=... ENV=...>" ["<other env set>" ...]
out=gpurun_out; mkdir -p $out
i=0
for envs in "$@"; do
  echo "== $envs"
  env $envs python tools/profile_pass.py --what unet ${AB_ARGS:-} > $out/ab_$i.json 2> $out/ab_$i.err || tail -3 $out/ab_$i.err
  python - <<PY
import json
d=json.load(open("$out/ab_$i.json"))
print("unet_ms", round(d["unet_ms"],3), "sustained", round(d.get("unet_sustained_ms", 0),3), " ".join(f"{k}={v}" for k,v in d["unet_layers_ms"].items()))
PY
  i=$((i+1))
done
