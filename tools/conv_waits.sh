#!/bin/bash
# Runs on the GPU box: rebuilds the library with -DSD_CONV_STATS into a scratch copy and prints, per conv op of one
# UNet pass, the cycles each warp role spent waiting (code 1 producer, 2 MMA<-epilogue, 3 MMA<-TMA, 4 epilogue<-MMA).
set -e
rm -rf /tmp/sdstats && cp -r . /tmp/sdstats && cd /tmp/sdstats
export SD_EXTRA_NVCC_FLAGS=-DSD_CONV_STATS      # exported: the build digest covers the flags, the run below must see the same ones
python -m stroke_derenderer_b200.build --force > /dev/null 2>&1
"$@" python tools/conv_waits.py 2> /tmp/waits.err | tail -1 > /tmp/waits.json || { tail -5 /tmp/waits.err; exit 1; }
python - <<'PY'
import json
ms = json.load(open("/tmp/waits.json"))["layers_ms"]
print(f"{'op':28s} {'ms':>7s} | per-SM kcycles waited: producer  mma<-epi  mma<-tma  epi<-mma(/128)")
for ln in open("/tmp/waits.err"):
    if not ln.startswith('{"op"'): continue
    d = json.loads(ln); w = d["wait"]
    if sum(w) == 0: continue
    n = 148.0
    print(f"{d['op']:28s} {ms.get(d['op'], 0):7.4f} | {w[1]/n/1e3:9.1f} {w[2]/n/1e3:9.1f} {w[3]/n/1e3:9.1f} {w[4]/n/128/1e3:9.1f}")
PY
