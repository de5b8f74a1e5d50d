"""Per-kernel SASS evidence: counts of the tcgen05 / TMEM / TMA mnemonics in the built library
(cuobjdump -sass), written as a markdown table.  Runs on the build box, no GPU needed.
    python tools/sass_summary.py [lib.so] > profiles/r02_sass_summary.md
"""
import re
import subprocess
import sys
from collections import OrderedDict
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
lib = Path(sys.argv[1]) if len(sys.argv) > 1 else ROOT / "stroke_derenderer_b200" / "libsd_b200.so"
MNEMONICS = ["UTCHMMA", "UTCHMMA.2CTA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "SYNCS", "LDGSTS", "ATOMS", "RED", "REDUX", "MATCH"]

txt = subprocess.run(["cuobjdump", "-sass", str(lib)], capture_output=True, text=True).stdout
kernels = OrderedDict()
cur = None
for ln in txt.splitlines():
    m = re.match(r"\s*Function : (\S+)", ln)
    if m:
        cur = m.group(1)
        kernels[cur] = {k: 0 for k in MNEMONICS}
        kernels[cur]["_n"] = 0
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
    if m and cur:
        op = m.group(1)
        kernels[cur]["_n"] += 1
        base = op.split(".")[0]
        if base in kernels[cur]:
            kernels[cur][base] += 1
        if op.startswith("UTCHMMA") and ".2CTA" in op:
            kernels[cur]["UTCHMMA.2CTA"] += 1


def demangle(n):
    r = subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
    r = re.sub(r"\(.*", "", r)
    return r.replace("sd::", "")


print(f"# SASS mnemonic counts per kernel of `{lib.name}` (cuobjdump -sass, sm_100a)\n")
print("UTCHMMA = tcgen05.mma, UTCBAR = tcgen05.commit, LDTM / STTM = tcgen05.ld / st (TMEM), UTMALDG / UTMASTG = TMA tensor "
      "load / store, SYNCS = mbarrier ops, ATOMS = shared-memory atomics, RED = global reductions, REDUX / MATCH = warp redux / match.\n")
print("| kernel | instr | " + " | ".join(MNEMONICS) + " |")
print("|---|---|" + "---|" * len(MNEMONICS))
for name, c in kernels.items():
    print(f"| `{demangle(name)}` | {c['_n']} | " + " | ".join(str(c[k]) if c[k] else "" for k in MNEMONICS) + " |")
tot = {k: sum(c[k] for c in kernels.values()) for k in MNEMONICS}
print("| **total** | | " + " | ".join(str(tot[k]) for k in MNEMONICS) + " |")
