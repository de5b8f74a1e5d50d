"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel name,
launch count, total/mean device time and share.  Usage: summarize_launches.py in.csv out.md"""
import csv, re, sys
from collections import defaultdict
rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    if r.get("Metric Name") == "gpu__time_duration.sum":
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        us = v / 1000 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1000)
        name = re.sub(r"\(.*", "", r["Kernel Name"])
        rows.append((name, us, r["Grid Size"], r["Block Size"]))
agg = defaultdict(lambda: [0, 0.0])
for n, us, *_ in rows:
    agg[n][0] += 1; agg[n][1] += us
tot = sum(v[1] for v in agg.values())
out = [f"# ncu launch list summary ({len(rows)} launches, {tot/1000:.2f} ms of kernel time; cold-cache, serialised: compare SHARES)\n",
       "| kernel | launches | total ms | mean us | share |", "|---|---:|---:|---:|---:|"]
for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    out.append(f"| `{n}` | {c} | {t/1000:.3f} | {t/c:.1f} | {100*t/tot:.1f}% |")
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write("\n".join(out) + "\n")
print("\n".join(out))
