"""Summarise an `ncu -i x.ncu-rep --page raw --csv` dump: one row per launch with duration, tensor-pipe
activity, DRAM traffic and throughput.  Usage: summarize_ncu_raw.py raw.csv out.md [names.json [traffic.json kernel_regex [tiles_per_pass]]]
names.json (optional): list of stage names in launch order (e.g. the engine's layer list)."""
import csv
import json
import re
import sys

COLS = [
    ("gpu__time_duration.sum", "us", 1e-3),
    # tcgen05 kernels: sm__mem_tensor_cycles_active is the counter that tracks the UMMA datapath (it equals executed
    # MACs / 4096 per SM-cycle on every conv launch); the *_realtime tensor counters return stale values under replay
    ("sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor pipe %", 1),
    ("dram__bytes_read.sum", "DRAM rd MB", 1e-6),
    ("dram__bytes_write.sum", "DRAM wr MB", 1e-6),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %", 1),
    ("lts__t_bytes.sum", "L2 MB", 1e-6),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM %", 1),
    ("launch__registers_per_thread", "regs", 1),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ %", 1),
]


def num(s):
    try:
        return float(s.replace(",", ""))
    except Exception:
        return float("nan")


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr, units, data = rows[hi], rows[hi + 1], rows[hi + 2:]
    idx = {h: i for i, h in enumerate(hdr)}
    names = json.load(open(sys.argv[3])) if len(sys.argv) > 3 else None
    have = [(c, t, s) for c, t, s in COLS if c in idx]
    out = ["| # | stage | kernel | grid | " + " | ".join(t for _, t, _ in have) + " |", "|---" * (4 + len(have)) + "|"]
    tot_us = 0.0
    for n, r in enumerate(data):
        if len(r) < len(hdr):
            continue
        k = re.sub(r"\(.*", "", r[idx["Kernel Name"]])
        k = k.replace("void ", "").replace("sd::", "")
        vals = []
        for c, t, s in have:
            v = num(r[idx[c]])
            u = units[idx[c]]
            if c == "gpu__time_duration.sum":
                v = v / 1000 if u in ("ns", "nsecond") else (v if u in ("us", "usecond") else v * 1000)
                tot_us += v
                vals.append(f"{v:.1f}")
            elif "bytes" in c:
                mul = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1e-6)
                vals.append(f"{v * mul:.2f}")
            else:
                vals.append(f"{v:.1f}")
        stage = names[n] if names and n < len(names) else ""
        out.append(f"| {n} | {stage} | `{k}` | {r[idx['Grid Size']]} | " + " | ".join(vals) + " |")
    out.insert(0, f"# ncu --set full, per launch ({len(data)} launches, {tot_us / 1000:.2f} ms under ncu: serialised, cold-ish cache; "
                  f"compare shares and ratios, not absolutes)\n")
    open(sys.argv[2], "w").write("\n".join(out) + "\n")
    print("\n".join(out))
    if len(sys.argv) > 4:   # traffic json: DRAM bytes per launch of the kernels whose name matches argv[5] (regex)
        rx = re.compile(sys.argv[5] if len(sys.argv) > 5 else ".")
        tot = n = 0
        for r in data:
            if len(r) < len(hdr) or not rx.search(r[idx["Kernel Name"]]):
                continue
            b = 0.0
            for c in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                mul = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(units[idx[c]], 1.0)
                b += num(r[idx[c]]) * mul
            tot += b; n += 1
        tiles = int(sys.argv[6]) if len(sys.argv) > 6 else 256      # tiles of the captured pass (bench.py scales by it)
        json.dump({"kernels": rx.pattern, "launches": n, "dram_bytes_total": tot, "dram_bytes_per_launch": tot / max(n, 1), "tiles_per_pass": tiles,
                   "source": "ncu --set full --clock-control none, dram__bytes_read.sum + dram__bytes_write.sum"},
                  open(sys.argv[4], "w"), indent=1)


if __name__ == "__main__":
    main()
