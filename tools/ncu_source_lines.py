"""Attribute an ncu SASS source page (ncu -i rep --page source --csv --print-source sass) to CUDA source
lines by aligning it with `nvdisasm --print-line-info` of the same cubin (same instruction order).
Usage: ncu_source_lines.py sass.csv lib.so kernel_substring [top_n]"""
import csv, re, subprocess, sys, tempfile, os, glob
from collections import defaultdict

sass_csv, lib, kern = sys.argv[1:4]
# "mangled|demangled": substring of the cubin symbol | substring of the kernel name in the csv
kern, kern_csv = (kern.split("|") + [kern])[:2]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, stdout=subprocess.DEVNULL)
lines_of = None
for cub in glob.glob(tmp + "/*.cubin"):
    txt = subprocess.run(["nvdisasm", "--print-line-info", cub], capture_output=True, text=True).stdout
    if kern not in txt:
        continue
    # split per function
    cur_fn, cur_line, seq = None, None, defaultdict(list)
    for ln in txt.splitlines():
        m = re.match(r"\s*\.text\.(\S+):", ln)
        if m:
            cur_fn = m.group(1); cur_line = None; continue
        m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', ln)
        if m:
            inl = "inlined" in m.group(3)
            cur_line = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m and cur_fn:
            seq[cur_fn].append((cur_line, m.group(2).strip()))
    for fn, s in seq.items():
        if kern in fn:
            lines_of = s
            break
rows = list(csv.reader(open(sass_csv)))
# a report with several launches holds one section per launch: pick section SECTION (default 0) whose name matches
secs = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name" and kern_csv in r[1]]
sel = secs[int(os.environ.get("SECTION", "0"))] if secs else 0
hi = next(i for i in range(sel, len(rows)) if rows[i] and rows[i][0] == "Address")
hdr = rows[hi]
body = []
for r in rows[hi + 1:]:
    if r and r[0] == "Kernel Name":
        break
    if len(r) == len(hdr):
        body.append(r)
ci, cs = hdr.index("Instructions Executed"), hdr.index("# Samples")
stall_cols = [j for j, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
assert lines_of and len(lines_of) == len(body), (len(lines_of or []), len(body))
agg = defaultdict(lambda: [0, 0, defaultdict(int)])
for (line, _), r in zip(lines_of, body):
    a = agg[line]
    a[0] += int(r[ci]); a[1] += int(r[cs])
    for j in stall_cols:
        v = int(r[j] or 0)
        if v: a[2][hdr[j][6:]] += v
ti, ts = sum(a[0] for a in agg.values()), sum(a[1] for a in agg.values())
src = {}
print(f"total warp-instructions {ti}, samples {ts}")
for line, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    f, n = line if line else ("?", 0)
    if f not in src:
        cand = glob.glob(os.path.join(os.path.dirname(os.path.abspath(lib)), "csrc", f))
        src[f] = open(cand[0]).read().splitlines() if cand else []
    text = src[f][n - 1].strip()[:90] if 0 < n <= len(src[f]) else ""
    st = ",".join(f"{k}:{v}" for k, v in sorted(a[2].items(), key=lambda kv: -kv[1])[:3])
    print(f"{100*a[1]/ts:5.1f}% smp {100*a[0]/ti:5.1f}% inst  {f}:{n:<4} {text}   [{st}]")
