"""Hottest CUDA source lines by stall samples, with the top stall reasons (ncu --page source --csv --print-source cuda,sass)."""
import csv
import sys
from collections import defaultdict

rows = list(csv.reader(open(sys.argv[1])))
srcdir = sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 24
fname = None
hdr = None
agg = defaultdict(lambda: [0, 0])
stall = defaultdict(lambda: defaultdict(int))
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        ie = hdr.index("Instructions Executed")
        it = hdr.index("# Samples")
        sidx = {n: i for i, n in enumerate(hdr) if n.startswith("stall_") and "Not Issued" not in n}
        continue
    if hdr is None or len(r) <= ie:
        continue
    try:
        line = int(r[0])
    except ValueError:
        continue
    if r[2] == "":
        continue
    try:
        agg[(fname, line)][0] += int(r[ie])
        agg[(fname, line)][1] += int(r[it])
        for n, i in sidx.items():
            stall[(fname, line)][n] += int(r[i] or 0)
    except ValueError:
        pass
ts = sum(v[1] for v in agg.values()) or 1
src = {}
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    if k[0] not in src:
        try:
            src[k[0]] = open(f"{srcdir}/{k[0]}").read().split("\n")
        except OSError:
            src[k[0]] = []
    text = src[k[0]][k[1] - 1].strip()[:72] if 0 < k[1] <= len(src[k[0]]) else ""
    t2 = sorted(stall[k].items(), key=lambda kv: -kv[1])[:2]
    print(f"{k[0][:16]:16s}:{k[1]:4d} {100*v[1]/ts:5.1f}% | {text:72s} | {' '.join(f'{n[6:]}={c}' for n, c in t2)}")
