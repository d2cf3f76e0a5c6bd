"""Per-function roll-up of `ncu --page source --csv --print-source cuda,sass` for bg_lane.cuh-style inlined code:
instructions and lane utilisation per LANE_HD function (by source line ranges)."""
import csv
import re
import sys
from collections import defaultdict

path, header = sys.argv[1], sys.argv[2]
rows = list(csv.reader(open(path)))
fname = None
hdr = None
agg = defaultdict(lambda: [0, 0])
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        ie = hdr.index("Instructions Executed")
        it = hdr.index("Thread Instructions Executed")
        continue
    if hdr is None or len(r) <= it or r[2] == "":
        continue
    try:
        agg[(fname, int(r[0]))][0] += int(r[ie])
        agg[(fname, int(r[0]))][1] += int(r[it])
    except ValueError:
        pass
src = open(header).read().split("\n")
funcs = []
for i, l in enumerate(src, 1):
    m = re.match(r"^LANE_HD [\w:<> ]+?[ \*&](\w+)\(", l)
    if m:
        funcs.append((i, m.group(1)))
hname = header.split("/")[-1]


def fn(line):
    name = "?"
    for s, n in funcs:
        if s <= line:
            name = n
    return name


tot = sum(v[0] for v in agg.values())
byf = defaultdict(lambda: [0, 0])
for (f, l), v in agg.items():
    key = fn(l) if f == hname else f
    byf[key][0] += v[0]
    byf[key][1] += v[1]
print("total warp instructions", tot)
for k, v in sorted(byf.items(), key=lambda kv: -kv[1][0])[:30]:
    print(f"{k:24s} {100*v[0]/tot:5.1f}%  lanes {v[1]/max(v[0],1):5.1f}")
