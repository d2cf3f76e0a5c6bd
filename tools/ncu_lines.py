"""Aggregates `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass` per CUDA source line:
executed warp instructions, thread instructions (lane utilisation) and stall samples."""
import csv
import sys
from collections import defaultdict

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(open(path)))
fname = None
hdr = None
agg = defaultdict(lambda: [0, 0, 0, ""])
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]
        continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        hdr = r
        ie = hdr.index("Instructions Executed")
        it = hdr.index("Thread Instructions Executed")
        isamp = hdr.index("# Samples")
        continue
    if hdr is None or len(r) <= it:
        continue
    try:
        line = int(r[0])
    except ValueError:
        continue
    if r[2] == "":  # the CUDA line itself (no SASS address): remember its text
        agg[(fname, line)][3] = r[1].strip()
        continue
    try:
        agg[(fname, line)][0] += int(r[ie])
        agg[(fname, line)][1] += int(r[it])
        agg[(fname, line)][2] += int(r[isamp])
    except ValueError:
        pass
tot_i = sum(v[0] for v in agg.values()) or 1
tot_s = sum(v[2] for v in agg.values()) or 1
print(f"total warp instructions {tot_i}, samples {tot_s}")
for (f, l), v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    lanes = v[1] / v[0] if v[0] else 0
    print(f"{f}:{l:<4d} {100*v[0]/tot_i:5.1f}% inst  {100*v[2]/tot_s:5.1f}% samples  {lanes:5.1f} lanes  | {v[3][:90]}")
