"""Issue-side reading of a kernel from a committed ncu export: profiles/<out>.json, which bench.py attaches to the
`roofline` of the line as `roofline.issue` (ncu replays kernels, so this is never measured inside bench.py).

    ncu -i capture.ncu-rep --page raw --csv > raw.csv
    python tools/ncu_issue.py raw.csv <plies the launch executed> profiles/r02_lane_run_rollouts_issue.json [capture name]

`plies` = the `rollout_plies_played_per_search` (or plies per step for --workload playout) that bench.py printed for the
same configuration WITHOUT ncu; the counters are per launch."""
import csv
import json
import sys

rows = list(csv.reader(open(sys.argv[1])))
plies = float(sys.argv[2])
h, v = rows[0], rows[2]


def get(name):
    return float(v[h.index(name)].replace(",", ""))


inst = get("smsp__inst_executed.sum")
lanes = get("smsp__thread_inst_executed_per_inst_executed.ratio")
issue = get("smsp__issue_active.avg.pct_of_peak_sustained_active") / 100.0
out = {
    "bound": "issue",
    "kernel": v[h.index("Kernel Name")] if "Kernel Name" in h else None,
    "issue_active": round(issue, 4),                       # share of cycles an SM sub-partition issues an instruction
    "lanes_per_inst": round(lanes, 2),                     # active threads per warp instruction, of 32
    "thread_inst_frac": round(issue * lanes / 32.0, 4),    # thread instructions issued / the machine's thread-instruction rate
    "warp_inst": int(inst),
    "plies_executed": int(plies),
    "warp_inst_per_ply": round(inst / plies, 1),
    "warps_active_frac": round(get("sm__warps_active.avg.pct_of_peak_sustained_active") / 100.0, 4),
    "registers_per_thread": int(get("launch__registers_per_thread")),
    "gpu_time_ms_under_ncu": round(get("gpu__time_duration.sum") * (1e-6 if "ns" in rows[1][h.index("gpu__time_duration.sum")] else 1e-3 if "us" in rows[1][h.index("gpu__time_duration.sum")] else 1.0), 4),
    "source": (sys.argv[4] if len(sys.argv) > 4 else sys.argv[1]) + " (ncu --set full --clock-control none)",
}
json.dump(out, open(sys.argv[3], "w"), indent=1)
print(json.dumps(out))
