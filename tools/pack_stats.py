"""Phase statistics of lane_pack_kernel (a -DDIEE_LANE_STATS build): python tools/pack_stats.py [games]"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import bench
from die_e_b200 import _ffi
import orc
N = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
ctx = _ffi.Context(0)
lib = _ffi.lib()
out = (ctypes.c_ulonglong * 16)()
HAVE = hasattr(lib, 'diee_debug_lane_stats')
states = bench.midgame_states(ctx, _ffi, 0, N)
cfg = orc.mcts_cfg(iterations=100, c=2.0, limit=400, mode=_ffi.MODE_PASS_CHILD)
for rep in range(2):
    if HAVE: lib.diee_debug_lane_stats(out, 1); lib.diee_debug_lane_stats2((ctypes.c_ulonglong * 16)(), 1)
    ctx.mcts_search(_ffi.GAME_BACKGAMMON, states, states["player"].copy(), cfg, 0xD1EE, 0, 0)
    tree_ms, roll_ms = ctx.search_timing()
    work = ctx.search_work()
    if HAVE: lib.diee_debug_lane_stats(out, 1)
print(f"games {N}: tree {tree_ms:.3f} ms, rollouts {roll_ms:.3f} ms, plies played {work} ({work / roll_ms / 1e6:.2f} G plies/s)")
if HAVE and out[0]:
    kinds = ("two dice", "doubles", "bar", "table", "walk", "turnover")
    print(f"batches {out[0]}, games per batch {out[1] / out[0]:.2f}, polls without a batch {out[2]} ({out[2] / out[0]:.2f} per batch)")
    print("games played per kind: " + ", ".join(f"{k} {out[8 + i] / out[1]:.3f}" for i, k in enumerate(kinds)))
    out2 = (ctypes.c_ulonglong * 16)()
    lib.diee_debug_lane_stats2(out2, 1)
    print("cycles per visit (warp 0 of every CTA): " + ", ".join(f"{k} {out2[i] / max(out2[8 + i], 1):.0f} ({out2[8 + i]})" for i, k in enumerate(kinds)))
    print("longest rollout: %d plies played" % out[15])
    print(f"next ply of the same kind: {out2[6] / max(out[1], 1):.3f} of the plies; of another closed-form kind: {out2[7] / max(out[1], 1):.3f}")
