"""prints the timeline of one sliced search (needs a -DDIEE_TRACE build of the library): tools/variant_run-style usage
   DIEE_TREE_SMS=64 DIEE_SEARCH_SLICES=4 python tools/trace_run.py"""
import ctypes, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from die_e_b200 import _ffi as ffi
ctx = ffi.Context(0)
G = int(os.environ.get("G", 1024))
h = bench.midgame_states(ctx, ffi, 0, G)
cfg = np.zeros(1, dtype=ffi.MCTS_CFG); cfg[0] = (100, 2.0, 400, 0.3, 0.25, ffi.MODE_PASS_CHILD)
for i in range(4):
    ctx.mcts_search(ffi.GAME_BACKGAMMON, h, h["player"].copy(), cfg, bench.SEED, 0, i)
out = np.zeros(64, dtype=np.float32)
k = ffi.lib().diee_debug_search_trace(out.ctypes.data_as(ctypes.c_void_p), 64)
for s in range(k // 4):
    t = out[4 * s:4 * s + 4]
    print(f"slice {s}: tree {t[0]:.3f} -> {t[1]:.3f} ms ({t[1]-t[0]:.3f})   rollouts {t[2]:.3f} -> {t[3]:.3f} ms ({t[3]-t[2]:.3f})")
