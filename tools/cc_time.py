"""times one check-current search of $GAMES games (environment selects the form): python tools/cc_time.py"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import bench, orc
from die_e_b200 import _ffi as ffi
ctx = ffi.Context(0)
n = int(os.environ.get("GAMES", "16384"))
cfg = orc.mcts_cfg(iterations=100, c=2.0, limit=400, mode=ffi.MODE_PASS_CHILD | ffi.MODE_ROLLOUT_CHECK_CURRENT)
h = bench.midgame_states(ctx, ffi, 0, n)
dev = torch.device("cuda:0")
ds = torch.from_numpy(h.view(np.uint8).reshape(n, 32)).to(dev); dp = torch.from_numpy(h["player"].copy()).to(dev)
db = torch.zeros(n, dtype=torch.int32, device=dev); dst = torch.zeros(n, dtype=torch.int32, device=dev)
best = 1e9
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    ctx.mcts_search_dev(ffi.GAME_BACKGAMMON, ds.data_ptr(), n, dp.data_ptr(), cfg, 0xD1EE, 0, 7, db.data_ptr(), dst.data_ptr(), 0)
    ctx.sync(); best = min(best, time.perf_counter() - t0)
print(f"{best*1e3:.2f} ms, {n*100/best/1e6:.2f} M simulations/s, best checksum {int(db.to(torch.int64).sum())}, plies {ctx.search_work()}")
