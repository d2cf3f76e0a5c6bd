#!/bin/bash
timeout 300 python -m pytest tests/test_gpu_mcts.py -q -x 2>&1 | tail -8
for n in 1024 16384 65536; do for pe in 0 1; do echo -n "games $n PERSISTENT=$pe: "; DIEE_CC_PERSISTENT=$pe timeout 200 python - <<PY
import os, sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np, torch
import bench, orc
from die_e_b200 import _ffi as ffi
ctx = ffi.Context(0)
n = $n
cfg = orc.mcts_cfg(iterations=100, c=2.0, limit=400, mode=ffi.MODE_PASS_CHILD | ffi.MODE_ROLLOUT_CHECK_CURRENT)
h = bench.midgame_states(ctx, ffi, 0, n)
dev = torch.device("cuda:0")
ds = torch.from_numpy(h.view(np.uint8).reshape(n, 32)).to(dev); dp = torch.from_numpy(h["player"].copy()).to(dev)
db = torch.zeros(n, dtype=torch.int32, device=dev); dst = torch.zeros(n, dtype=torch.int32, device=dev)
import time
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    ctx.mcts_search_dev(ffi.GAME_BACKGAMMON, ds.data_ptr(), n, dp.data_ptr(), cfg, 0xD1EE, 0, rep, db.data_ptr(), dst.data_ptr(), 0)
    ctx.sync(); dt = time.perf_counter() - t0
print(f"{dt*1e3:.2f} ms, {n*100/dt/1e6:.2f} M simulations/s, status sum {int(dst.abs().sum())}, best checksum {int(db.to(torch.int64).sum())}")
PY
done; done
