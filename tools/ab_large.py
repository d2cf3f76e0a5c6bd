"""A/B of library builds on ONE box: time the reference-exact search at 1,024 / 8,192 games with CUDA events.
usage: python tools/ab_large.py lib1.so lib2.so ...   (each is copied over die_e_b200/libdiee_cuda.so in a subprocess)"""
import os, shutil, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if sys.argv[1] == "--child":
    sys.path.insert(0, ROOT)
    import numpy as np, torch
    import bench
    from die_e_b200 import _ffi as ffi
    ctx = ffi.Context(0)
    dev = torch.device("cuda", 0)
    stream = torch.cuda.Stream(device=dev); torch.cuda.set_stream(stream); ctx.set_stream(stream.cuda_stream)
    cfg = np.zeros(1, dtype=ffi.MCTS_CFG); cfg[0] = (100, 2.0, 400, 0.3, 0.25, ffi.MODE_PASS_CHILD)
    for G in (1024, 8192):
        h = bench.midgame_states(ctx, ffi, 0, G)
        ds = torch.from_numpy(h.view(np.uint8).reshape(G, 32)).to(dev); dp = torch.from_numpy(h["player"].copy()).to(dev)
        db = torch.zeros(G, dtype=torch.int32, device=dev); dst = torch.zeros(G, dtype=torch.int32, device=dev)
        res = []
        for rep in range(3):
            for i in range(3):
                ctx.mcts_search_dev(ffi.GAME_BACKGAMMON, ds.data_ptr(), G, dp.data_ptr(), cfg, bench.SEED, 0, i, db.data_ptr(), dst.data_ptr(), 0)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for i in range(5):
                ctx.mcts_search_dev(ffi.GAME_BACKGAMMON, ds.data_ptr(), G, dp.data_ptr(), cfg, bench.SEED, 0, 3 + i, db.data_ptr(), dst.data_ptr(), 0)
            e1.record(stream); e1.synchronize()
            res.append(round(G * 100 * 5 / (e0.elapsed_time(e1) / 1e3) / 1e6, 2))
        print(f"  {G} games: {res} M sims/s", flush=True)
    sys.exit(0)
orig = os.path.join(ROOT, "die_e_b200", "libdiee_cuda.so")
shutil.copy(orig, "/tmp/libdiee_keep.so")
for lib in sys.argv[1:]:
    print(lib, flush=True)
    if lib != "current":
        shutil.copy(lib, orig)
    subprocess.run([sys.executable, __file__, "--child"])
    shutil.copy("/tmp/libdiee_keep.so", orig)
