"""The exchange step over NCCL through the C ABI (diee_comm_init / diee_traj_allgather / diee_net_broadcast):
two ranks, one per GPU, each in its own process.  Needs two GPUs (skipped on a one-GPU box; the gloo twin of the same
logic is tests/test_parallel_gloo.py)."""
import multiprocessing as mp

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _records(rank, n_rec):
    from die_e_b200 import _ffi
    rng = np.random.default_rng(100 + rank)
    rec = np.zeros(n_rec, dtype=_ffi.TRAJ)
    n_pi_each = rng.integers(1, 9, n_rec)
    rec["game_id"] = 1000 * rank + np.arange(n_rec)
    rec["ply"] = rng.integers(0, 200, n_rec)
    rec["outcome"] = rng.integers(-1, 2, n_rec)
    rec["n_pi"] = n_pi_each
    rec["pi_offset"] = np.concatenate([[0], np.cumsum(n_pi_each)[:-1]]) if n_rec else []
    rec["state"]["pts"] = rng.integers(-3, 4, (n_rec, 24))
    n_pi = int(n_pi_each.sum())
    return rec, rng.integers(0, 1352, n_pi).astype(np.uint16), rng.random(n_pi, dtype=np.float32)


def _worker(rank, world, uid, q):
    try:
        from die_e_b200 import _ffi
        ctx = _ffi.Context(rank)
        ctx.comm_init(world, rank, uid)
        sizes = [37, 5]   # ragged on purpose
        rec, ids, vals = _records(rank, sizes[rank])
        g_rec, g_ids, g_vals = ctx.traj_allgather(rec, ids, vals, rec_cap=100, pi_cap=1000)
        want = [_records(r, sizes[r]) for r in range(world)]
        base = 0
        off = 0
        for r in range(world):
            w_rec, w_ids, w_vals = want[r]
            got = g_rec[off:off + len(w_rec)]
            assert (got["game_id"] == w_rec["game_id"]).all() and (got["outcome"] == w_rec["outcome"]).all()
            assert got["state"].tobytes() == w_rec["state"].tobytes()
            assert (got["pi_offset"] == w_rec["pi_offset"] + base).all()
            assert (g_ids[base:base + len(w_ids)] == w_ids).all() and (g_vals[base:base + len(w_vals)] == w_vals).all()
            base += len(w_ids)
            off += len(w_rec)
        assert len(g_rec) == sum(sizes) and len(g_ids) == base
        # an empty contribution, and a capacity that is too small
        e_rec, e_ids, e_vals = ctx.traj_allgather(rec[:0] if rank else rec, ids[:0] if rank else ids, vals[:0] if rank else vals, 100, 1000)
        assert len(e_rec) == sizes[0]
        try:
            ctx.traj_allgather(rec, ids, vals, rec_cap=10, pi_cap=1000)
            raise AssertionError("expected an overflow")
        except _ffi.DieeError as err:
            assert err.code == _ffi.ERR_OVERFLOW
        # the verdict is COLLECTIVE: only rank 0's buffers are too small, yet BOTH ranks get the overflow and nobody is left
        # waiting inside the payload all-gather (round-1 advisor finding); the communicator stays usable afterwards
        try:
            ctx.traj_allgather(rec, ids, vals, rec_cap=10 if rank == 0 else 100, pi_cap=1000)
            raise AssertionError("expected an overflow on both ranks")
        except _ffi.DieeError as err:
            assert err.code == _ffi.ERR_OVERFLOW
        again = ctx.traj_allgather(rec, ids, vals, rec_cap=100, pi_cap=1000)
        assert len(again[0]) == sum(sizes)
        # weights: everybody ends up with rank 1's tensors
        tens = [np.full(7, float(rank), np.float32), np.arange(12, dtype=np.float32).reshape(3, 4) * (rank + 1)]
        ctx.net_broadcast(tens, root=1)
        assert (tens[0] == 1.0).all() and (tens[1] == np.arange(12, dtype=np.float32).reshape(3, 4) * 2).all()
        ctx.comm_destroy()
        q.put((rank, "ok"))
    except Exception as e:  # noqa: BLE001
        import traceback
        q.put((rank, "FAILED: " + repr(e) + "\n" + traceback.format_exc()))


def test_traj_allgather_and_weight_broadcast_over_nccl():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from die_e_b200 import _ffi
    uid = _ffi.comm_unique_id()
    ctxm = mp.get_context("spawn")
    q = ctxm.Queue()
    procs = [ctxm.Process(target=_worker, args=(r, 2, uid, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=300) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == {0: "ok", 1: "ok"}, res
