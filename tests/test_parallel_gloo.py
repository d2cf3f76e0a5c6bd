"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: game-id sharding and the trajectory all-gather."""
import os
import socket

import numpy as np
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _make_shard(rank, n_rec):
    from die_e_b200 import _ffi
    rng = np.random.default_rng(100 + rank)
    rec = np.zeros(n_rec, dtype=_ffi.TRAJ)
    n_pi = rng.integers(1, 20, size=n_rec)
    rec["n_pi"] = n_pi
    rec["pi_offset"] = np.concatenate([[0], np.cumsum(n_pi)[:-1]]) if n_rec else []
    rec["game_id"] = rank * 1000 + np.arange(n_rec)
    rec["outcome"] = rng.integers(-1, 2, size=n_rec)
    rec["state"]["pts"] = rng.integers(-5, 6, size=(n_rec, 24))
    tot = int(n_pi.sum())
    return rec, rng.integers(0, 1352, size=tot).astype(np.uint16), rng.random(tot).astype(np.float32)


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from die_e_b200 import parallel
    rec, ids, vals = _make_shard(rank, [7, 3][rank])           # ragged shard sizes
    arec, aids, avals = parallel.allgather_trajectories(rec, ids, vals)
    w = parallel.broadcast_weights([np.full(5, rank + 1, np.float32), np.arange(3, dtype=np.float32) * (rank + 1)], src=1)
    q.put((rank, arec.tobytes(), aids.tobytes(), avals.tobytes(), [x.tolist() for x in w]))
    dist.destroy_process_group()


def test_allgather_trajectories_and_broadcast():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in procs])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    from die_e_b200 import _ffi
    shards = [_make_shard(r, [7, 3][r]) for r in range(2)]
    want_rec = np.concatenate([s[0] for s in shards]).copy()
    want_rec["pi_offset"][7:] += len(shards[0][1])
    want_ids = np.concatenate([s[1] for s in shards])
    want_vals = np.concatenate([s[2] for s in shards])
    for rank, rec_b, ids_b, vals_b, w in res:
        assert rec_b == want_rec.tobytes() and ids_b == want_ids.tobytes() and vals_b == want_vals.tobytes()
        assert w == [[2.0] * 5, [0.0, 2.0, 4.0]]
    # each record's pi slice is still its own after rebasing
    rec = np.frombuffer(res[0][1], dtype=_ffi.TRAJ)
    ids = np.frombuffer(res[0][2], dtype=np.uint16)
    o, k = int(rec["pi_offset"][8]), int(rec["n_pi"][8])
    so, sk = int(shards[1][0]["pi_offset"][1]), int(shards[1][0]["n_pi"][1])
    assert k == sk and (ids[o:o + k] == shards[1][1][so:so + sk]).all()


def test_shard_game_ids():
    from die_e_b200 import parallel
    for total, world in [(131072, 8), (10, 4), (3, 8)]:
        spans = [parallel.shard_game_ids(total, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == total
        assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
    assert parallel.shard_game_ids(131072, 3, 8) == (49152, 65536)
