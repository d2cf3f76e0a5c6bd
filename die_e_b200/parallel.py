"""Multi-GPU plumbing.  Games are independent, so self-play shards by global game id
with no data-path collective; the one exchange step is the all-gather of finished trajectories into
every rank's replay buffer at the end of a self-play wave (NCCL over NVLink on GPUs, gloo in CPU
tests), plus a weight broadcast when a new model is promoted."""
import numpy as np
import torch
import torch.distributed as dist

from . import _ffi


def shard_game_ids(n_games_total, rank, world):
    """contiguous block of global game ids owned by `rank` (streams are keyed by game id, so results do not
    depend on the number of GPUs)"""
    per = (n_games_total + world - 1) // world
    lo = min(n_games_total, rank * per)
    hi = min(n_games_total, lo + per)
    return lo, hi


def _dev(group=None):
    backend = dist.get_backend(group)
    return torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")


def allgather_trajectories(rec, pi_ids, pi_vals, group=None):
    """every rank contributes its packed records; every rank receives all of them, rank-major, with
    pi_offset rebased into the concatenated pi arrays.  Ragged sizes: counts first, then padded slabs."""
    world = dist.get_world_size(group)
    dev = _dev(group)
    counts = torch.tensor([len(rec), len(pi_ids)], dtype=torch.int64, device=dev)
    all_counts = [torch.zeros(2, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(all_counts, counts, group=group)
    all_counts = torch.stack(all_counts).cpu().numpy()
    max_rec, max_pi = int(all_counts[:, 0].max()), int(all_counts[:, 1].max())

    def gather(arr, max_n, itemsize):
        buf = np.zeros(max(1, max_n) * itemsize, dtype=np.uint8)
        raw = np.ascontiguousarray(arr).view(np.uint8).reshape(-1)
        buf[: len(raw)] = raw
        t = torch.from_numpy(buf).to(dev)
        outs = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(outs, t, group=group)
        return [o.cpu().numpy() for o in outs]
    recs = gather(rec, max_rec, _ffi.TRAJ.itemsize)
    ids = gather(pi_ids, max_pi, 2)
    vals = gather(pi_vals, max_pi, 4)
    out_rec, out_ids, out_vals, base = [], [], [], 0
    for r in range(world):
        nr, npi = int(all_counts[r, 0]), int(all_counts[r, 1])
        rr = recs[r][: nr * _ffi.TRAJ.itemsize].view(_ffi.TRAJ).copy()
        rr["pi_offset"] += base
        out_rec.append(rr)
        out_ids.append(ids[r][: npi * 2].view(np.uint16))
        out_vals.append(vals[r][: npi * 4].view(np.float32))
        base += npi
    return np.concatenate(out_rec), np.concatenate(out_ids), np.concatenate(out_vals)


def allgather_trajectories_abi(ctx, rec, pi_ids, pi_vals, rec_cap, pi_cap):
    """the same exchange through the C ABI (diee_traj_allgather: NCCL bound inside the library, no torch involved);
    `ctx` must have joined a communicator with ctx.comm_init(nranks, rank, _ffi.comm_unique_id() of rank 0)"""
    return ctx.traj_allgather(rec, pi_ids, pi_vals, rec_cap, pi_cap)


def broadcast_weights(tensors, src=0, group=None):
    """new model -> every rank (94 MB fp32 for the backgammon net)"""
    dev = _dev(group)
    out = []
    for t in tensors:
        x = torch.from_numpy(np.ascontiguousarray(t, dtype=np.float32)).to(dev)
        dist.broadcast(x, src=src, group=group)
        out.append(x.cpu().numpy())
    return out
