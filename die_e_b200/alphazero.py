"""Host mirror of the AlphaZero self-play path: `alpha_mcts_parallel` (src/mcts/alpha_mcts.rs:91-202),
`get_prob_tensor_parallel` (src/mcts/utils.rs:42-58), `AlphaZero::self_play_parallel`
(src/alphazero/alpha_parallel.rs:101-231) and `MemoryFragment` (src/alphazero/alphazero.rs:69-73).
Training (`AlphaZero::train`, alphazero.rs:202-261) stays in tch and is out of scope; the arrays
`memory_to_arrays` returns are exactly what `save_training_data` (alphazero.rs:149-176) stores as
ps.ot / states.ot / outcomes.ot."""
import numpy as np

from . import _ffi
from .mcts import MctsConfig


class AlphaZeroConfig:
    """alphazero.rs:25-45 (fields on the self-play path)"""

    def __init__(self, temperature=1.25, num_self_play_batches=1024, learn_iterations=100, num_epochs=4,
                 training_batch_size=256, self_play_iterations=4):
        self.temperature, self.num_self_play_batches = temperature, num_self_play_batches
        self.learn_iterations, self.num_epochs = learn_iterations, num_epochs
        self.training_batch_size, self.self_play_iterations = training_batch_size, self_play_iterations

    @classmethod
    def from_config(cls, conf):
        return cls(float(conf["temperature"]), int(conf["num_self_play_batches"]), int(conf["learn_iterations"]),
                   int(conf["num_epochs"]), int(conf["training_batch_size"]), int(conf["self_play_iterations"]))


class RootChildren:
    """what consumers of the NodeStore read after alpha_mcts_parallel: each root's children
    {action_taken, visits} in legal-move order (utils.rs:42-58)"""

    def __init__(self, ids, moves, visits, counts, status):
        self.ids, self.moves, self.visits, self.counts, self.status = ids, moves, visits, counts, status

    def prob_tensor(self):
        """get_prob_tensor_parallel: dense [N,1352] f32, rows = visits / sum(visits)"""
        n = len(self.counts)
        out = np.zeros((n, _ffi.ACTION_SPACE), dtype=np.float32)
        for g in range(n):
            k = int(self.counts[g])
            if k == 0:
                out[g] = np.nan  # 0/0 in the reference
                continue
            s = np.float32(0)
            for v in self.visits[g, :k]:
                s = np.float32(s + v)
            out[g, self.ids[g, :k]] = self.visits[g, :k] / s
        return out


def alpha_mcts_parallel(states, net, mcts_config, seed=0, game_ids=None, epoch=0, ctx=None, max_nodes=0):
    """`alpha_mcts_parallel(store, states, net, mcts_config, pb)`: fills a fresh tree per game on the
    device and returns the roots' children (the store itself stays in HBM)."""
    states = np.ascontiguousarray(states, dtype=_ffi.BG_STATE).reshape(-1)
    ctx = ctx or net.ctx
    ids = np.arange(len(states), dtype=np.uint32) if game_ids is None else np.asarray(game_ids, dtype=np.uint32)
    r = ctx.alpha_search(net._net if hasattr(net, "_net") else net, states, ids, mcts_config.record(), seed, epoch, max_nodes)
    for i, st in enumerate(r[4]):
        if st != _ffi.OK:
            raise _ffi.DieeError(int(st), f"game {i}: node pool exhausted")
    return RootChildren(*r)


class MemoryFragment:
    """alphazero.rs:69-73: outcome, ps (dense [1352] f32), state ([1,6,4,6] f32)"""

    def __init__(self, outcome, ps, state):
        self.outcome, self.ps, self.state = outcome, ps, state


def memory_to_arrays(rec, pi_ids, pi_vals, ctx=None):
    """packed trajectory records -> (ps [M,1352] f32, states [M,6,4,6] f32, outcomes [M] i8)"""
    ctx = ctx or _ffi.default_context()
    m = len(rec)
    ps = np.zeros((m, _ffi.ACTION_SPACE), dtype=np.float32)
    for i in range(m):
        o, k = int(rec["pi_offset"][i]), int(rec["n_pi"][i])
        ps[i, pi_ids[o:o + k]] = pi_vals[o:o + k]
    states = ctx.bg_encode_states(np.ascontiguousarray(rec["state"])) if m else np.zeros((0, 6, 4, 6), np.float32)
    return ps, states, rec["outcome"].astype(np.int8)


# ---- replay-buffer files: what the (unchanged) tch training loop reads back ----
# `Tensor::save` (tch) is libtorch's `torch::save(tensor, path)`: a TorchScript archive holding ONE tensor under the
# key "0".  The reference writes three of them per self-play iteration (alphazero.rs:149-176):
#   ps.ot [M,1352] f32, states.ot [M,6,4,6] f32 (Tensor::concat of the [1,6,4,6] states), outcomes.ot [M] i8
# under ./data/<game>/run-<id>/lrn-<i>/sp-<j>/ (alpha_parallel.rs:18-21,43-44,60-62).  The container is pinned by files
# written with libtorch's own C++ serializer (tests/golden/ot_writer.cpp = the calls tch's at_save makes,
# tests/test_ot_fixture.py); a file written by tch itself does not exist here (no Rust toolchain).
def _save_tensor_ot(path, array):
    import torch

    class Holder(torch.nn.Module):
        pass
    m = Holder()
    t = torch.from_numpy(np.ascontiguousarray(array))
    if t.is_floating_point():
        m.register_parameter("0", torch.nn.Parameter(t, requires_grad=False))
    else:
        m.register_buffer("0", t)
    torch.jit.script(m).save(str(path))


def _load_tensor_ot(path):
    import torch
    mod = torch.jit.load(str(path), map_location="cpu")
    named = dict(list(mod.named_parameters()) + list(mod.named_buffers()))
    if len(named) != 1:
        raise ValueError(f"{path}: expected a single saved tensor, found {sorted(named)}")
    return next(iter(named.values())).detach().cpu().numpy()


def sp_dir(game_name, run_id, learn_iteration, self_play_iteration, base="."):
    """./data/<game>/run-<id>/lrn-<i>/sp-<j>  (alpha_parallel.rs:18-21,43-44,60-62)"""
    import os
    return os.path.join(base, "data", game_name, f"run-{run_id}", f"lrn-{learn_iteration}", f"sp-{self_play_iteration}")


def save_training_data(data, path, ctx=None):
    """`AlphaZero::save_training_data(&self, data: &[MemoryFragment], path)` (alphazero.rs:149-176).
    `data` is a list of MemoryFragment or the packed triple (records, pi_ids, pi_vals) of self_play_parallel(packed=True)."""
    import os
    if not os.path.exists(path):
        raise FileNotFoundError(f"path: {path} does not exist!")  # the reference panics
    if isinstance(data, tuple):
        ps, states, outcomes = memory_to_arrays(*data, ctx=ctx)
    else:
        ps = np.stack([f.ps for f in data]) if data else np.zeros((0, _ffi.ACTION_SPACE), np.float32)
        states = np.concatenate([f.state for f in data]) if data else np.zeros((0, 6, 4, 6), np.float32)
        outcomes = np.array([f.outcome for f in data], dtype=np.int8)
    _save_tensor_ot(os.path.join(path, "ps.ot"), ps.astype(np.float32))
    _save_tensor_ot(os.path.join(path, "states.ot"), states.astype(np.float32))
    _save_tensor_ot(os.path.join(path, "outcomes.ot"), outcomes.astype(np.int8))


def load_training_data(path):
    """`AlphaZero::load_training_data(path) -> Vec<MemoryFragment>` (alphazero.rs:178-200)"""
    import os
    if not os.path.exists(path):
        raise FileNotFoundError(f"path: {path} does not exist!")
    ps = _load_tensor_ot(os.path.join(path, "ps.ot"))
    states = _load_tensor_ot(os.path.join(path, "states.ot"))
    outcomes = _load_tensor_ot(os.path.join(path, "outcomes.ot"))
    return [MemoryFragment(int(outcomes[i]), ps[i], states[i:i + 1]) for i in range(len(ps))]


class AlphaZero:
    """alphazero.rs:60-67: model + configs.  Only the self-play half is served by this engine."""

    def __init__(self, model, config=None, mcts_config=None, seed=0xD1EE, ctx=None):
        self.model, self.config, self.mcts_config = model, config or AlphaZeroConfig(), mcts_config or MctsConfig()
        self.seed, self.ctx = seed, ctx or model.ctx
        self._next_game_id = 0

    def self_play_parallel(self, packed=False, max_nodes=0):
        """alpha_parallel.rs:101-231 -> Vec<MemoryFragment> (or the packed records with packed=True)"""
        n = self.config.num_self_play_batches
        rec, pi_ids, pi_vals, _ = self.ctx.selfplay_run(self.model._net, n, self.mcts_config.record(), self.config.temperature,
                                                        self.seed, self._next_game_id, max_nodes)
        self._next_game_id += n
        if packed:
            return rec, pi_ids, pi_vals
        ps, states, outcomes = memory_to_arrays(rec, pi_ids, pi_vals, self.ctx)
        return [MemoryFragment(int(outcomes[i]), ps[i], states[i:i + 1]) for i in range(len(rec))]
