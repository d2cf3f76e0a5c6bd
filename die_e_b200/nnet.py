"""`ResNet`: the reference's policy/value net (src/alphazero/nnet.rs) served by the CUDA engine.

Weights are a flat list of float32 arrays in the reference's registration order (nnet.rs:62-98,
ResBlock::new :37-44); they come from a tch VarStore `.ot` file (`from_path`), from a seeded
synthetic init (`ResNet.new`), or from any training job that exports that order.  Training itself
stays in tch (out of scope)."""
import re

import numpy as np

from . import _ffi

ACTION_SPACE = 1352


def layer_spec(filters=256, blocks=19):
    """[(kind, shapes...)] in registration order.  kind in {conv, bn, linear}"""
    spec = [("conv", (filters, 6, 3, 3)), ("bn", filters)]
    for _ in range(blocks):
        spec += [("conv", (filters, filters, 3, 3)), ("conv", (filters, filters, 3, 3)), ("bn", filters), ("bn", filters)]
    spec += [("conv", (32, filters, 3, 3)), ("bn", 32), ("linear", (ACTION_SPACE, 32 * 24))]
    spec += [("conv", (3, filters, 3, 3)), ("bn", 3), ("linear", (1, 3 * 24))]
    return spec


def tensor_shapes(filters=256, blocks=19):
    """shapes of the flat tensor list diee_net_create expects"""
    out = []
    for kind, sh in layer_spec(filters, blocks):
        if kind == "conv":
            out += [sh, (sh[0],)]
        elif kind == "linear":
            out += [sh, (sh[0],)]
        else:
            out += [(sh,)] * 4  # weight(gamma), bias(beta), running_mean, running_var
    return out


def synthetic_tensors(seed=0xD1EE, filters=256, blocks=19, bn_stats="identity"):
    """SURVEY 8(d) M-inputs: conv/linear weights U(+-sqrt(6/fan_in))*0.5, biases 0; BatchNorm either the
    identity (gamma=1, beta=0, mean=0, var=1) or non-trivial statistics that exercise the folding"""
    rng = np.random.default_rng(seed)
    out = []
    for kind, sh in layer_spec(filters, blocks):
        if kind in ("conv", "linear"):
            fan_in = int(np.prod(sh[1:]))
            lim = np.sqrt(6.0 / fan_in) * 0.5
            out.append(rng.uniform(-lim, lim, size=sh).astype(np.float32))
            out.append(np.zeros(sh[0], dtype=np.float32))
        elif bn_stats == "identity":
            out += [np.ones(sh, np.float32), np.zeros(sh, np.float32), np.zeros(sh, np.float32), np.ones(sh, np.float32)]
        else:
            out += [rng.uniform(0.5, 1.5, sh).astype(np.float32), rng.normal(0, 0.1, sh).astype(np.float32),
                    rng.normal(0, 0.1, sh).astype(np.float32), rng.uniform(0.5, 1.5, sh).astype(np.float32)]
    return out


# ---- tch VarStore (.ot) naming: every layer registers on the ROOT path (nnet.rs:62-96), so the
# names collide and tch disambiguates with "__<number of variables registered so far>".  The suffix is
# strictly increasing, so sorting each base name by suffix recovers registration order whatever order
# tch creates weight/bias inside one layer.  The CONTAINER is pinned by an archive written with
# libtorch's own OutputArchive (tests/test_ot_fixture.py); the suffix scheme is restated from tch 0.13.0's published
# source and cannot be run here (no Rust toolchain; the reference ships no .ot file).
def _tch_names(filters, blocks):
    names, count, seen = [], 0, set()

    def reg(base):
        nonlocal count
        name = base if base not in seen else f"{base}__{count}"
        seen.add(name)
        seen.add(base)
        count += 1
        return name
    for kind, _ in layer_spec(filters, blocks):
        if kind in ("conv", "linear"):
            b = reg("bias")
            w = reg("weight")
            names += [w, b]
        else:
            rm, rv, w, b = reg("running_mean"), reg("running_var"), reg("weight"), reg("bias")
            names += [w, b, rm, rv]
    return names


def save_ot(path, tensors, filters=256, blocks=19):
    """writes a TorchScript archive of named tensors the way tch's VarStore::save does"""
    import torch

    class Holder(torch.nn.Module):
        pass
    m = Holder()
    for name, t in zip(_tch_names(filters, blocks), tensors):
        m.register_parameter(name, torch.nn.Parameter(torch.from_numpy(np.ascontiguousarray(t)), requires_grad=False))
    torch.jit.script(m).save(path)


def load_ot(path):
    """-> (tensors in registration order, filters, blocks)"""
    import torch
    mod = torch.jit.load(path, map_location="cpu")
    named = {k: v.detach().cpu().numpy() for k, v in list(mod.named_parameters()) + list(mod.named_buffers())}
    groups = {"weight": [], "bias": [], "running_mean": [], "running_var": []}
    for k, v in named.items():
        m = re.fullmatch(r"(?:.*[.|])?(weight|bias|running_mean|running_var)(?:__(\d+))?", k)
        if not m:
            raise ValueError(f"unexpected variable name {k!r} in {path}")
        groups[m.group(1)].append((int(m.group(2)) if m.group(2) else -1, v))
    for g in groups.values():
        g.sort(key=lambda kv: kv[0])
    n_bn = len(groups["running_mean"])
    blocks = (n_bn - 3) // 2
    filters = int(groups["running_mean"][0][1].shape[0])
    it = {k: iter(v for _, v in g) for k, g in groups.items()}
    out = []
    for kind, sh in layer_spec(filters, blocks):
        w, b = next(it["weight"]), next(it["bias"])
        want = tuple(sh) if kind != "bn" else (sh,)
        if tuple(w.shape) != want:
            raise ValueError(f"{path}: expected a {kind} weight of shape {want}, found {tuple(w.shape)}")
        out += [w, b]
        if kind == "bn":
            out += [next(it["running_mean"]), next(it["running_var"])]
    return [np.ascontiguousarray(t, dtype=np.float32) for t in out], filters, blocks


def states_from_tensor(x):
    """inverse of as_tensor (backgammon_logic.rs:198-252): f32 [N,6,4,6] -> packed states"""
    x = np.asarray(x, dtype=np.float32).reshape(-1, 6, 24)
    s = np.zeros(len(x), dtype=_ffi.BG_STATE)
    s["pts"] = x[:, 0].astype(np.int8)
    s["player"] = x[:, 1, 0].astype(np.int8)
    s["bar"][:, 0], s["bar"][:, 1] = x[:, 2, 0], x[:, 2, 12]
    s["off"][:, 0], s["off"][:, 1] = x[:, 3, 0], x[:, 3, 12]
    s["roll"][:, 0], s["roll"][:, 1] = x[:, 4, 0], x[:, 4, 12]
    s["second"] = x[:, 5, 0].astype(np.uint8)
    return s


class ResNet:
    """nnet.rs:48-155.  forward_t / forward_policy take packed states (or the reference's
    [N,6,4,6] float tensor, which is converted back losslessly)."""

    def __init__(self, tensors, filters=256, blocks=19, ctx=None):
        self.tensors, self.filters, self.blocks = tensors, filters, blocks
        self.ctx = ctx or _ffi.default_context()
        self._net = _ffi.Net(self.ctx, tensors)

    @classmethod
    def new(cls, seed=0xD1EE, filters=256, blocks=19, bn_stats="identity", ctx=None):
        return cls(synthetic_tensors(seed, filters, blocks, bn_stats), filters, blocks, ctx)

    @classmethod
    def from_path(cls, model_path, ctx=None):  # nnet.rs:109-118
        tensors, filters, blocks = load_ot(str(model_path))
        return cls(tensors, filters, blocks, ctx)

    def save(self, path):
        save_ot(str(path), self.tensors, self.filters, self.blocks)

    def _states(self, xs):
        xs = np.asarray(xs)
        return xs if xs.dtype == _ffi.BG_STATE else states_from_tensor(xs)

    def forward_t(self, xs, train=False):  # nnet.rs:120-133
        assert not train, "training stays in tch (out of scope)"
        policy, value = self._net.forward(self._states(xs))
        return policy, value.reshape(-1, 1)

    def forward_policy(self, xs, train=False):  # nnet.rs:150-155
        return self.forward_t(xs, train)[0]
