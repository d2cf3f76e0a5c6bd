"""GPU tests of the policy/value net (tcgen05 implicit-GEMM convolutions) through the C ABI.

Floating point: the product computes in bf16 with fp32 accumulation.  Two bars are asserted:
  (1) against the torch oracle that follows the SAME numeric recipe (net_oracle.forward_bf16_emulated:
      folded BatchNorm, bf16 weights/activations, wide accumulation) the outputs agree to within
      accumulation-order noise for shallow nets (value |dv| <= 2e-3, policy |dp| <= 0.5 % of the row
      maximum for <= 2 blocks); through all 39 convolutions single bf16 rounding flips get amplified, so
      the 19-block bar is the bf16 noise floor itself (|dv| <= 6e-2, |dp| <= 10 % of the row maximum);
  (2) against the fp64 oracle of the reference architecture the bf16 error itself is bounded
      (value |dv| <= 0.1, policy total-variation distance <= 0.05) and printed.
north_star's 1e-5 (relative, vs the fp32 reference) is NOT met by the bf16 path and is not claimed for it.
The modes inside the tolerance are DIEE_NET_FP32 (fp32 FMAs on the CUDA cores, the reference's own arithmetic) and
DIEE_NET_SPLIT3 (tensor cores: three bf16 planes per operand, the leading one an integer digit whose products are
accumulated exactly): test_fp32_mode_matches_fp32_reference asserts both against the fp64 oracle next to the
reference's own fp32 rounding cost."""
import numpy as np
import pytest

import net_oracle
import positions

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from die_e_b200 import _ffi
    return _ffi.Context(0)


def _inputs(oracle, n, seed=1):
    states = positions.midgame_positions(seed=seed, n=n, max_adv=100)
    x = np.concatenate([oracle.bg_as_tensor(states[i:i + 1]) for i in range(n)])
    return states, x


@pytest.mark.parametrize("filters,blocks,n,bn", [(128, 1, 20, "identity"), (256, 2, 37, "random"), (256, 19, 48, "random")])
def test_forward_matches_oracles(ctx, oracle, filters, blocks, n, bn):
    from die_e_b200 import _ffi, nnet
    tens = nnet.synthetic_tensors(seed=7, filters=filters, blocks=blocks, bn_stats=bn)
    net = _ffi.Net(ctx, tens)
    net.set_precision(_ffi.NET_BF16)
    states, x = _inputs(oracle, n)
    p, v = net.forward(states)
    assert np.isfinite(p).all() and np.isfinite(v).all()
    assert np.allclose(p.sum(1), 1.0, atol=1e-4)
    pe, ve = net_oracle.forward_bf16_emulated(tens, x, blocks)
    p64, v64 = net_oracle.forward(tens, x, blocks)
    d_emul_v = np.abs(v - ve).max()
    d_emul_p = (np.abs(p - pe).max(1) / pe.max(1)).max()
    d64_v = np.abs(v - v64).max()
    tv64 = 0.5 * np.abs(p - p64).sum(1).max()
    print(f"\n[net F={filters} B={blocks} n={n}] vs bf16-emulated: dv={d_emul_v:.2e} dp/max={d_emul_p:.2e} | "
          f"vs fp64: dv={d64_v:.2e} TV={tv64:.2e} (oracle's own bf16 cost: dv={np.abs(ve - v64).max():.2e})")
    if blocks <= 2:
        assert d_emul_v <= 2e-3 and d_emul_p <= 5e-3
    else:
        assert d_emul_v <= 6e-2 and d_emul_p <= 1e-1
    assert d64_v <= 0.1 and tv64 <= 0.05
    assert (p.argmax(1) == pe.argmax(1)).mean() >= 0.9
    net.close()


@pytest.mark.parametrize("filters,blocks,n,bn", [(128, 1, 20, "identity"), (256, 2, 37, "random"), (256, 19, 48, "random")])
def test_fp32_mode_matches_fp32_reference(ctx, oracle, filters, blocks, n, bn):
    """N1 in the parity mode (DIEE_NET_FP32) and in the 24-bit tensor-core mode (DIEE_NET_SPLIT3).

    Tolerance.  north_star asks for 1e-5 relative on the value estimate against the fp32 reference.  Two correct
    fp32 forwards differ by their summation order, so the bar is written against the fp64 oracle and next to
    the reference's OWN fp32 rounding cost (torch fp32 vs fp64, printed): the fp32 mode must be within
    max(1e-5, 4 x that cost) on the value (relative to max(|v|, 0.1): the value head is a 72-term sum that cancels, so an absolute floor of 1e-6 on a quantity in [-1, 1]) and on the policy (fraction of the row
    maximum -- a softmax row reaches down to 1e-9, a per-entry relative bar means nothing there).  With these
    synthetic weights the reference's own cost is 3e-7 at 2 blocks and 1.6e-5 at 19 blocks.
    SPLIT3 -- the tensor-core mode -- is held to the SAME bar: its full-size products are accumulated exactly (integer
    digit planes, csrc/net_kernels.cu) and only terms 2^-7 and smaller meet the tensor core's truncating accumulation."""
    import torch
    from die_e_b200 import _ffi, nnet
    tens = nnet.synthetic_tensors(seed=7, filters=filters, blocks=blocks, bn_stats=bn)
    net = _ffi.Net(ctx, tens)
    states, x = _inputs(oracle, n)
    p64, v64 = net_oracle.forward(tens, x, blocks)
    p32, v32 = net_oracle.forward(tens, x, blocks, dtype=torch.float32)   # what the reference (tch, fp32) computes

    def err(p, v):
        return ((np.abs(v - v64) / np.maximum(np.abs(v64), 1e-1)).max(), (np.abs(p - p64).max(1) / p64.max(1)).max())
    ref_v, ref_p = err(p32, v32)
    out = {}
    for name, mode in (("fp32", _ffi.NET_FP32), ("split3", _ffi.NET_SPLIT3), ("bf16", _ffi.NET_BF16)):
        net.set_precision(mode)
        p, v = net.forward(states)
        assert np.isfinite(p).all() and np.isfinite(v).all() and np.allclose(p.sum(1), 1.0, atol=1e-4)
        out[name] = err(p, v) + (p,)
    print(f"\n[net F={filters} B={blocks} n={n}] error vs fp64 (value rel, policy/rowmax): "
          f"torch fp32 {ref_v:.2e} {ref_p:.2e} | ours fp32 {out['fp32'][0]:.2e} {out['fp32'][1]:.2e} | "
          f"split3 {out['split3'][0]:.2e} {out['split3'][1]:.2e} | bf16 {out['bf16'][0]:.2e} {out['bf16'][1]:.2e}")
    assert out["fp32"][0] <= max(1e-5, 4 * ref_v) and out["fp32"][1] <= max(1e-5, 4 * ref_p)
    assert out["split3"][0] <= max(1e-5, 4 * ref_v) and out["split3"][1] <= max(1e-5, 4 * ref_p)
    assert (out["split3"][2].argmax(1) == p64.argmax(1)).all()
    assert (out["fp32"][2].argmax(1) == p64.argmax(1)).all()
    net.close()


def test_batch_edges_and_determinism(ctx, oracle):
    from die_e_b200 import _ffi, nnet
    tens = nnet.synthetic_tensors(seed=9, filters=128, blocks=1, bn_stats="random")
    net = _ffi.Net(ctx, tens)
    states, x = _inputs(oracle, 50, seed=3)
    for mode in (_ffi.NET_SPLIT3, _ffi.NET_BF16, _ffi.NET_FP32):
        net.set_precision(mode)
        p_all, v_all = net.forward(states)
        for n in (1, 15, 16, 17, 33):
            p, v = net.forward(states[:n])
            assert (p == p_all[:n]).all() and (v == v_all[:n]).all(), mode   # results do not depend on the batch tiling
        order = np.arange(49, -1, -1)
        p_r, v_r = net.forward(states[order])                        # ... nor on where in the batch a board sits
        assert (p_r == p_all[order]).all() and (v_r == v_all[order]).all(), mode
    net.set_precision(_ffi.NET_SPLIT3)
    p0, v0 = net.forward(states[:0])
    assert p0.shape == (0, 1352)
    bad = states[:2].copy()
    bad["roll"][1] = (0, 0)
    with pytest.raises(_ffi.DieeError):
        net.forward(bad)
    with pytest.raises(_ffi.DieeError):
        _ffi.Net(ctx, tens[:-1])
    net.close()


def test_every_cta_tile_shape_gives_the_same_bits(ctx, oracle, monkeypatch):
    """the tower picks its CTA tile (boards x output channels) from the batch size; every shape accumulates each
    output element over the same K order, so all of them must agree bit for bit (incl. the padded last M-tile of the
    8- and 4-board tiles and a batch that does not fill its last tile)"""
    from die_e_b200 import _ffi, nnet
    tens = nnet.synthetic_tensors(seed=11, filters=256, blocks=2, bn_stats="random")
    net = _ffi.Net(ctx, tens)
    states, _ = _inputs(oracle, 43, seed=5)
    for mode in (_ffi.NET_BF16, _ffi.NET_SPLIT3):      # both tensor-core modes pick their tile the same way
        net.set_precision(mode)
        monkeypatch.setenv("DIEE_CONV_TILE", "16,128")
        p_ref, v_ref = net.forward(states)
        for tile in ("8,128", "8,64", "8,32", "4,32"):
            monkeypatch.setenv("DIEE_CONV_TILE", tile)
            p, v = net.forward(states)
            assert (p == p_ref).all() and (v == v_ref).all(), (mode, tile)
        monkeypatch.delenv("DIEE_CONV_TILE")
        p, v = net.forward(states)   # the automatic choice
        assert (p == p_ref).all() and (v == v_ref).all(), mode
    net.close()


def test_resnet_host_object(ctx, oracle, tmp_path):
    from die_e_b200 import nnet
    net = nnet.ResNet.new(seed=5, filters=128, blocks=1, bn_stats="random", ctx=ctx)
    states, x = _inputs(oracle, 9)
    p, v = net.forward_t(x)                      # the reference's call shape: [N,6,4,6] float tensor
    p2, v2 = net.forward_t(states)
    assert p.shape == (9, 1352) and v.shape == (9, 1) and (p == p2).all() and (v == v2).all()
    assert (net.forward_policy(states) == p).all()
    path = tmp_path / "best_model.ot"
    net.save(path)
    net2 = nnet.ResNet.from_path(path, ctx=ctx)
    p3, v3 = net2.forward_t(states)
    assert (p3 == p).all() and (v3 == v).all()


@pytest.mark.parametrize("n", [256, 300, 1024])
def test_cta_pair_convolution_gives_the_same_bits(ctx, oracle, monkeypatch, n):
    """from 256 boards on the tower's 16-board x 128-channel tile runs on CTA pairs (cta_group::2: one M = 256 MMA over two
    SMs, each CTA staging its own boards and half of the weight tile).  Every output element is accumulated over the same K
    order by the same instruction kind, so the pair form must agree with the single-CTA form bit for bit -- in bf16 and in
    the split-precision mode, for a batch that fills its last pair (256, 1,024) and one that does not (300 = 18.75 tiles)"""
    from die_e_b200 import _ffi, nnet
    tens = nnet.synthetic_tensors(seed=13, filters=256, blocks=2, bn_stats="random")
    net = _ffi.Net(ctx, tens)
    base = positions.midgame_positions(seed=9, n=64, max_adv=100)
    states = np.concatenate([base] * ((n + 63) // 64))[:n]
    states["roll"][:, 0] = 1 + (np.arange(n) % 6)          # make the boards differ
    for mode in (_ffi.NET_BF16, _ffi.NET_SPLIT3):
        net.set_precision(mode)
        monkeypatch.setenv("DIEE_CONV_2CTA", "0")
        monkeypatch.setenv("DIEE_CONV_TILE", "16,128")
        p_ref, v_ref = net.forward(states)
        monkeypatch.setenv("DIEE_CONV_2CTA", "1")
        p, v = net.forward(states)
        assert (p == p_ref).all() and (v == v_ref).all(), mode
        monkeypatch.delenv("DIEE_CONV_2CTA")
        monkeypatch.delenv("DIEE_CONV_TILE")
    net.close()
