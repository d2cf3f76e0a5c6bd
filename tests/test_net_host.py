"""CPU checks of the net's host side: tensor order/shapes, .ot round trip with tch's collision
naming, as_tensor inversion, and the oracle's own consistency (fp64 vs fp32, folded vs unfolded)."""
import numpy as np
import pytest

import net_oracle


def test_param_counts_match_survey():
    from die_e_b200 import nnet
    shapes = nnet.tensor_shapes(256, 19)
    assert len(shapes) == 250                              # 250 named tensors
    tens = nnet.synthetic_tensors(filters=256, blocks=1)
    assert len(tens) == 22 + 12
    total = sum(int(np.prod(s)) for s in shapes)
    bn_buffers = 2 * (256 * (1 + 2 * 19) + 32 + 3)
    assert bn_buffers == 20038 and total - bn_buffers == 23577594   # SURVEY N1 figures


def test_ot_round_trip(tmp_path):
    from die_e_b200 import nnet
    tens = nnet.synthetic_tensors(seed=3, filters=128, blocks=2, bn_stats="random")
    p = str(tmp_path / "model_0.ot")
    nnet.save_ot(p, tens, 128, 2)
    back, filters, blocks = nnet.load_ot(p)
    assert (filters, blocks) == (128, 2) and len(back) == len(tens)
    for a, b in zip(tens, back):
        assert a.shape == b.shape and (a == b).all()
    names = nnet._tch_names(128, 2)
    assert names[0] == "weight" and names[1] == "bias" and len(set(names)) == len(names)
    assert sum(n.startswith("running_mean") for n in names) == 1 + 2 * 2 + 2


def test_states_from_tensor_inverts_as_tensor(oracle):
    import positions
    from die_e_b200 import nnet
    states = positions.reachable_positions(seed=4, n_games=3)
    x = np.concatenate([oracle.bg_as_tensor(states[i:i + 1]) for i in range(len(states))])
    back = nnet.states_from_tensor(x)
    assert back.tobytes() == states.tobytes()


def test_oracle_self_consistency(oracle):
    import positions
    from die_e_b200 import nnet
    tens = nnet.synthetic_tensors(seed=1, filters=128, blocks=2, bn_stats="random")
    states = positions.midgame_positions(seed=1, n=6)
    x = np.concatenate([oracle.bg_as_tensor(states[i:i + 1]) for i in range(len(states))])
    import torch
    p64, v64 = net_oracle.forward(tens, x, 2)
    p32, v32 = net_oracle.forward(tens, x, 2, dtype=torch.float32)
    assert np.allclose(p64.sum(1), 1.0) and (np.abs(v64) <= 1).all()
    assert np.abs(p64 - p32).max() < 1e-5 and np.abs(v64 - v32).max() < 1e-5
    pb, vb = net_oracle.forward_bf16_emulated(tens, x, 2)
    assert np.abs(pb - p64).max() < 5e-2 and np.abs(vb - v64).max() < 1e-1
