"""CPU checks for the AlphaZero path: the host-side Dirichlet sampler of the product equals the oracle's
(bit for bit), and the oracle's search/self-play behave sanely with a synthetic net callback."""
import numpy as np


def test_dirichlet_product_equals_oracle(oracle):
    from die_e_b200 import _ffi
    for seed, epoch, alpha in [(1, 0, 0.3), (0xD1EE, 7, 0.3), (5, 3, 1.0), (9, 1, 2.5), (2, 2, 0.03)]:
        a = _ffi.dirichlet(seed, epoch, alpha)
        b = oracle.dirichlet(seed, epoch, alpha)
        assert a.tobytes() == b.tobytes()
        assert abs(float(a.sum()) - 1.0) < 1e-5 and (a >= 0).all()
    assert _ffi.dirichlet(1, 0, 0.3).tobytes() != _ffi.dirichlet(1, 1, 0.3).tobytes()


def _uniform_eval(states):
    n = len(states)
    p = np.full((n, 1352), 1.0 / 1352, dtype=np.float32)
    v = (states["off"][:, 0].astype(np.float32) - states["off"][:, 1].astype(np.float32)) / 15.0
    return p, v * states["player"].astype(np.float32) * -1.0


def test_oracle_alpha_search_invariants(oracle):
    import positions
    states = positions.midgame_positions(seed=3, n=6, max_adv=60)
    cfg = oracle.mcts_cfg(iterations=30, c=2.0, limit=400)
    cb = oracle.make_eval(_uniform_eval)
    nodes, n_nodes, status = oracle.alpha_mcts_parallel(states, np.arange(6), cfg, 11, 0, cb, 1 + 31 * 40)
    assert (status == 0).all()
    for g in range(6):
        t = nodes[g, :n_nodes[g]]
        root = t[0]
        nc = int(root["n_children"])
        assert nc == len(oracle.bg_valid_moves(states[g:g + 1]))
        ch = t[root["first_child"]:root["first_child"] + nc]
        if nc:
            assert abs(float(ch["prior"].sum()) - 1.0) < 1e-4
            # every iteration passes through the root: visits = 1 + iterations (+ stale hits on game 0, Q9)
            assert root["visits"] >= 31
            assert ch["visits"].sum() == root["visits"] - 1 or g == 0


def test_oracle_self_play_records(oracle):
    cfg = oracle.mcts_cfg(iterations=4, c=2.0, limit=60)
    cb = oracle.make_eval(_uniform_eval)
    rec, pi_ids, pi_vals, waves = oracle.self_play(3, cfg, 1.25, 5, 100, cb, 1 + 5 * 40)
    assert len(rec) > 0 and waves >= 60 or waves > 0
    assert set(np.unique(rec["outcome"])) <= {-1, 0, 1}
    assert (rec["game_id"] >= 100).all() and (rec["game_id"] < 103).all()
    for r in rec[:50]:
        ids = pi_ids[r["pi_offset"]:r["pi_offset"] + r["n_pi"]]
        want = [oracle.bg_encode(np.array([r["state"]]), m) for m in oracle.bg_valid_moves(np.array([r["state"]]))]
        assert list(ids) == want


def test_training_data_files_round_trip(tmp_path):
    """ps.ot / states.ot / outcomes.ot as the tch training loop reads them (alphazero.rs:149-200)"""
    import numpy as np
    import pytest
    import torch
    from die_e_b200 import alphazero
    rng = np.random.default_rng(3)
    data = []
    for i in range(7):
        ps = np.zeros(1352, dtype=np.float32)
        ps[rng.integers(0, 1352, 5)] = rng.random(5, dtype=np.float32)
        data.append(alphazero.MemoryFragment(int(rng.integers(-1, 2)), ps, rng.integers(-3, 4, (1, 6, 4, 6)).astype(np.float32)))
    d = tmp_path / alphazero.sp_dir("backgammon", "abc", 2, 1, base="")
    with pytest.raises(FileNotFoundError):
        alphazero.save_training_data(data, d)
    d.mkdir(parents=True)
    assert str(d).endswith("data/backgammon/run-abc/lrn-2/sp-1")
    alphazero.save_training_data(data, d)
    # each file is a libtorch archive holding one tensor under the key "0" (what tch's Tensor::load expects)
    for name, shape, dtype in (("ps.ot", (7, 1352), torch.float32), ("states.ot", (7, 6, 4, 6), torch.float32), ("outcomes.ot", (7,), torch.int8)):
        mod = torch.jit.load(str(d / name))
        named = dict(list(mod.named_parameters()) + list(mod.named_buffers()))
        assert list(named) == ["0"] and tuple(named["0"].shape) == shape and named["0"].dtype == dtype
    back = alphazero.load_training_data(d)
    assert len(back) == 7
    for a, b in zip(data, back):
        assert a.outcome == b.outcome and (a.ps == b.ps).all() and (a.state == b.state).all() and b.state.shape == (1, 6, 4, 6)


def test_reference_mcts_test_invariants(oracle):
    """tests/mcts_test.rs of the reference (its only two tests on the search path), on the oracle twins:
    :17-33  turn_policy_to_probs_tensor_parallel -- with a RANDOM policy every row of masked, renormalised priors sums
            to 1 (here: the priors of every root's children, one root per game, and of every expanded node);
    :41-59  get_prob_tensor_parallel -- a root whose children all have 10 visits gives a row that sums to 1."""
    import positions
    rng = np.random.default_rng(7)

    def random_eval(states):
        n = len(states)
        return rng.random((n, 1352), dtype=np.float32), np.zeros(n, dtype=np.float32)
    states = positions.midgame_positions(seed=9, n=10, max_adv=60)
    cfg = oracle.mcts_cfg(iterations=12, c=2.0, limit=400, alpha=0.3, eps=0.0)
    nodes, n_nodes, status = oracle.alpha_mcts_parallel(states, np.arange(10), cfg, 3, 0, oracle.make_eval(random_eval), 1 + 13 * 130)
    assert (status == 0).all()
    rows = 0
    for g in range(10):
        t = nodes[g, :n_nodes[g]]
        for nd in t:
            nc = int(nd["n_children"])
            if nc:
                ch = t[nd["first_child"]:nd["first_child"] + nc]
                assert abs(float(ch["prior"].astype(np.float64).sum()) - 1.0) <= 1e-5
                rows += 1
    assert rows >= 10
    # get_prob_tensor_parallel: children with 10 visits each
    t = nodes[0, :n_nodes[0]].copy()
    nc = int(t[0]["n_children"])
    assert nc > 0
    t["visits"][t[0]["first_child"]:t[0]["first_child"] + nc] = 10.0
    ids, pi = oracle.root_pi(t, 1.0)
    assert len(ids) == nc and abs(float(pi.astype(np.float64).sum()) - 1.0) <= 1e-5
    assert np.allclose(pi, 1.0 / nc, rtol=1e-6)
