"""GPU parity tests of the AlphaZero search and the self-play driver against the CPU oracle.

The oracle's net is a callback that evaluates each batch with the PRODUCT's net on the GPU, so both
sides see identical net outputs ("visit counts given identical net outputs"); the net
forward is batch-tiling invariant (tests/test_gpu_net.py), so a state gets the same outputs whatever
batch it is in.  Bar: bit-exact -- every node's parent / child range / visits / value / prior /
action / state, the per-game status, and every self-play record (state, outcome, ply, pi ids, pi values)."""
import numpy as np
import pytest

import positions

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from die_e_b200 import _ffi
    return _ffi.Context(0)


@pytest.fixture(scope="module")
def net(ctx):
    from die_e_b200 import _ffi, nnet
    return _ffi.Net(ctx, nnet.synthetic_tensors(seed=21, filters=128, blocks=1, bn_stats="random"))


def _eval_cb(oracle, net):
    return oracle.make_eval(lambda st: net.forward(st))


def _cmp_pools(oracle, got_nodes, got_n, want_nodes, want_n):
    assert (got_n == want_n).all()
    for g in range(len(got_n)):
        k = int(want_n[g])
        a, b = got_nodes[g, :k], want_nodes[g, :k]
        for f in ("parent", "first_child", "n_children"):
            assert (a[f] == b[f]).all(), (g, f)
        for f in ("visits", "value", "prior", "action", "state"):
            assert a[f].tobytes() == b[f].tobytes(), (g, f)


@pytest.mark.parametrize("n,iters,seed", [(16, 25, 3), (40, 12, 9)])
def test_alpha_search_bit_exact(ctx, net, oracle, n, iters, seed):
    states = positions.midgame_positions(seed=seed, n=n, max_adv=110)
    ids = np.arange(100, 100 + n, dtype=np.uint32)
    cfg = oracle.mcts_cfg(iterations=iters, c=2.0, limit=400, alpha=0.3, eps=0.25)
    max_nodes = 1 + (iters + 1) * 40
    r_ids, r_moves, r_vis, r_cnt, status, nodes, n_nodes = ctx.alpha_search(net, states, ids, cfg, seed, epoch=4,
                                                                            max_nodes=max_nodes, dump=True)
    o_nodes, o_n, o_status = oracle.alpha_mcts_parallel(states, ids, cfg, seed, 4, _eval_cb(oracle, net), max_nodes)
    assert (status == o_status).all() and (status == 0).all()
    _cmp_pools(oracle, nodes, n_nodes, o_nodes, o_n)
    for g in range(n):
        root = o_nodes[g, 0]
        nc = int(root["n_children"])
        assert r_cnt[g] == nc
        ch = o_nodes[g, root["first_child"]:root["first_child"] + nc]
        assert r_vis[g, :nc].tobytes() == ch["visits"].tobytes()
        assert r_moves[g, :nc].tobytes() == ch["action"].tobytes()
        assert [int(x) for x in r_ids[g, :nc]] == [oracle.bg_encode(states[g:g + 1], m) for m in oracle.moves_to_list(ch["action"], nc)]


@pytest.mark.parametrize("prec", ["fp32", "split3"])
def test_alpha_search_end_to_end_against_an_independent_fp32_net(ctx, oracle, prec):
    """End to end, nothing shared: the GPU search with the net in a mode inside the fp32 tolerance (DIEE_NET_FP32 on the
    CUDA cores, DIEE_NET_SPLIT3 -- the default -- on the tensor cores) against the
    oracle search whose net is torch's fp32 CPU forward of the same weights (what the reference's tch computes).
    The two nets agree to ~1e-6, so the trees can only differ where two PUCT scores are closer than that.
    Bar: root visit counts identical for >= 90 % of the games; where a near-tie went the other way the total-variation
    distance of the root visit distributions stays <= 0.1 (one visit of a 30-iteration search is 1/30, and the torch CPU
    forward itself moves in its last bits with the thread count); value estimates of the root children
    (value / visits) within 1e-5 relative to max(|q|, 0.1) wherever the visit counts agree -- the metric
    tests/test_gpu_net.py uses for the value head itself: q is a mean of tanh outputs, each a 72-term sum that cancels,
    and two correct fp32 forwards already differ by ~2e-7 ABSOLUTE on it, so a purely relative bar on a q near zero
    would test the rounding of the reference's own arithmetic, not this engine."""
    import torch
    import net_oracle
    from die_e_b200 import _ffi, nnet
    blocks = 2
    tens = nnet.synthetic_tensors(seed=33, filters=128, blocks=blocks, bn_stats="random")
    gnet = _ffi.Net(ctx, tens)
    gnet.set_precision(_ffi.NET_FP32 if prec == "fp32" else _ffi.NET_SPLIT3)
    n, iters, seed = 24, 30, 5
    states = positions.midgame_positions(seed=seed, n=n, max_adv=100)
    ids = np.arange(500, 500 + n, dtype=np.uint32)
    cfg = oracle.mcts_cfg(iterations=iters, c=2.0, limit=400, alpha=0.3, eps=0.25)
    max_nodes = 1 + (iters + 1) * 40

    def torch_fp32(st):
        x = np.concatenate([oracle.bg_as_tensor(st[i:i + 1]) for i in range(len(st))])
        p, v = net_oracle.forward(tens, x, blocks, dtype=torch.float32)
        return p.astype(np.float32), v.astype(np.float32)
    o_nodes, o_n, o_status = oracle.alpha_mcts_parallel(states, ids, cfg, seed, 2, oracle.make_eval(torch_fp32), max_nodes)
    r_ids, r_moves, r_vis, r_cnt, status, nodes, n_nodes = ctx.alpha_search(gnet, states, ids, cfg, seed, epoch=2,
                                                                            max_nodes=max_nodes, dump=True)
    assert (status == o_status).all()
    same, worst_tv, worst_q = 0, 0.0, 0.0
    for g in range(n):
        root = o_nodes[g, 0]
        nc = int(root["n_children"])
        assert r_cnt[g] == nc
        ch = o_nodes[g, root["first_child"]:root["first_child"] + nc]
        assert r_moves[g, :nc].tobytes() == ch["action"].tobytes()       # legal moves and their order never depend on the net
        a, b = r_vis[g, :nc].astype(np.float64), ch["visits"].astype(np.float64)
        if nc:
            worst_tv = max(worst_tv, 0.5 * np.abs(a / max(a.sum(), 1) - b / max(b.sum(), 1)).sum())
        if (a == b).all():
            same += 1
            gch = nodes[g, nodes[g, 0]["first_child"]:nodes[g, 0]["first_child"] + nc]
            vis = b > 0
            qa = gch["value"][vis].astype(np.float64) / b[vis]
            qb = ch["value"][vis].astype(np.float64) / b[vis]
            if vis.any():
                worst_q = max(worst_q, (np.abs(qa - qb) / np.maximum(np.abs(qb), 1e-1)).max())
    print(f"\n[alpha end-to-end {prec}] identical root visit counts {same}/{n}, worst TV {worst_tv:.2e}, worst value rel {worst_q:.2e}")
    assert same >= 0.9 * n and worst_tv <= 0.1 and worst_q <= 1e-5, (same, worst_tv, worst_q)
    gnet.close()


def test_alpha_search_endgame_quirks(ctx, net, oracle):
    """positions one move from the end: terminal leaves (value +-1 w.r.t. the root player), stale slots that
    hit game 0's root (Q9), a root with no legal move (Q12) and an already finished game"""
    def pts(d):
        p = [0] * 24
        for k, v in d.items():
            p[k] = v
        return p
    states = np.concatenate([
        oracle.make_state(pts({3: -2, 20: 3}), off=(13, 12), roll=(5, 3), player=-1),
        oracle.make_state(pts({0: -1, 23: 1}), off=(14, 14), roll=(3, 4), player=-1),
        oracle.make_state(pts({0: -1, 23: 1}), off=(14, 14), roll=(6, 6), player=1),
        oracle.make_state(pts({20: -1, 19: 2, 18: 2, 2: 1}), off=(14, 10), roll=(1, 2), player=-1),
        oracle.make_state(pts({5: -2, 4: -3, 18: 2, 19: 3}), off=(10, 10), roll=(2, 1), player=1),
        oracle.make_state(pts({1: -1}), off=(14, 15), roll=(2, 1), player=-1),
    ])
    ids = np.arange(6, dtype=np.uint32)
    cfg = oracle.mcts_cfg(iterations=20, c=2.0, limit=400)
    r = ctx.alpha_search(net, states, ids, cfg, 77, epoch=0, max_nodes=600, dump=True)
    o_nodes, o_n, o_status = oracle.alpha_mcts_parallel(states, ids, cfg, 77, 0, _eval_cb(oracle, net), 600)
    assert (r[4] == o_status).all()
    _cmp_pools(oracle, r[5], r[6], o_nodes, o_n)
    # pool exhaustion is reported per game, identically
    r2 = ctx.alpha_search(net, states[:2], ids[:2], cfg, 77, epoch=0, max_nodes=8, dump=True)
    o2 = oracle.alpha_mcts_parallel(states[:2], ids[:2], cfg, 77, 0, _eval_cb(oracle, net), 8)
    assert (r2[4] == o2[2]).all()


@pytest.mark.parametrize("n_games,iters,limit", [(12, 6, 400), (10, 4, 40)])
def test_self_play_bit_exact(ctx, net, oracle, n_games, iters, limit):
    cfg = oracle.mcts_cfg(iterations=iters, c=2.0, limit=limit, alpha=0.3, eps=0.25)
    max_nodes = 1 + (iters + 1) * 128
    rec, pi_ids, pi_vals, waves = ctx.selfplay_run(net, n_games, cfg, 1.25, seed=31, first_game_id=1000, max_nodes=max_nodes)
    o_rec, o_ids, o_vals, o_waves = oracle.self_play(n_games, cfg, 1.25, 31, 1000, _eval_cb(oracle, net), max_nodes)
    assert waves == o_waves and len(rec) == len(o_rec) and len(rec) > 0
    assert rec.tobytes() == o_rec.tobytes()
    assert pi_ids.tobytes() == o_ids.tobytes() and pi_vals.tobytes() == o_vals.tobytes()
    if limit == 400:
        assert set(np.unique(rec["outcome"])) == {-1, 1}          # every game was played to a winner
        assert len(np.unique(rec["game_id"])) == n_games
    else:
        assert (rec["outcome"] == 0).any()                         # round-capped games (Q10)


# ---------------------------------------------------------------- NON-PARITY throughput modes
def test_virtual_loss_search_with_one_leaf_is_the_reference_search(ctx, net, oracle):
    """diee_alpha_search_vl with one leaf per step and no virtual loss runs the virtual-loss kernels but must give the
    reference's search exactly, as long as no terminal leaf / no-move leaf is in reach (where quirks Q9 / Q12 differ)"""
    states = positions.midgame_positions(seed=17, n=32, max_adv=30)     # early positions: 20 iterations cannot reach the end
    ids = np.arange(700, 732, dtype=np.uint32)
    cfg = oracle.mcts_cfg(iterations=20, c=2.0, limit=400, alpha=0.3, eps=0.25)
    a = ctx.alpha_search(net, states, ids, cfg, 11, epoch=2)
    b = ctx.alpha_search_vl(net, states, ids, cfg, 11, epoch=2, leaves_per_game=1, virtual_loss=0.0)
    for x, y in zip(a, b):
        assert np.asarray(x).tobytes() == np.asarray(y).tobytes()


@pytest.mark.parametrize("K,vl", [(4, 1.0), (8, 0.5)])
def test_virtual_loss_search_properties(ctx, net, oracle, K, vl):
    """K leaves per game and step: same legal moves at the root, every simulation accounted for (root children visits
    sum to the number of evaluated or terminal leaves <= iterations), deterministic, and close to the one-leaf search
    (total-variation distance of the root visit distributions)"""
    n, iters = 48, 64
    states = positions.midgame_positions(seed=19, n=n, max_adv=100)
    ids = np.arange(n, dtype=np.uint32)
    cfg = oracle.mcts_cfg(iterations=iters, c=2.0, limit=400, alpha=0.3, eps=0.25)
    ref = ctx.alpha_search(net, states, ids, cfg, 5, epoch=1)
    got = ctx.alpha_search_vl(net, states, ids, cfg, 5, epoch=1, leaves_per_game=K, virtual_loss=vl)
    again = ctx.alpha_search_vl(net, states, ids, cfg, 5, epoch=1, leaves_per_game=K, virtual_loss=vl)
    for x, y in zip(got, again):
        assert np.asarray(x).tobytes() == np.asarray(y).tobytes()
    assert (got[4] == 0).all() and (got[3] == ref[3]).all()
    tv = []
    for g in range(n):
        nc = int(ref[3][g])
        assert got[1][g, :nc].tobytes() == ref[1][g, :nc].tobytes() and (got[0][g, :nc] == ref[0][g, :nc]).all()
        s = float(got[2][g, :nc].sum())
        assert s <= iters + 1e-3 and (nc == 0 or s >= iters * 0.5), (g, s)   # duplicates of a pending leaf are skipped, not re-evaluated
        if nc:
            a, b = got[2][g, :nc] / max(s, 1), ref[2][g, :nc] / max(float(ref[2][g, :nc].sum()), 1)
            tv.append(0.5 * float(np.abs(a - b).sum()))
    print(f"\n[virtual loss K={K} vl={vl}] mean TV distance to the one-leaf search {np.mean(tv):.3f}, max {np.max(tv):.3f}")
    assert np.mean(tv) < 0.25


def test_refilled_self_play(ctx, net, oracle):
    """DIEE_SP_REFILL: a finished game's slot starts a new game (new id) so the forward batch stays full; the FIRST game of
    every slot is the reference's game record for record; later games are complete, labelled games of fresh ids"""
    n_games, iters, target = 12, 4, 30
    cfg = oracle.mcts_cfg(iterations=iters, c=2.0, limit=400, alpha=0.3, eps=0.25)
    base, b_ids, b_vals, _ = ctx.selfplay_run(net, n_games, cfg, 1.25, seed=31, first_game_id=1000)
    from die_e_b200 import _ffi
    rec, pi_ids, pi_vals, rep = ctx.selfplay_run_ex(net, n_games, cfg, 1.25, 31, 1000, flags=_ffi.SP_REFILL, target_games=target)
    assert rep["games_finished"] >= target and rep["games_finished"] < target + n_games
    assert rep["game_moves"] == rep["waves"] * n_games or rep["games_cut"] < n_games        # every slot busy in every wave
    ids = np.unique(rec["game_id"])
    assert ids.min() == 1000 and ids.max() >= 1000 + target - 1
    # a refilled batch couples games differently from wave `k` on (the Dirichlet epoch is the wave number and slot 0's root
    # feeds quirk Q9), so only games that ran entirely while no slot had been refilled yet are the reference's own
    first_finish = min(int(base[base["game_id"] == g]["ply"].max()) for g in range(1000, 1000 + n_games))
    for g in range(1000, 1000 + n_games):
        a, b = base[base["game_id"] == g], rec[rec["game_id"] == g]
        k = int((a["ply"] <= first_finish).sum())
        assert a[:k]["state"].tobytes() == b[:k]["state"].tobytes(), g
    done = [g for g in ids if (rec[rec["game_id"] == g]["outcome"] != 0).any()]
    assert len(done) >= target - 2                                     # (a round-capped game is labelled 0)
    for g in done:
        r = rec[rec["game_id"] == g]
        assert set(np.unique(r["outcome"])) <= {-1, 1} and (np.diff(r["ply"].astype(int)) > 0).all()
