"""Parity at the sizes BASELINE.json names -- what bench.py times is what is checked here.

  C3  pure MCTS, 1,024 games x iterations=100, c=2, limit=400 (configs[2]): one launch over the whole batch, 32 sampled
      games compared with the oracle node by node (visits / values bit-exact), every rollout's end state included.
  C4  alpha_mcts_parallel, 1,024 games x iterations=100, ResNet 256x19 (configs[3]): the WHOLE lock-step search against
      the oracle (the games of a batch are coupled through slot 0 -- quirk Q9 -- so no subset can be checked alone);
      the oracle's net is a callback into the product's GPU net, so both sides see identical net outputs.
  P5  self_play_parallel with 256 games, every record compared.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

SEED = 0xD1EE


@pytest.fixture(scope="module")
def ctx():
    from die_e_b200 import _ffi
    return _ffi.Context(0)


def _bench_inputs(ctx, n, first_gid=0):
    import bench
    from die_e_b200 import _ffi
    return bench.midgame_states(ctx, _ffi, first_gid, n)


def test_bench_inputs_equal_the_reference_arms(ctx, oracle):
    """both arms of bench.py run on the same positions: the GPU-made batch == the oracle-made one"""
    import argparse
    import bench
    got = _bench_inputs(ctx, 1024)
    want = bench.host_states_for_reference(argparse.Namespace(games=1024, workload="mcts"), 0, 1024)
    assert got.tobytes() == want.tobytes()
    got7 = _bench_inputs(ctx, 64, first_gid=7 * 1024)           # another rank's shard: keyed by GLOBAL game id
    want7 = bench.host_states_for_reference(argparse.Namespace(games=64, workload="mcts"), 7 * 1024, 64)
    assert got7.tobytes() == want7.tobytes()


@pytest.mark.parametrize("mode", [2, 3])
def test_c3_pure_mcts_at_1024_games(ctx, oracle, mode):
    from die_e_b200 import _ffi as ffi
    n = 1024
    states = _bench_inputs(ctx, n)
    players = states["player"].copy()
    cfg = oracle.mcts_cfg(iterations=100, c=2.0, limit=400, mode=mode)
    best, status, stats, nodes, nstates, n_nodes, finals = ctx.mcts_search(ffi.GAME_BACKGAMMON, states, players, cfg, SEED, 0, 7,
                                                                          dump=True)
    assert (status == 0).all()
    # the device-resident call bench.py times gives the same best moves
    best2, status2, _ = ctx.mcts_search(ffi.GAME_BACKGAMMON, states, players, cfg, SEED, 0, 7)
    assert best2.tobytes() == best.tobytes() and (status2 == 0).all()
    sample = np.linspace(0, n - 1, 32).astype(int)
    rollouts = 0
    for i in sample:
        rc, obest, onodes, ostates, ofin = oracle.mcts_search_bg(states[i:i + 1], int(players[i]), cfg, SEED, int(i), 7, want_finals=True)
        k = len(onodes)
        assert rc == 0 and n_nodes[i] == k, i
        got = nodes[i, :k]
        for f in ("parent", "n_moves", "n_untried"):
            assert (got[f] == onodes[f]).all(), (i, f)
        assert got["visits"].tobytes() == onodes["visits"].tobytes(), i
        assert got["value"].tobytes() == onodes["value"].tobytes(), i
        assert got["action"].tobytes() == onodes["action"].tobytes(), i
        assert nstates[i, :k].tobytes() == ostates.tobytes(), i
        assert best[i:i + 1].tobytes() == obest.tobytes(), i
        assert finals[i].tobytes() == ofin.tobytes(), i
        rollouts += int(np.frombuffer(ofin.tobytes(), dtype=np.uint8).reshape(-1, 32).any(axis=1).sum())
    assert rollouts > 2000
    if mode == 2:
        assert int(stats["rollout_plies"].sum()) % 400 == 0     # Q5: a reference-exact rollout always runs to the cap
        assert 0 < ctx.search_work() < int(stats["rollout_plies"].sum())


def test_c4_alpha_search_at_1024_games_100_iterations_256x19(ctx, oracle):
    from die_e_b200 import _ffi, nnet
    net = _ffi.Net(ctx, nnet.synthetic_tensors(seed=SEED, filters=256, blocks=19, bn_stats="identity"))
    n, iters = 1024, 100
    states = _bench_inputs(ctx, n)
    ids = np.arange(n, dtype=np.uint32)
    cfg = oracle.mcts_cfg(iterations=iters, c=2.0, limit=400, alpha=0.3, eps=0.25)
    max_nodes = 1 + (iters + 1) * 64      # smaller slabs than the default keep the two pool dumps at ~400 MB each; an
    #                                       exhausted slab is reported per game, identically on both sides
    r_ids, r_moves, r_vis, r_cnt, status, nodes, n_nodes = ctx.alpha_search(net, states, ids, cfg, SEED, epoch=3,
                                                                            max_nodes=max_nodes, dump=True)
    evals = [0]

    def ev(st):
        evals[0] += len(st)
        return net.forward(st)
    o_nodes, o_n, o_status = oracle.alpha_mcts_parallel(states, ids, cfg, SEED, 3, oracle.make_eval(ev), max_nodes)
    assert evals[0] == n * (iters + 1)
    assert (status == o_status).all()
    assert (status == 0).sum() >= 0.9 * n
    assert (n_nodes == o_n).all()
    for g in range(n):
        if status[g] != 0:
            continue
        k = int(o_n[g])
        a, b = nodes[g, :k], o_nodes[g, :k]
        for f in ("parent", "first_child", "n_children"):
            assert (a[f] == b[f]).all(), (g, f)
        for f in ("visits", "value", "prior", "action", "state"):
            assert a[f].tobytes() == b[f].tobytes(), (g, f)
        root = b[0]
        nc = int(root["n_children"])
        assert r_cnt[g] == nc
        ch = b[root["first_child"]:root["first_child"] + nc]
        assert r_vis[g, :nc].tobytes() == ch["visits"].tobytes()
        assert r_moves[g, :nc].tobytes() == ch["action"].tobytes()
    net.close()


def test_self_play_at_256_games(ctx, oracle):
    from die_e_b200 import _ffi, nnet
    net = _ffi.Net(ctx, nnet.synthetic_tensors(seed=21, filters=128, blocks=1, bn_stats="random"))
    n_games, iters = 256, 4
    cfg = oracle.mcts_cfg(iterations=iters, c=2.0, limit=400, alpha=0.3, eps=0.25)
    max_nodes = 1 + (iters + 1) * 128
    rec, pi_ids, pi_vals, waves = ctx.selfplay_run(net, n_games, cfg, 1.25, seed=SEED, first_game_id=4096, max_nodes=max_nodes)
    o_rec, o_ids, o_vals, o_waves = oracle.self_play(n_games, cfg, 1.25, SEED, 4096, oracle.make_eval(lambda st: net.forward(st)),
                                                     max_nodes)
    assert waves == o_waves and len(rec) == len(o_rec) and len(rec) > 20 * n_games
    assert rec.tobytes() == o_rec.tobytes()
    assert pi_ids.tobytes() == o_ids.tobytes() and pi_vals.tobytes() == o_vals.tobytes()
    assert len(np.unique(rec["game_id"])) == n_games
    # the time-boxed form is a prefix of the same run: the first `w` waves' searches, then every game hands in its records
    rec_w, ids_w, vals_w, rep = ctx.selfplay_run_ex(net, n_games, cfg, 1.25, SEED, 4096, max_nodes=max_nodes, max_waves=5)
    assert rep["waves"] == 5 and rep["games_cut"] == n_games and rep["game_moves"] == 5 * n_games
    assert (rec_w["outcome"] == 0).all() and len(rec_w) <= 5 * n_games
    first5 = {(int(r["game_id"]), int(r["ply"])): r["state"].tobytes() for r in rec if r["ply"] < 5}
    for r in rec_w:
        assert first5[(int(r["game_id"]), int(r["ply"]))] == r["state"].tobytes()
    net.close()


def test_alpha_search_rejects_bad_dirichlet_parameters(ctx):
    """Dirichlet::new(..).unwrap() panics for alpha <= 0 (noise.rs:29); here an error -- never a hang in the Gamma
    rejection loop (alpha < -2/3) or silent NaN priors (alpha == 0)"""
    from die_e_b200 import _ffi, nnet
    import positions
    net = _ffi.Net(ctx, nnet.synthetic_tensors(seed=21, filters=128, blocks=1, bn_stats="random"))
    states = positions.midgame_positions(seed=1, n=4, max_adv=20)
    ids = np.arange(4, dtype=np.uint32)
    for alpha, eps in ((0.0, 0.25), (-1.0, 0.25), (float("nan"), 0.25), (0.3, -0.1), (0.3, 1.5)):
        cfg = np.zeros(1, dtype=_ffi.MCTS_CFG)
        cfg[0] = (4, 2.0, 400, alpha, eps, 0)
        with pytest.raises(_ffi.DieeError) as e:
            ctx.alpha_search(net, states, ids, cfg, 1)
        assert e.value.code == _ffi.ERR_INVALID
        with pytest.raises(_ffi.DieeError):
            ctx.selfplay_run(net, 4, cfg, 1.25, seed=1)
    net.close()
