"""Edge cases of the round-2 entry points through the C ABI: empty batches, bad arguments, capacity errors, and the
BASELINE configs[1] size (65,536 playouts) sampled against the oracle."""
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


def test_c2_playouts_at_65536_games_sampled_against_the_oracle(ctx, oracle):
    import bench
    from die_e_b200 import _ffi
    n = 65536
    starts = bench.initial_states(_ffi, 0, n)
    winners, plies, finals = ctx.bg_playout(starts, seed=0xD1EE, first_game_id=0, round_limit=400, want_finals=True)
    for g in np.linspace(0, n - 1, 96).astype(int):
        w, p, f = oracle.bg_playout(starts[g:g + 1], 0xD1EE, int(g), 400)
        assert (int(winners[g]), int(plies[g])) == (w, p) and finals[g:g + 1].tobytes() == f.tobytes(), g
    assert set(np.unique(winners)) <= {-1, 0, 1} and (winners != 0).mean() > 0.999


def test_empty_batches(ctx, net):
    from die_e_b200 import _ffi
    cfg = np.zeros(1, dtype=_ffi.MCTS_CFG)
    cfg[0] = (8, 2.0, 40, 0.3, 0.25, _ffi.MODE_PASS_CHILD)
    none = np.zeros(0, dtype=_ffi.BG_STATE)
    best, status, stats = ctx.mcts_search(_ffi.GAME_BACKGAMMON, none, np.zeros(0, np.int8), cfg, 1)
    assert len(best) == 0 and len(status) == 0
    r = ctx.alpha_search(net, none, np.zeros(0, np.uint32), cfg, 1)
    assert all(len(x) == 0 for x in r)
    r = ctx.alpha_search_vl(net, none, np.zeros(0, np.uint32), cfg, 1)
    assert all(len(x) == 0 for x in r)
    rec, ids, vals, rep = ctx.selfplay_run_ex(net, 0, cfg, 1.25, 1, 0, rec_cap=4, pi_cap=4)
    assert len(rec) == 0 and rep["waves"] == 0
    w, p = ctx.bg_playout(none, seed=1)
    assert len(w) == 0
    p_, v_ = net.forward(none)
    assert p_.shape == (0, 1352)


def test_bad_arguments_are_errors_not_crashes(ctx, net):
    from die_e_b200 import _ffi
    cfg = np.zeros(1, dtype=_ffi.MCTS_CFG)
    cfg[0] = (8, 2.0, 40, 0.3, 0.25, 0)
    states = positions.midgame_positions(seed=3, n=4, max_adv=30)
    ids = np.arange(4, dtype=np.uint32)
    for kw in (dict(leaves_per_game=0), dict(leaves_per_game=65), dict(virtual_loss=-1.0)):
        with pytest.raises(_ffi.DieeError) as e:
            ctx.alpha_search_vl(net, states, ids, cfg, 1, **kw)
        assert e.value.code == _ffi.ERR_INVALID
    # a refilled run without a stopping rule, an unknown flag, a target without refill, a negative time box
    for kw in (dict(flags=_ffi.SP_REFILL), dict(flags=8), dict(target_games=5), dict(max_waves=-1), dict(leaves_per_game=99)):
        with pytest.raises(_ffi.DieeError) as e:
            ctx.selfplay_run_ex(net, 4, cfg, 1.25, 1, 0, rec_cap=64, pi_cap=4096, **kw)
        assert e.value.code == _ffi.ERR_INVALID
    # record buffers too small: an overflow, reported (not a write past the end)
    with pytest.raises(_ffi.DieeError) as e:
        ctx.selfplay_run_ex(net, 4, cfg, 1.25, 1, 0, rec_cap=2, pi_cap=4096, max_waves=3)
    assert e.value.code == _ffi.ERR_OVERFLOW
    unrolled = states.copy()
    unrolled["roll"][2] = (0, 0)
    with pytest.raises(_ffi.DieeError) as e:
        ctx.alpha_search_vl(net, unrolled, ids, cfg, 1)
    assert e.value.code == _ffi.ERR_NOT_ROLLED
    with pytest.raises(_ffi.DieeError):
        _ffi.Arena(ctx, 0, 1, 400)
    arena = _ffi.Arena(ctx, 8, 1, 400)
    with pytest.raises(_ffi.DieeError) as e:
        arena.round(7, _ffi.AGENT_RANDOM)
    assert e.value.code == _ffi.ERR_INVALID
    with pytest.raises(_ffi.DieeError):
        arena.round(_ffi.AGENT_MCTS, _ffi.AGENT_RANDOM, None)     # Agent::Mcts needs a config
    assert arena.round(_ffi.AGENT_RANDOM, _ffi.AGENT_RANDOM)[0] >= 0   # still usable
    arena.close()


def test_arena_round_limit_and_tiny_arenas(ctx, oracle):
    import orc_arena
    from die_e_b200.versus import Agent, Player, play
    for n_games, limit, seed in ((1, 400, 3), (2, 400, 4), (5, 7, 5), (33, 25, 6)):
        res = play("backgammon", Player(Agent.Random), Player(Agent.Random), seed=seed, num_games=n_games, round_limit=limit, ctx=ctx)
        w1, w2, winners, rounds = orc_arena.play_backgammon("random", "random", None, 1.0, seed, n_games, limit)
        assert (res.winners == winners).all() and (res.rounds == rounds).all() and (res.wins_p1, res.wins_p2) == (w1, w2)
