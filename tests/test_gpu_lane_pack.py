"""The packed lane kernel (lane_pack_kernel: a CTA re-packs its games by sub-case every ply; lane_kernels.cu) against the
lane-resident kernel and the oracle.  A game's dice and choices are keyed by (game id, ply), so which thread plays a ply
must change nothing: winners, ply counts, end states and whole search dumps are compared bit for bit.
DIEE_LANE_PACK=2 forces the packed kernel for every job size, =0 keeps the lane-resident one."""
import numpy as np
import pytest

import positions

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from die_e_b200 import _ffi
    return _ffi.Context(0)


def _opening(n, seed):
    from die_e_b200 import _ffi
    starts = np.zeros(n, dtype=_ffi.BG_STATE)
    starts["pts"][:] = [2, 0, 0, 0, 0, -5, 0, -3, 0, 0, 0, 5, -5, 0, 0, 0, 3, 0, 5, 0, 0, 0, 0, -2]
    starts["player"] = -1
    starts["roll"] = np.random.default_rng(seed).integers(1, 7, size=(n, 2))
    return starts


@pytest.mark.parametrize("n", [1, 31, 384, 385, 1000, 5000])
def test_packed_playouts_equal_lane_resident_ones(ctx, oracle, monkeypatch, n):
    rng = np.random.default_rng(n)
    starts = positions.midgame_positions(seed=100 + n, n=min(n, 600), max_adv=150)
    starts = starts[rng.integers(0, len(starts), size=n)]
    monkeypatch.setenv("DIEE_LANE_PACK", "0")
    w0, p0, f0 = ctx.bg_playout(starts, seed=7, first_game_id=11, round_limit=400, want_finals=True)
    monkeypatch.setenv("DIEE_LANE_PACK", "2")
    w1, p1, f1 = ctx.bg_playout(starts, seed=7, first_game_id=11, round_limit=400, want_finals=True)
    assert w1.tobytes() == w0.tobytes() and p1.tobytes() == p0.tobytes() and f1.tobytes() == f0.tobytes()
    for g in np.linspace(0, n - 1, min(n, 24)).astype(int):
        w, p, s = oracle.bg_playout(starts[g:g + 1], 7, 11 + int(g), 400)
        assert (w1[g], p1[g]) == (w, p) and f1[g:g + 1].tobytes() == s.tobytes(), g
    # a short cap (nobody finishes), no plies at all, and starts that are already over
    w2, p2, f2 = ctx.bg_playout(starts, seed=7, first_game_id=11, round_limit=9, want_finals=True)
    monkeypatch.setenv("DIEE_LANE_PACK", "0")
    w3, p3, f3 = ctx.bg_playout(starts, seed=7, first_game_id=11, round_limit=9, want_finals=True)
    assert w2.tobytes() == w3.tobytes() and p2.tobytes() == p3.tobytes() and f2.tobytes() == f3.tobytes()
    monkeypatch.setenv("DIEE_LANE_PACK", "2")
    w4, p4 = ctx.bg_playout(f1, seed=1, round_limit=400)
    assert (p4[w1 != 0] == 0).all() and (w4[w1 != 0] == w1[w1 != 0]).all()
    w5, p5 = ctx.bg_playout(starts, seed=1, round_limit=0)
    assert (p5 == 0).all()


@pytest.mark.parametrize("mode", [0, 2])
def test_packed_rollouts_give_the_same_search(ctx, oracle, monkeypatch, mode):
    from die_e_b200 import _ffi
    states = positions.midgame_positions(seed=41, n=96, max_adv=140)
    players = states["player"].copy()
    cfg = oracle.mcts_cfg(iterations=40, c=2.0, limit=400, mode=mode)
    monkeypatch.setenv("DIEE_LANE_PACK", "0")
    ref = ctx.mcts_search(_ffi.GAME_BACKGAMMON, states, players, cfg, 5, 100, 3, dump=True)
    monkeypatch.setenv("DIEE_LANE_PACK", "2")
    got = ctx.mcts_search(_ffi.GAME_BACKGAMMON, states, players, cfg, 5, 100, 3, dump=True)
    assert ref[0].tobytes() == got[0].tobytes() and ref[1].tobytes() == got[1].tobytes()
    assert np.asarray(ref[6]).tobytes() == np.asarray(got[6]).tobytes()       # every rollout's end state
    assert ref[2]["rollout_plies"].tobytes() == got[2]["rollout_plies"].tobytes()
    for i in (0, 17, 95):
        rc, obest, onodes, ostates, ofin = oracle.mcts_search_bg(states[i:i + 1], int(players[i]), cfg, 5, 100 + i, 3, want_finals=True)
        assert got[1][i] == rc
        if rc == 0:
            assert got[6][i].tobytes() == ofin.tobytes(), i


def test_packed_kernel_is_what_a_large_job_runs(ctx, monkeypatch):
    """160,000 playouts (more than 1,024 items per SM): the default picks the packed kernel; same results as the
    lane-resident one, and the played-plies counter agrees"""
    n = 160000
    starts = _opening(n, seed=3)
    monkeypatch.setenv("DIEE_LANE_PACK", "0")
    w0, p0, f0 = ctx.bg_playout(starts, seed=0xD1EE, round_limit=100000, want_finals=True)
    work0 = ctx.search_work()
    launches0 = ctx.launch_count()
    monkeypatch.delenv("DIEE_LANE_PACK")
    w1, p1, f1 = ctx.bg_playout(starts, seed=0xD1EE, round_limit=100000, want_finals=True)
    work1 = ctx.search_work()
    assert w1.tobytes() == w0.tobytes() and p1.tobytes() == p0.tobytes() and f1.tobytes() == f0.tobytes()
    assert (w1 != 0).all() and work0 == work1 == int(p1.sum())
    assert ctx.launch_count() - launches0 == 1   # one launch for the whole job
