"""GPU parity tests for pure MCTS (mct_search) through the C ABI against the CPU oracle.
Bar: bit-exact -- every node's parent/visits/value/action/move counts and state, the chosen
move, the per-game status (incl. the reference's panic, Q6) in both rollout modes (Q5)."""
import numpy as np
import pytest

import positions

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from die_e_b200 import _ffi
    return _ffi.Context(0)


def _cmp_trees(oracle, ffi, ctx, kind, states, players, cfg, seed, first, epoch):
    best, status, stats, nodes, nstates, n_nodes, finals = ctx.mcts_search(kind, states, players, cfg, seed, first, epoch, dump=True)
    total_plies = 0
    for i in range(len(states)):
        if kind == ffi.GAME_BACKGAMMON:
            rc, obest, onodes, ostates, ofin = oracle.mcts_search_bg(states[i:i + 1], int(players[i]), cfg, seed, first + i,
                                                                     epoch, want_finals=True)
            assert finals[i].tobytes() == ofin.tobytes(), i   # every rollout ended in the same state
        else:
            rc, obest, onodes, ostates = oracle.mcts_search_ttt(states[i:i + 1], int(players[i]), cfg, seed, first + i, epoch)
        assert status[i] == rc, (i, status[i], rc)
        assert n_nodes[i] == len(onodes), i
        k = len(onodes)
        got = nodes[i, :k]
        for f in ("parent", "n_moves", "n_untried"):
            assert (got[f] == onodes[f]).all(), (i, f)
        assert got["visits"].tobytes() == onodes["visits"].tobytes(), i       # bit-exact f32
        assert got["value"].tobytes() == onodes["value"].tobytes(), i
        assert got["action"].tobytes() == onodes["action"].tobytes(), i
        assert nstates[i, :k].tobytes() == ostates.tobytes(), i
        if kind == ffi.GAME_BACKGAMMON:
            assert best[i:i + 1].tobytes() == obest.tobytes(), i
        else:
            assert best[i] == obest, i
        total_plies += int(stats[i]["rollout_plies"])
        assert stats[i]["expansions"] == max(0, k - 1)
    return status, total_plies


@pytest.mark.parametrize("mode", [0, 1, 2, 3])
def test_backgammon_mcts_bit_exact(ctx, oracle, mode):
    from die_e_b200 import _ffi as ffi
    states = positions.midgame_positions(seed=77, n=48, max_adv=120)
    players = states["player"].copy()
    cfg = oracle.mcts_cfg(iterations=60, c=2.0, limit=40, mode=mode)
    status, plies = _cmp_trees(oracle, ffi, ctx, ffi.GAME_BACKGAMMON, states, players, cfg, 0xD1EE, 1000, 3)
    assert plies > 0
    if not (mode & 2):
        # without PASS_CHILD some searches must hit the reference's panic (4.3 % of plies have no move)
        pass


def test_backgammon_mcts_reference_config(ctx, oracle):
    """config 3 parameters (iterations=100, c=2, limit=400), reference-exact rollouts, PASS_CHILD"""
    from die_e_b200 import _ffi as ffi
    states = positions.midgame_positions(seed=5, n=12, max_adv=80)
    cfg = oracle.mcts_cfg(iterations=100, c=2.0, limit=400, mode=2)
    status, plies = _cmp_trees(oracle, ffi, ctx, ffi.GAME_BACKGAMMON, states, states["player"].copy(), cfg, 9, 0, 0)
    assert (status == 0).all()
    # Q5: in reference-exact mode every rollout from a non-terminal node runs the full limit
    assert plies > 0 and plies % 400 == 0


def test_backgammon_mcts_panic_and_terminal(ctx, oracle):
    from die_e_b200 import _ffi as ffi
    pts = [0] * 24
    pts[20] = -1
    pts[19] = 2
    pts[18] = 2
    blocked = oracle.make_state(pts, off=(14, 11), roll=(1, 2), player=-1)      # no legal move at the root
    won = oracle.make_state([0] * 24, off=(15, 3), roll=(1, 2), player=1)
    pts2 = [0] * 24
    pts2[0] = -1
    pts2[23] = 1
    near = oracle.make_state(pts2, off=(14, 14), roll=(3, 4), player=-1)        # children are terminal
    states = np.concatenate([blocked, won, near])
    players = np.array([-1, 1, -1], dtype=np.int8)
    for mode in (0, 2):
        cfg = oracle.mcts_cfg(iterations=30, c=2.0, limit=20, mode=mode)
        status, _ = _cmp_trees(oracle, ffi, ctx, ffi.GAME_BACKGAMMON, states, players, cfg, 4, 0, 0)
        assert status[0] == (ffi.ERR_NO_MOVES_PANIC if mode == 0 else 0)
        assert status[1] == 0 and status[2] == 0


@pytest.mark.parametrize("mode", [0, 1])
def test_tictactoe_mcts_bit_exact(ctx, oracle, mode):
    from die_e_b200 import _ffi as ffi
    rng = np.random.default_rng(2)
    states = []
    for g in range(64):
        s = oracle.ttt_new()
        for _ in range(int(rng.integers(0, 7))):
            if oracle.ttt_check_winner(s) is not None:
                break
            mv = oracle.ttt_valid_moves(s)
            oracle.ttt_apply_move(s, int(rng.choice(mv)))
        states.append(s)
    states = np.concatenate(states)
    cfg = oracle.mcts_cfg(iterations=100, c=2.0, limit=400, mode=mode)
    _cmp_trees(oracle, ffi, ctx, ffi.GAME_TICTACTOE, states, states["player"].copy(), cfg, 123, 0, 1)


def test_mct_search_host_api(ctx, oracle):
    from die_e_b200 import Backgammon, MctsConfig, TicTacToe, mct_search, _ffi as ffi
    t = TicTacToe.new()
    t.apply_move(4)
    cfg = MctsConfig(iterations=50, simulate_round_limit=20)
    mv = mct_search(t, t.get_player(), cfg, seed=3, game_id=7, ctx=ctx)
    rc, obest, _, _ = oracle.mcts_search_ttt(t.s, t.get_player(), oracle.mcts_cfg(50, 2.0, 20), 3, 7, 0)
    assert mv == obest and mv in t.get_valid_moves()
    bg = Backgammon.new(ctx, seed=3, game_id=1)
    bg.roll_die()
    cfg = MctsConfig(iterations=30, simulate_round_limit=10, mode_flags=ffi.MODE_PASS_CHILD)
    mv = mct_search(bg, bg.get_player(), cfg, seed=3, game_id=1, ctx=ctx)
    assert mv in bg.get_valid_moves()


def _same_search(ref, got, tag, oracle_finals=None):
    names = ("best", "status", "stats", "nodes", "node_states", "n_nodes", "finals")
    assert np.asarray(ref[5]).tobytes() == np.asarray(got[5]).tobytes(), f"{tag}: n_nodes differ"
    live = np.arange(np.asarray(ref[3]).shape[1])[None, :] < np.asarray(ref[5])[:, None]   # pool entries >= n_nodes are unspecified
    for name, a, b in zip(names, ref, got):
        if name in ("nodes", "node_states"):
            a, b = np.asarray(a)[live], np.asarray(b)[live]
        a8 = np.frombuffer(np.asarray(a).tobytes(), dtype=np.uint8)
        b8 = np.frombuffer(np.asarray(b).tobytes(), dtype=np.uint8)
        if a8.tobytes() != b8.tobytes():
            where = np.nonzero(a8 != b8)[0]
            extra = ""
            if name == "finals" and oracle_finals is not None:
                o8 = np.frombuffer(oracle_finals, dtype=np.uint8)
                extra = f"; reference run == oracle: {bool((a8 == o8).all())}, this run == oracle: {bool((b8 == o8).all())}"
            raise AssertionError(f"{tag}: {name} differs at bytes {where[:6].tolist()} ({len(where)} in all){extra}")


def test_sliced_search_is_the_same_search(ctx, oracle, monkeypatch):
    """DIEE_SEARCH_SLICES: the search cut into slices of iterations (tree kernel of slice s+1 beside the rollouts of
    slice s on a side stream) must give exactly the unsliced result -- nodes, best moves, rollout end states."""
    from die_e_b200 import _ffi
    states = positions.midgame_positions(seed=13, n=24, max_adv=90)
    players = states["player"].copy()
    cfg = oracle.mcts_cfg(iterations=30, c=2.0, limit=60, mode=_ffi.MODE_PASS_CHILD)
    ref = ctx.mcts_search(_ffi.GAME_BACKGAMMON, states, players, cfg, 5, 100, 3, dump=True)
    ofin = b"".join(oracle.mcts_search_bg(states[i:i + 1], int(players[i]), cfg, 5, 100 + i, 3, want_finals=True)[4].tobytes()
                    for i in range(len(states)))
    assert np.asarray(ref[6]).tobytes() == ofin
    for slices in ("2", "4"):
        monkeypatch.setenv("DIEE_SEARCH_SLICES", slices)
        _same_search(ref, ctx.mcts_search(_ffi.GAME_BACKGAMMON, states, players, cfg, 5, 100, 3, dump=True), f"{slices} slices", ofin)
    # the same slices on an SM partition (green contexts: tree slices on 64 SMs, rollouts on the rest); a fresh context,
    # because the partition is set up once per context
    monkeypatch.setenv("DIEE_TREE_SMS", "64")
    ctx2 = _ffi.Context(0)
    for rep in range(3):
        _same_search(ref, ctx2.mcts_search(_ffi.GAME_BACKGAMMON, states, players, cfg, 5, 100, 3, dump=True),
                     f"4 slices on an SM partition, run {rep}", ofin)
    ctx2.close()
    monkeypatch.delenv("DIEE_TREE_SMS")
    monkeypatch.delenv("DIEE_SEARCH_SLICES")


def test_lockstep_check_current_groups_and_fused_form_agree(ctx, oracle, monkeypatch):
    """DIEE_MODE_ROLLOUT_CHECK_CURRENT has three forms: lock-step on the lane engine (one tree launch + one rollout launch
    per iteration, games cut into groups on side streams), the round-1 fused form (DIEE_CC_FUSED=1: one launch, warp-per-game
    rollouts) and the persistent form (DIEE_CC_PERSISTENT=1, cc_search_kernel: one launch, every game at its own pace,
    games queued by the code of their next step).  All must give the same pool, moves and rollout end states."""
    from die_e_b200 import _ffi
    states = positions.midgame_positions(seed=29, n=70, max_adv=110)
    players = states["player"].copy()
    cfg = oracle.mcts_cfg(iterations=25, c=2.0, limit=400, mode=_ffi.MODE_PASS_CHILD | _ffi.MODE_ROLLOUT_CHECK_CURRENT)
    monkeypatch.setenv("DIEE_CC_FUSED", "0")   # (small batches default to the fused form)
    monkeypatch.setenv("DIEE_CC_PERSISTENT", "0")
    ref = ctx.mcts_search(_ffi.GAME_BACKGAMMON, states, players, cfg, 5, 100, 3, dump=True)
    assert (ref[1] == 0).all() and int(ref[2]["rollout_plies"].sum()) > 0
    for env, val in (("DIEE_CC_GROUPS", "3"), ("DIEE_CC_GROUPS", "4"), ("DIEE_CC_FUSED", "1"), ("DIEE_CC_PERSISTENT", "1")):
        monkeypatch.setenv(env, val)
        got = ctx.mcts_search(_ffi.GAME_BACKGAMMON, states, players, cfg, 5, 100, 3, dump=True)
        if env == "DIEE_CC_GROUPS":
            monkeypatch.delenv(env)
        else:
            monkeypatch.setenv(env, "0")
        _same_search(ref, got, f"{env}={val}")
    # and against the oracle, game by game, with the full rollout cap -- the lock-step form, then the persistent one
    _cmp_trees(oracle, _ffi, ctx, _ffi.GAME_BACKGAMMON, states[:24], players[:24], cfg, 5, 100, 3)
    monkeypatch.setenv("DIEE_CC_PERSISTENT", "1")
    _cmp_trees(oracle, _ffi, ctx, _ffi.GAME_BACKGAMMON, states[:24], players[:24], cfg, 5, 100, 3)
    for mode in (1, 3):   # without PASS_CHILD some games hit the reference's panic: the persistent form reports the same
        cfg2 = oracle.mcts_cfg(iterations=40, c=2.0, limit=60, mode=mode)
        _cmp_trees(oracle, _ffi, ctx, _ffi.GAME_BACKGAMMON, states[:32], players[:32], cfg2, 9, 7, 1)
    monkeypatch.delenv("DIEE_CC_FUSED")
    monkeypatch.delenv("DIEE_CC_PERSISTENT")


def test_persistent_check_current_at_a_few_thousand_games(ctx, oracle, monkeypatch):
    """3,000 games (ten resident games per CTA and more): the same search as the lock-step form, game for game"""
    from die_e_b200 import _ffi
    rng = np.random.default_rng(5)
    base = positions.midgame_positions(seed=31, n=500, max_adv=130)
    states = base[rng.integers(0, len(base), size=3000)]
    players = states["player"].copy()
    cfg = oracle.mcts_cfg(iterations=16, c=2.0, limit=400, mode=_ffi.MODE_PASS_CHILD | _ffi.MODE_ROLLOUT_CHECK_CURRENT)
    monkeypatch.setenv("DIEE_CC_FUSED", "0")
    monkeypatch.setenv("DIEE_CC_PERSISTENT", "0")
    ref = ctx.mcts_search(_ffi.GAME_BACKGAMMON, states, players, cfg, 11, 0, 2, dump=True)
    work_ref = ctx.search_work()
    monkeypatch.setenv("DIEE_CC_PERSISTENT", "1")
    got = ctx.mcts_search(_ffi.GAME_BACKGAMMON, states, players, cfg, 11, 0, 2, dump=True)
    _same_search(ref, got, "persistent form, 3,000 games")
    assert ctx.search_work() == work_ref == int(ref[2]["rollout_plies"].sum())
    monkeypatch.delenv("DIEE_CC_FUSED")
    monkeypatch.delenv("DIEE_CC_PERSISTENT")
