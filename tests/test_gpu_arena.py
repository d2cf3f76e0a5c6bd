"""GPU parity tests of the arena (`versus::play`, src/versus.rs:160-318 -- SURVEY 8(f) rank 1) against its oracle
twin (tests/orc_arena.py: the same rounds played one game at a time on the CPU oracle).  Bar: identical -- every
game's winner and the round it ended in, hence wins / draws / winrate."""
import numpy as np
import pytest

import orc_arena

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from die_e_b200 import _ffi
    return _ffi.Context(0)


def _same(res, want):
    w1, w2, winners, rounds = want
    assert (res.winners == winners).all() and (res.rounds == rounds).all()
    assert (res.wins_p1, res.wins_p2, res.draws) == (w1, w2, res.n_games - w1 - w2)


def test_backgammon_random_vs_random(ctx, oracle):
    from die_e_b200.versus import Agent, Player, play
    res = play("backgammon", Player(Agent.Random), Player(Agent.Random), seed=11, num_games=64, round_limit=400, ctx=ctx)
    _same(res, orc_arena.play_backgammon("random", "random", None, 1.0, 11, 64, 400))
    assert res.wins_p1 + res.wins_p2 + res.draws == 64 and res.draws <= 2
    # the round limit: games still running when it is reached are draws once they have played a move
    res = play("backgammon", Player(Agent.Random), Player(Agent.Random), seed=12, num_games=16, round_limit=30, ctx=ctx)
    _same(res, orc_arena.play_backgammon("random", "random", None, 1.0, 12, 16, 30))
    assert res.draws > 0 and "Winrate" in str(res)


def test_backgammon_mcts_vs_random(ctx, oracle):
    from die_e_b200 import _ffi
    from die_e_b200.mcts import MctsConfig
    from die_e_b200.versus import Agent, Player, play
    for mode in (_ffi.MODE_PASS_CHILD, _ffi.MODE_PASS_CHILD | _ffi.MODE_ROLLOUT_CHECK_CURRENT):
        cfg = MctsConfig(iterations=16, c=2.0, simulate_round_limit=40, mode_flags=mode)
        ocfg = oracle.mcts_cfg(iterations=16, c=2.0, limit=40, mode=mode)
        res = play("backgammon", Player(Agent.Mcts), Player(Agent.Random), cfg, seed=21, num_games=12, round_limit=400, ctx=ctx)
        _same(res, orc_arena.play_backgammon("mcts", "random", ocfg, 1.0, 21, 12, 400))
        res = play("backgammon", Player(Agent.Random), Player(Agent.Mcts), cfg, seed=22, num_games=8, round_limit=400, ctx=ctx)
        _same(res, orc_arena.play_backgammon("random", "mcts", ocfg, 1.0, 22, 8, 400))


def test_tictactoe_mcts_vs_random_is_config_1(ctx, oracle):
    """BASELINE configs[0]: Tic-Tac-Toe MCTS agent vs random, iterations=100, random rollouts (`die-e -g tic-tac-toe play`)"""
    from die_e_b200.mcts import MctsConfig
    from die_e_b200.versus import Agent, Player, play
    cfg = MctsConfig(iterations=100, c=2.0, simulate_round_limit=400)
    ocfg = oracle.mcts_cfg(iterations=100, c=2.0, limit=400)
    res = play("tictactoe", Player(Agent.Mcts), Player(Agent.Random), cfg, seed=31, num_games=60, ctx=ctx)
    _same(res, orc_arena.play_tictactoe("mcts", "random", ocfg, 31, 60, 400))
    # the reference's own size: 400 games, first 200 started by player 1
    import time
    t0 = time.perf_counter()
    full = play("tictactoe", Player(Agent.Mcts), Player(Agent.Random), cfg, seed=0xD1EE, ctx=ctx)
    dt = time.perf_counter() - t0
    assert full.n_games == 400 and full.wins_p1 + full.wins_p2 + full.draws == 400
    t0 = time.perf_counter()
    orc_arena.play_tictactoe("mcts", "random", ocfg, 0xD1EE, 400, 400)
    dt_cpu = time.perf_counter() - t0
    print(f"\n[C1] tictactoe MCTS(100) vs random, 400 games: {full.wins_p1} / {full.wins_p2} / {full.draws} "
          f"(winrate {full.winrate:.3f}); arena wall time {dt * 1e3:.0f} ms on the GPU engine, {dt_cpu * 1e3:.0f} ms oracle twin (1 core)")


def test_backgammon_model_vs_random(ctx, oracle):
    from die_e_b200 import _ffi, nnet
    from die_e_b200.mcts import MctsConfig
    from die_e_b200.versus import Agent, Player, play
    model = nnet.ResNet.new(seed=41, filters=128, blocks=1, bn_stats="random", ctx=ctx)
    cfg = MctsConfig(iterations=8, c=2.0, simulate_round_limit=400, dirichlet_alpha=0.3, dirichlet_epsilon=0.25)
    ocfg = oracle.mcts_cfg(iterations=8, c=2.0, limit=400, alpha=0.3, eps=0.25)
    cb = oracle.make_eval(lambda st: model._net.forward(st))   # identical net outputs on both sides
    res = play("backgammon", Player(Agent.Model, model), Player(Agent.Random), cfg, temp=1.25, seed=51, num_games=6,
               round_limit=60, ctx=ctx)
    _same(res, orc_arena.play_backgammon("model", "random", ocfg, 1.25, 51, 6, 60, eval_cb=cb, max_nodes=1 + 9 * 60))


def test_saved_games_round_trip(ctx, oracle, tmp_path):
    """Game<T> JSON (versus.rs:18-71,107-122) and the replay printout"""
    import json
    from die_e_b200.versus import Agent, Player, load_all_games, play, print_game, save_game
    res = play("backgammon", Player(Agent.Random), Player(Agent.Random), seed=7, num_games=6, ctx=ctx, keep_games=True,
               record_turns=True)
    assert len(res.games) == 6 and len({g.id for g in res.games}) == 6 and all(len(g.id) == 21 for g in res.games)
    for g in res.games:
        save_game(g, tmp_path)
    raw = json.load(open(tmp_path / f"{res.games[0].id}.json"))
    assert set(raw) == {"id", "player1", "player2", "turns", "winner", "initial_state"}
    assert set(raw["initial_state"]) == {"board", "roll", "player", "is_second_play", "id"}
    assert len(raw["initial_state"]["board"][0]) == 24 and raw["player1"] == "Random" and raw["winner"] in ("Random", "None")
    back = load_all_games(tmp_path)
    assert sorted(g.id for g in back) == sorted(g.id for g in res.games)
    # replaying the recorded turns on the oracle from the initial state reproduces the game's winner
    g0 = res.games[0]
    st = oracle.make_state(g0.initial_state["board"][0], tuple(g0.initial_state["board"][1]), tuple(g0.initial_state["board"][2]),
                           tuple(g0.initial_state["roll"]), g0.initial_state["player"], g0.initial_state["is_second_play"])
    idx = g0.initial_state["id"]
    for r, turn in enumerate(g0.turns):
        assert list(st["roll"][0]) == turn["roll"]
        o = oracle.philox(7, r, idx, oracle.STREAM_GAME, 0)
        if turn["action"]:
            oracle.bg_apply_move(st, oracle.list_to_move([tuple(p) for p in turn["action"]]), oracle.die(o[0]), oracle.die(o[1]))
        else:
            oracle.bg_skip_turn(st, oracle.die(o[0]), oracle.die(o[1]))
    assert oracle.bg_check_winner(st) == res.winners[idx]
    lines = []
    print_game(tmp_path / f"{g0.id}.json", out=lines.append)
    assert lines[0] == f"Game ID: {g0.id}" and any(l.startswith("Action:") for l in lines)


def test_device_resident_arena_equals_the_host_loop(ctx, oracle):
    """csrc/arena.cu (games resident in HBM, one 20-byte read-back per round) against the host-side loop of
    die_e_b200/versus.py -- which the tests above hold to the oracle twin: same winners, same rounds; plus the reference's
    own arena size (400 games, MCTS vs random) timed both ways"""
    import time
    from die_e_b200 import _ffi
    from die_e_b200.mcts import MctsConfig
    from die_e_b200.versus import Agent, Player, play
    cfg = MctsConfig(iterations=16, c=2.0, simulate_round_limit=40, mode_flags=_ffi.MODE_PASS_CHILD)
    for p1, p2, seed in ((Agent.Mcts, Agent.Random, 61), (Agent.Random, Agent.Mcts, 62), (Agent.Mcts, Agent.Mcts, 63)):
        dev = play("backgammon", Player(p1), Player(p2), cfg, seed=seed, num_games=24, round_limit=400, ctx=ctx, device_resident=True)
        host = play("backgammon", Player(p1), Player(p2), cfg, seed=seed, num_games=24, round_limit=400, ctx=ctx, device_resident=False)
        assert (dev.winners == host.winners).all() and (dev.rounds == host.rounds).all()
        assert (dev.wins_p1, dev.wins_p2, dev.draws) == (host.wins_p1, host.wins_p2, host.draws)
    # without PASS_CHILD the reference panics on a no-move node (node.rs:119-121): the arena reports it
    bad = MctsConfig(iterations=16, c=2.0, simulate_round_limit=40, mode_flags=0)
    with pytest.raises(_ffi.DieeError) as e:
        play("backgammon", Player(Agent.Mcts), Player(Agent.Mcts), bad, seed=64, num_games=64, round_limit=400, ctx=ctx)
    assert e.value.code == _ffi.ERR_NO_MOVES_PANIC
    # versus.rs:168-169: 400 games, 400 rounds; MCTS(iterations=100, limit=400) vs random
    cfg = MctsConfig(iterations=100, c=2.0, simulate_round_limit=400, mode_flags=_ffi.MODE_PASS_CHILD)
    t0 = time.perf_counter()
    dev = play("backgammon", Player(Agent.Mcts), Player(Agent.Random), cfg, seed=0xD1EE, ctx=ctx)
    t_dev = time.perf_counter() - t0
    t0 = time.perf_counter()
    host = play("backgammon", Player(Agent.Mcts), Player(Agent.Random), cfg, seed=0xD1EE, ctx=ctx, device_resident=False)
    t_host = time.perf_counter() - t0
    assert (dev.winners == host.winners).all() and (dev.rounds == host.rounds).all() and dev.n_games == 400
    print(f"\n[arena] backgammon MCTS(100) vs random, 400 games, {int(dev.rounds.max())} rounds: {dev.wins_p1} / {dev.wins_p2} / {dev.draws}; "
          f"device-resident {t_dev * 1e3:.0f} ms, host loop {t_host * 1e3:.0f} ms")
