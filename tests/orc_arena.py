"""ORACLE twin of the arena (TEST INFRASTRUCTURE): `versus::play` (src/versus.rs:160-268) and
`get_actions_for_player` (:270-318) restated over the CPU oracle's primitives (tests/orc.py -> oracle/liborc.so),
one game at a time like the reference, with the injected stream contract of include/diee.h (see the header of
die_e_b200/versus.py for the arena's draw sites).  Shares no code with the product."""
import numpy as np

import orc

RANDOM, MCTS, MODEL = "random", "mcts", "model"


def _winner_or_none(s):
    return orc.bg_check_winner(s)


def play_backgammon(agent1, agent2, cfg, temp, seed, num_games, round_limit, eval_cb=None, max_nodes=0):
    games = {}
    for g in range(num_games):                       # versus.rs:170-181
        s = orc.bg_new()
        o = orc.philox(seed, 0, g, orc.STREAM_INIT, 0)
        if g >= num_games // 2:
            orc.bg_skip_turn(s, orc.die(o[0]), orc.die(o[1]))
            s["roll"][0] = (orc.die(o[2]), orc.die(o[3]))
        else:
            s["roll"][0] = (orc.die(o[0]), orc.die(o[1]))
        games[g] = s
    winners = np.zeros(num_games, dtype=np.int8)
    rounds = np.zeros(num_games, dtype=np.int32)
    wins_p1 = wins_p2 = 0
    round_count = 0
    while games:
        ids_p1 = sorted(g for g, s in games.items() if int(s["player"][0]) == -1)
        ids_p2 = sorted(g for g, s in games.items() if int(s["player"][0]) != -1)
        actions = {}
        for side, (agent, ids) in enumerate(((agent1, ids_p1), (agent2, ids_p2))):
            if not ids:
                continue
            if agent == RANDOM:                       # versus.rs:307-316
                for g in ids:
                    mv = orc.bg_valid_moves(games[g])
                    o = orc.philox(seed, round_count, g, orc.STREAM_GAME, 0)
                    actions[g] = mv[orc.index(o[2], len(mv))] if mv else []
            elif agent == MCTS:                       # versus.rs:303-306
                for g in ids:
                    rc, best, _, _ = orc.mcts_search_bg(games[g], int(games[g]["player"][0]), cfg, seed, g, round_count)
                    assert rc == 0, (g, rc)
                    actions[g] = orc.moves_to_list(best, 1)[0]
            else:                                     # versus.rs:277-302
                st = np.concatenate([games[g] for g in ids])
                nodes, n_nodes, status = orc.alpha_mcts_parallel(st, np.asarray(ids, dtype=np.uint32), cfg, seed,
                                                                 2 * round_count + side, eval_cb, max_nodes)
                assert (status == 0).all()
                for i, g in enumerate(ids):
                    ids_k, pi = orc.root_pi(nodes[i], 1.0 / temp)
                    if len(ids_k) == 0 or not (np.asarray(pi, dtype=np.float32).sum() != 0):
                        actions[g] = []
                        continue
                    a = orc.weighted_select(ids_k, pi, seed, g, round_count)
                    actions[g] = orc.bg_decode(games[g], a)
        round_count += 1
        for g in ids_p1 + ids_p2:                     # versus.rs:215-248
            s = games[g]
            o = orc.philox(seed, round_count - 1, g, orc.STREAM_GAME, 0)
            d0, d1 = orc.die(o[0]), orc.die(o[1])
            if not actions[g]:
                orc.bg_skip_turn(s, d0, d1)
                continue
            orc.bg_apply_move(s, orc.list_to_move(actions[g]), d0, d1)
            w = _winner_or_none(s)
            if w is None and round_count >= round_limit:
                w = 0
            if w is not None:
                winners[g], rounds[g] = w, round_count
                wins_p1 += w == -1
                wins_p2 += w == 1
                del games[g]
    return int(wins_p1), int(wins_p2), winners, rounds


def play_tictactoe(agent1, agent2, cfg, seed, num_games, round_limit):
    games = {}
    for g in range(num_games):
        s = orc.ttt_new()
        if g >= num_games // 2:
            s["player"] = -s["player"]
        games[g] = s
    winners = np.zeros(num_games, dtype=np.int8)
    rounds = np.zeros(num_games, dtype=np.int32)
    wins_p1 = wins_p2 = 0
    round_count = 0
    while games:
        ids_p1 = sorted(g for g, s in games.items() if int(s["player"][0]) == -1)
        ids_p2 = sorted(g for g, s in games.items() if int(s["player"][0]) != -1)
        actions = {}
        for agent, ids in ((agent1, ids_p1), (agent2, ids_p2)):
            for g in ids:
                if agent == RANDOM:
                    mv = orc.ttt_valid_moves(games[g])
                    o = orc.philox(seed, round_count, g, orc.STREAM_GAME, 0)
                    actions[g] = mv[orc.index(o[2], len(mv))] if mv else 10
                else:
                    rc, best, _, _ = orc.mcts_search_ttt(games[g], int(games[g]["player"][0]), cfg, seed, g, round_count)
                    assert rc == 0
                    actions[g] = best
        round_count += 1
        for g in ids_p1 + ids_p2:
            s = games[g]
            if actions[g] == 10:
                s["player"] = -s["player"]
                continue
            orc.ttt_apply_move(s, actions[g])
            w = orc.ttt_check_winner(s)
            if w is None and round_count >= round_limit:
                w = 0
            if w is not None:
                winners[g], rounds[g] = w, round_count
                wins_p1 += w == -1
                wins_p2 += w == 1
                del games[g]
    return int(wins_p1), int(wins_p2), winners, rounds
