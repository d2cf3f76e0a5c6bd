"""Seeded reachable backgammon positions for differential tests (generated with the CPU oracle)."""
import numpy as np

import orc


def start_state(seed, game_id):
    s = orc.bg_new()
    w = orc.philox(seed, 0, game_id, orc.STREAM_INIT, 0)
    s["roll"][0] = (orc.die(w[0]), orc.die(w[1]))
    return s


def reachable_positions(seed, n_games, max_plies=500, every=1):
    """all positions met in n_games random games (before each ply), as one BG_STATE array"""
    out = []
    for g in range(n_games):
        s = start_state(seed, g)
        for ply in range(max_plies):
            if orc.bg_check_winner(s) is not None:
                break
            if ply % every == 0:
                out.append(s.copy())
            orc.bg_random_ply(s, orc.philox(seed, ply, g, orc.STREAM_GAME, 0))
        out.append(s.copy())  # terminal position (still has a roll)
    return np.concatenate(out)


def midgame_positions(seed, n, max_adv=80):
    """SURVEY 8(d) M-inputs: game g advanced k ~ U{0..max_adv} random plies"""
    rng = np.random.default_rng(seed)
    out = []
    for g in range(n):
        s = start_state(seed, g)
        k = int(rng.integers(0, max_adv + 1))
        for ply in range(k):
            if orc.bg_check_winner(s) is not None:
                break
            orc.bg_random_ply(s, orc.philox(seed, ply, g, orc.STREAM_GAME, 0))
        out.append(s.copy())
    return np.concatenate(out)
