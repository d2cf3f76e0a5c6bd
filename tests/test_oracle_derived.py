"""Derived vectors: not in the reference's tests; they pin the quirks
Q1/Q3/Q4 and absolute action ids, computed by an independent restatement at survey time."""
import numpy as np


def _board(d):
    pts = [0] * 24
    for k, val in d.items():
        pts[k] = val
    return pts


def test_q1_no_use_max_dice(oracle):
    s = oracle.make_state(_board({12: -1, 6: -1, 4: 2, 3: 2, 1: 2}), off=(13, 0), roll=(6, 2), player=-1)
    assert oracle.bg_valid_moves(s) == [[(12, 10), (6, 0)], [(12, 6)]]


def test_q3_signed_sum_bearoff(oracle):
    s = oracle.make_state(_board({0: -1, 3: 2, 5: -1}), off=(13, 0), roll=(5, 2), player=-1)
    assert oracle.bg_valid_moves(s) == [[(0, -1), (5, 0)], [(0, -1)]]
    s = oracle.make_state(_board({0: -1, 5: -1}), off=(13, 0), roll=(5, 2), player=-1)
    assert oracle.bg_valid_moves(s) == [[(5, 3), (3, -1)]]
    s = oracle.make_state(_board({20: 1, 22: 1}), off=(0, 13), roll=(6, 1), player=1)
    assert oracle.bg_valid_moves(s) == [[(20, 21), (21, -1)], [(22, 23), (20, -1)]]


def test_q4_doubles_are_two_move_plays(oracle):
    s = oracle.bg_new()
    s["roll"][0] = (6, 6)
    assert oracle.bg_valid_moves(s) == [[(7, 1), (7, 1)], [(7, 1), (12, 6)], [(7, 1), (23, 17)],
                                        [(12, 6), (12, 6)], [(12, 6), (23, 17)], [(23, 17), (23, 17)]]
    s = oracle.make_state(_board({20: -1}), roll=(1, 1), player=-1)
    assert oracle.bg_valid_moves(s) == [[(20, 19), (19, 18)]]


def test_opening_move_counts(oracle):
    want = [[9, 15, 16, 14, 8, 10], [15, 13, 17, 18, 8, 14], [16, 17, 13, 17, 9, 14],
            [14, 18, 17, 12, 9, 14], [8, 8, 9, 9, 3, 7], [10, 14, 14, 14, 7, 6]]
    for r0 in range(1, 7):
        for r1 in range(1, 7):
            s = oracle.bg_new()
            s["roll"][0] = (r0, r1)
            assert len(oracle.bg_valid_moves(s)) == want[r0 - 1][r1 - 1], (r0, r1)


def test_absolute_action_ids(oracle):
    vec = [((2, 1), -1, [(4, 2)], 654), ((2, 1), -1, [(4, 3)], 1330), ((2, 1), -1, [(-1, 22)], 674),
           ((2, 1), -1, [(-1, 23)], 1350), ((2, 1), -1, [(1, -1)], 651), ((2, 1), -1, [(0, -1)], 1326),
           ((2, 1), -1, [(23, 21), (5, 4)], 153), ((2, 1), -1, [(5, 4), (23, 21)], 1279),
           ((2, 1), 1, [(-1, 1), (-1, 0)], 648), ((2, 1), 1, [(-1, 0), (-1, 1)], 1324),
           ((6, 1), -1, [(-1, 18), (18, 17)], 492), ((6, 1), -1, [(-1, 23), (23, 17)], 1298),
           ((6, 1), 1, [(21, -1)], 671), ((4, 5), 1, [(0, 4), (0, 5)], 676),
           ((5, 3), 1, [(22, -1), (18, 21)], 490), ((6, 3), 1, [(22, -1)], 1348), ((6, 3), 1, [(21, -1)], 1347)]
    for roll, player, acts, want in vec:
        s = oracle.make_state([0] * 24, roll=roll, player=player)
        assert oracle.bg_encode(s, acts) == want, (roll, player, acts)


def test_as_tensor_layout(oracle):
    s = oracle.bg_new()
    s["roll"][0] = (3, 5)
    s["bar"][0] = (1, 2)
    s["off"][0] = (4, 6)
    s["second"][0] = 1
    t = oracle.bg_as_tensor(s)
    assert t.shape == (1, 6, 4, 6)
    assert t[0, 0].reshape(-1).tolist() == [float(x) for x in s["pts"][0]]
    assert (t[0, 1] == -1).all()
    assert (t[0, 2, :2] == 1).all() and (t[0, 2, 2:] == 2).all()
    assert (t[0, 3, :2] == 4).all() and (t[0, 3, 2:] == 6).all()
    assert (t[0, 4, :2] == 3).all() and (t[0, 4, 2:] == 5).all()
    assert (t[0, 5] == 1).all()


def test_random_games_invariants(oracle):
    """checker conservation + codec round trip and injectivity on reachable positions"""
    rng = np.random.default_rng(1)
    for g in range(40):
        s = oracle.bg_new()
        blk = oracle.philox(7, 0, g, oracle.STREAM_INIT, 0)
        s["roll"][0] = (oracle.die(blk[0]), oracle.die(blk[1]))
        for ply in range(500):
            if oracle.bg_check_winner(s) is not None:
                break
            mv = oracle.bg_valid_moves(s)
            ids = [oracle.bg_encode(s, m) for m in mv]
            assert len(set(ids)) == len(ids)
            for m, i in zip(mv, ids):
                assert oracle.bg_decode(s, i) == m
            oracle.bg_random_ply(s, oracle.philox(7, ply, g, oracle.STREAM_GAME, 0))
            p = s["pts"][0].astype(int)
            assert -p[p < 0].sum() + s["bar"][0][0] + s["off"][0][0] == 15
            assert p[p > 0].sum() + s["bar"][0][1] + s["off"][0][1] == 15
        assert oracle.bg_check_winner(s) is not None
