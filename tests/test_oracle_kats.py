"""The reference's own tests (transcribed, tests/golden/) run against the CPU oracle.

This is the oracle's pin: every vector in tests/backgammon_test.rs, tests/tictactoe_test.rs
and tests/encoding_test.rs of alibasaran/die-e.  (The reference's fourth test file, tests/mcts_test.rs, holds two
normalisation invariants and no vectors; they run on the oracle in tests/test_alpha_host.py.)
"""
import json
import os

import pytest

import kat_shim
import ref_backgammon_kats
import ref_tictactoe_kats

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def shim(oracle):
    return kat_shim.make_shim(kat_shim.OracleBackend())


@pytest.mark.parametrize("name,fn,stale", ref_backgammon_kats.CASES, ids=[c[0] for c in ref_backgammon_kats.CASES])
def test_reference_backgammon_test(shim, name, fn, stale):
    if stale:
        # tests/backgammon_test.rs:918-925 contradicts the current source (doubles are two
        # 2-move plays, backgammon_logic.rs:406-409,179-185): the source is the truth.
        with pytest.raises(AssertionError):
            fn(shim)
        bg = shim.Backgammon.new()
        bg.board[0] = [0] * 24
        bg.board[0][20] = -1
        bg.roll = [1, 1]
        kat_shim.assert_eq(bg.get_valid_moves(), [[(20, 19), (19, 18)]])
    else:
        fn(shim)


@pytest.mark.parametrize("name,fn,stale", ref_tictactoe_kats.CASES, ids=[c[0] for c in ref_tictactoe_kats.CASES])
def test_reference_tictactoe_test(shim, name, fn, stale):
    fn(shim)


def test_reference_encoding_round_trips(oracle):
    cases = json.load(open(os.path.join(GOLDEN, "ref_encoding_kats.json")))["cases"]
    assert len(cases) == 46
    for c in cases:
        s = oracle.make_state([0] * 24, roll=c["roll"], player=c["player"])
        acts = [tuple(a) for a in c["actions"]]
        enc = oracle.bg_encode(s, acts)
        assert 0 <= enc <= 1351
        assert oracle.bg_decode(s, enc) == acts, c
    # the one absolute id the reference's test names imply (encoding_test.rs:80 "..._when_enc_is_0")
    s = oracle.make_state([0] * 24, roll=(2, 1), player=-1)
    assert oracle.bg_encode(s, [(0, -1), (0, -1)]) == 0
    assert oracle.bg_encode(s, []) == 1351
