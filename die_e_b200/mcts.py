"""Pure MCTS host API: `mct_search` (src/mcts/simple_mcts.rs:10-39) and its batched form -- the
reference maps `mct_search` over games with rayon (src/versus.rs:303-306); here one kernel launch
runs every game's search, one warp per game."""
import numpy as np

from . import _ffi
from .backgammon import Backgammon, _move_list
from .tictactoe import TicTacToe


class MctsConfig:
    """lib.rs:33-52 (+ mode flags exposing reference quirks Q5/Q6, default reference-exact)"""

    def __init__(self, iterations=100, c=2.0, simulate_round_limit=400, dirichlet_alpha=0.3, dirichlet_epsilon=0.25,
                 mode_flags=0):
        self.iterations, self.c, self.simulate_round_limit = iterations, c, simulate_round_limit
        self.dirichlet_alpha, self.dirichlet_epsilon, self.mode_flags = dirichlet_alpha, dirichlet_epsilon, mode_flags

    @classmethod
    def from_config(cls, conf):  # lib.rs:43-51 (keys of config-example.toml)
        return cls(int(conf["iterations"]), float(conf["exploration_const"]), int(conf["simulate_round_limit"]),
                   float(conf["dirichlet_alpha"]), float(conf["dirichlet_epsilon"]))

    def record(self):
        r = np.zeros(1, dtype=_ffi.MCTS_CFG)
        r[0] = (self.iterations, self.c, self.simulate_round_limit, self.dirichlet_alpha, self.dirichlet_epsilon,
                self.mode_flags)
        return r


def mct_search_batch(games, players, cfg, seed=0, first_game_id=0, epoch=0, ctx=None):
    """one independent `mct_search` per game (all of one kind); returns the list of chosen moves.
    A game whose search hits the reference's panic (node.rs:119-121) raises DieeError."""
    if not games:
        return []
    ctx = ctx or _ffi.default_context()
    if isinstance(games[0], Backgammon):
        kind, states = _ffi.GAME_BACKGAMMON, np.concatenate([g.s for g in games])
    else:
        kind, states = _ffi.GAME_TICTACTOE, np.concatenate([g.s for g in games])
    best, status, _ = ctx.mcts_search(kind, states, np.array(players, dtype=np.int8), cfg.record(), seed, first_game_id, epoch)
    for i, st in enumerate(status):
        if st == _ffi.ERR_NO_MOVES_PANIC:
            raise _ffi.DieeError(int(st), f"game {i}: expand() called on node with no expandable moves")
        if st != _ffi.OK:
            raise _ffi.DieeError(int(st), f"game {i}: search failed")
    if kind == _ffi.GAME_BACKGAMMON:
        return [_move_list(best[i]) for i in range(len(games))]
    return [int(b) for b in best]


def mct_search(state, player, cfg, seed=0, game_id=0, epoch=0, ctx=None):
    """`pub fn mct_search<T>(state: T, player: i8, mcts_config: &MctsConfig) -> T::Move`"""
    return mct_search_batch([state], [player], cfg, seed, game_id, epoch, ctx)[0]
