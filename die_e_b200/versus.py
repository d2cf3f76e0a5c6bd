"""The arena: `versus::play` and `get_actions_for_player` (src/versus.rs:124-318) over the GPU engine.

The reference keeps 400 games in a HashMap, and every round partitions the live games by the side to
move, asks each side's agent for one action per game (rayon `par_iter` over `mct_search` for Agent::Mcts,
a uniform choice for Agent::Random, `alpha_mcts_parallel` + temperature + categorical draw for
Agent::Model), applies the actions and retires finished games.  Here a round is a handful of batched
calls: ONE search launch for all of a side's games (one warp per game), one move-generation launch for a
random side, one apply launch for the whole round.  What stays on the host is the bookkeeping the
reference also does on the host (partition, win/draw accounting, the 400/400 constants).

Behaviour kept from the reference (versus.rs:168-268):
  * games with index >= n/2 start with `skip_turn` (the other side begins) and, for dice games, one more roll;
  * an EMPTY_MOVE action is a `skip_turn` and is NOT followed by a winner or round-limit test (:222-225);
  * after an applied move: winner, else a draw once `round_count >= round_limit` (:231-235);
  * player 1 is the side -1; winrate = wins_p1 / n_games; draws = n_games - wins.

Randomness: the reference draws from `thread_rng()` in HashMap order; here every draw is keyed by the
game's index (include/diee.h stream contract), so results do not depend on batch composition:
  INIT    block (0, game):      o0,o1 = first roll (first half) or the skip_turn's roll (second half), o2,o3 = its roll_die
  GAME    block (round, game):  o0,o1 = the roll after this round's action, o2 = the random agent's choice
  EXPAND / ROLLOUT (search):    game id = game index, epoch = round
  DIRICHLET (Agent::Model):     epoch = 2 * round + (0 for player 1's batch, 1 for player 2's)
  SAMPLE  block (round, game):  the categorical draw of Agent::Model
"""
import ctypes
import ctypes.util
import enum

import numpy as np

from . import _ffi
from .mcts import MctsConfig

_libm = ctypes.CDLL(ctypes.util.find_library("m") or "libm.so.6")
_libm.powf.restype = ctypes.c_float
_libm.powf.argtypes = [ctypes.c_float, ctypes.c_float]


class Agent(enum.Enum):  # versus.rs:124-130 (clap ValueEnum)
    Model = "model"
    Mcts = "mcts"
    Random = "random"
    Nobody = "none"


class Player:  # versus.rs:124-127
    def __init__(self, player_type, model=None):
        self.player_type, self.model = player_type, model


_AGENT_SERDE = {Agent.Random: "Random", Agent.Mcts: "Mcts", Agent.Model: "Model", Agent.Nobody: "None"}  # serde variant names


class Game:
    """`Game<T>` (versus.rs:27-50) and its JSON wire format (serde derive): {"id", "player1", "player2", "turns",
    "winner", "initial_state"}; a Backgammon state serialises as {"board": [[24 x i8], [hit0, hit1], [col0, col1]],
    "roll": [a, b], "player": +-1, "is_second_play": bool, "id": n} (backgammon_logic.rs:53-60), a TicTacToe state as
    {"board": [9 x i8], "player": +-1, "id": n}.  The reference never fills `turns` (versus.rs:47 only creates it)
    and types `Turn::action` as Vec<T>; files written here keep `turns` empty so the reference's `load_game` reads
    them, unless the arena is asked to record turns (an extension: roll, action as [from, to] pairs, agent)."""

    def __init__(self, game_id, player1, player2, initial_state, winner=Agent.Nobody, turns=None):
        self.id, self.player1, self.player2 = game_id, player1, player2
        self.initial_state, self.winner, self.turns = initial_state, winner, turns or []

    def to_json(self):
        return {"id": self.id, "player1": _AGENT_SERDE[self.player1], "player2": _AGENT_SERDE[self.player2],
                "turns": self.turns, "winner": _AGENT_SERDE[self.winner], "initial_state": self.initial_state}

    @classmethod
    def from_json(cls, d):
        rev = {v: k for k, v in _AGENT_SERDE.items()}
        return cls(d["id"], rev[d["player1"]], rev[d["player2"]], d["initial_state"], rev[d["winner"]], d.get("turns", []))


def _bg_state_json(s, idx):
    return {"board": [[int(v) for v in s["pts"]], [int(s["bar"][0]), int(s["bar"][1])], [int(s["off"][0]), int(s["off"][1])]],
            "roll": [int(s["roll"][0]), int(s["roll"][1])], "player": int(s["player"]), "is_second_play": bool(s["second"]),
            "id": int(idx)}


def _game_id(seed, idx):
    """a 21-character id like nanoid's (versus.rs:45), but reproducible: drawn from the INIT stream of the game"""
    alphabet = "useandom-26T198340PX75pxJACKVERYMINDBUSHWOLF_GQZbfghjklqvwyzrict"  # nanoid's url alphabet
    out = []
    for blk in range(6):
        for w in _ffi.philox(seed, 1 + blk, idx, _ffi.STREAM_INIT, 0):
            out.append(alphabet[int(w) & 63])
    return "".join(out[:21])


def save_game(game, game_path):
    """versus.rs:52-61: <game_path>/<id>.json, pretty-printed"""
    import json
    import os
    path = os.path.join(game_path, f"{game.id}.json")
    with open(path, "w") as f:
        f.write(json.dumps(game.to_json(), indent=2))
    return path


def load_game(path):
    """versus.rs:63-71"""
    import json
    with open(path) as f:
        return Game.from_json(json.load(f))


def load_all_games(path):
    """versus.rs:107-122: every *.json file of the directory"""
    import os
    return [load_game(os.path.join(path, n)) for n in sorted(os.listdir(path)) if n.endswith(".json") and os.path.isfile(os.path.join(path, n))]


def print_game(path, out=print):
    """the `replay` subcommand (versus.rs:73-105, main.rs:208-213), without the interactive pause"""
    game = load_game(path)
    out(f"Game ID: {game.id}")
    out(f"Player 1: {_AGENT_SERDE[game.player1]}, Player 2: {_AGENT_SERDE[game.player2]}")
    out(f"Game winner: {_AGENT_SERDE[game.winner]}")
    out("Initial State:")
    out(str(game.initial_state))
    for turn in game.turns:
        out(f"Player: {turn.get('player')}")
        out(f"Roll: {turn.get('roll')}")
        out(f"Action: {turn.get('action')}")
    return game


class PlayResult:  # versus.rs:130-152
    games = ()  # Vec<Game<T>> in the order the games were retired (filled when the arena keeps games)

    def __init__(self, player1, player2, wins_p1, wins_p2, n_games, winners, rounds):
        self.player1, self.player2, self.wins_p1, self.wins_p2, self.n_games = player1, player2, wins_p1, wins_p2, n_games
        self.draws = n_games - (wins_p1 + wins_p2)
        self.winrate = wins_p1 / n_games if n_games else 0.0
        self.winners = winners   # per game: -1 / +1 / 0 (draw)   (Game::winner, versus.rs:236-246)
        self.rounds = rounds     # per game: the round in which it was retired

    def __str__(self):
        return (f"Player 1: {self.player1}\nPlayer 2: {self.player2}\nWins Player 1: {self.wins_p1}\n"
                f"Wins Player 2: {self.wins_p2}\nDraws: {self.draws}\nNumber of Games: {self.n_games}\n"
                f"Winrate: {self.winrate * 100.0}%\n")


# ---------------------------------------------------------------- backgammon
def _bg_finished_dummy():
    s = np.zeros(1, dtype=_ffi.BG_STATE)
    s["off"][0] = (15, 0)
    s["roll"][0] = (1, 2)
    s["player"] = -1
    return s


def _bg_initial(n_games, seed):
    s = np.zeros(n_games, dtype=_ffi.BG_STATE)
    # Backgammon::new (backgammon_logic.rs:80-94)
    start = [0] * 24
    for pt, v in ((23, -2), (12, -5), (7, -3), (5, -5), (0, 2), (11, 5), (16, 3), (18, 5)):
        start[pt] = v
    s["pts"][:] = start
    s["player"] = -1
    for g in range(n_games):
        o = _ffi.philox(seed, 0, g, _ffi.STREAM_INIT, 0)
        if g >= n_games // 2:
            s["player"][g] = 1                                   # skip_turn (:192-196); its roll is overwritten
            s["roll"][g] = (_ffi.die_of(o[2]), _ffi.die_of(o[3]))      # roll_die
        else:
            s["roll"][g] = (_ffi.die_of(o[0]), _ffi.die_of(o[1]))
    return s


def _weighted_select(ids, weights, seed, game, rnd):
    """AlphaZero::weighted_select_tensor_idx (alphazero.rs:129-137): f64 cumulative weights in action-id order"""
    dense = np.zeros(_ffi.ACTION_SPACE, dtype=np.float64)
    dense[ids] = weights.astype(np.float64)
    total = 0.0
    for j in range(_ffi.ACTION_SPACE):
        total += dense[j]
    o = _ffi.philox(seed, rnd, game, _ffi.STREAM_SAMPLE, 0)
    u = float(((int(o[0]) << 21) ^ (int(o[1]) >> 11)) & ((1 << 53) - 1)) * (1.0 / 9007199254740992.0)
    chosen, cum, last = u * total, 0.0, -1
    for j in range(_ffi.ACTION_SPACE):
        if dense[j] == 0.0:
            continue
        cum += dense[j]
        last = j
        if cum > chosen:
            return j
    return last


def _bg_actions(ctx, player, side, states, live_ids, all_states, cfg, temp, seed, rnd):
    """get_actions_for_player (versus.rs:270-318) for one side's games -> MOVE array"""
    n = len(live_ids)
    moves = np.full(n, _ffi.NONE, dtype=np.int8).repeat(4).view(_ffi.MOVE).reshape(n)
    if n == 0:
        return moves
    if player.player_type is Agent.Random:
        mv, cnt = ctx.bg_valid_moves(states)
        for i, g in enumerate(live_ids):
            if cnt[i] > 0:
                o = _ffi.philox(seed, rnd, int(g), _ffi.STREAM_GAME, 0)
                moves[i] = mv[i, _ffi.index_of(o[2], int(cnt[i]))]
        return moves
    if player.player_type is Agent.Mcts:
        # one launch for the side: a dense array indexed by game so that every search's streams are keyed by
        # its game index; games that do not take part are a finished dummy (their root is terminal: no work)
        dense = np.repeat(_bg_finished_dummy(), len(all_states))
        dense[live_ids] = states
        best, status, _ = ctx.mcts_search(_ffi.GAME_BACKGAMMON, dense, dense["player"].copy(), cfg.record(), seed, 0, rnd)
        for g in live_ids:
            if status[g] == _ffi.ERR_NO_MOVES_PANIC:
                raise _ffi.DieeError(int(status[g]), f"game {g}: expand() called on node with no expandable moves "
                                                     "(node.rs:119-121; use MODE_PASS_CHILD for arena play)")
            if status[g] != _ffi.OK:
                raise _ffi.DieeError(int(status[g]), f"game {g}: search failed")
        return best[live_ids]
    if player.player_type is Agent.Model:
        from .alphazero import alpha_mcts_parallel
        roots = alpha_mcts_parallel(states, player.model, cfg, seed, game_ids=np.asarray(live_ids, dtype=np.uint32),
                                    epoch=2 * rnd + side, ctx=ctx)
        tinv = ctypes.c_float(1.0 / temp)
        for i, g in enumerate(live_ids):
            k = int(roots.counts[i])
            if k == 0:
                continue  # no children: EMPTY_MOVE
            vis = roots.visits[i, :k]
            s = np.float32(0)
            for v in vis:
                s = np.float32(s + v)
            w = np.array([_libm.powf(ctypes.c_float(np.float32(v) / s), tinv) for v in vis], dtype=np.float32)
            if not (w.sum() != 0):
                continue
            a = _weighted_select(roots.ids[i, :k].astype(np.int64), w, seed, int(g), rnd)
            moves[i] = ctx.bg_decode_moves(states[i:i + 1], np.array([a], dtype=np.uint16))[0]
        return moves
    raise ValueError("Agent::None cannot play (versus.rs:316)")


def play_backgammon_device(player1, player2, mcts_config=None, seed=0xD1EE, num_games=400, round_limit=400, ctx=None):
    """the arena with the games resident on the device (csrc/arena.cu): per round a handful of launches and one
    20-byte read-back; Agent::Mcts and Agent::Random.  Same results as play_backgammon(device_resident=False)."""
    ctx = ctx or _ffi.default_context()
    cfg = mcts_config or MctsConfig()
    kinds = {Agent.Random: _ffi.AGENT_RANDOM, Agent.Mcts: _ffi.AGENT_MCTS}
    a1, a2 = kinds[player1.player_type], kinds[player2.player_type]
    arena = _ffi.Arena(ctx, num_games, seed, round_limit)
    retired = 0
    while retired < num_games:
        retired, wins_p1, wins_p2, _ = arena.round(a1, a2, cfg.record())
    states, winners, rounds = arena.read()
    arena.close()
    res = PlayResult(player1.player_type, player2.player_type, int(wins_p1), int(wins_p2), num_games, winners, rounds)
    res.final_states = states
    return res


def play_backgammon(player1, player2, mcts_config=None, temp=1.0, seed=0xD1EE, num_games=400, round_limit=400, ctx=None,
                    keep_games=False, record_turns=False, device_resident=None):
    """device_resident: None = on the device whenever both agents can (Mcts / Random) and no per-game files are kept"""
    can = (player1.player_type in (Agent.Random, Agent.Mcts) and player2.player_type in (Agent.Random, Agent.Mcts)
           and not keep_games and not record_turns)
    if device_resident is None:
        device_resident = can
    if device_resident:
        if not can:
            raise ValueError("the device-resident arena plays Agent::Mcts / Agent::Random and keeps no per-game files")
        return play_backgammon_device(player1, player2, mcts_config, seed, num_games, round_limit, ctx)
    ctx = ctx or _ffi.default_context()
    cfg = mcts_config or MctsConfig()
    states = _bg_initial(num_games, seed)
    games = [Game(_game_id(seed, g), player1.player_type, player2.player_type, _bg_state_json(states[g], g))
             for g in range(num_games)] if keep_games else None
    retired = []
    live = np.ones(num_games, dtype=bool)
    winners = np.zeros(num_games, dtype=np.int8)
    rounds = np.zeros(num_games, dtype=np.int32)
    wins_p1 = wins_p2 = 0
    round_count = 0
    empty = np.full(4, _ffi.NONE, dtype=np.int8).view(_ffi.MOVE)[0]
    while live.any():
        ids = np.nonzero(live)[0]
        ids_p1 = ids[states["player"][ids] == -1]
        ids_p2 = ids[states["player"][ids] != -1]
        a1 = _bg_actions(ctx, player1, 0, states[ids_p1], ids_p1, states, cfg, temp, seed, round_count)
        a2 = _bg_actions(ctx, player2, 1, states[ids_p2], ids_p2, states, cfg, temp, seed, round_count)
        order = np.concatenate([ids_p1, ids_p2])
        acts = np.concatenate([a1, a2])
        rolls = np.zeros((len(order), 2), dtype=np.uint8)
        for i, g in enumerate(order):
            o = _ffi.philox(seed, round_count, int(g), _ffi.STREAM_GAME, 0)
            rolls[i] = (_ffi.die_of(o[0]), _ffi.die_of(o[1]))
        if keep_games and record_turns:
            from .backgammon import _move_list
            for i, g in enumerate(order):
                mover = player1 if states["player"][g] == -1 else player2
                games[g].turns.append({"roll": [int(states["roll"][g][0]), int(states["roll"][g][1])],
                                       "action": [list(p) for p in _move_list(acts[i])], "player": _AGENT_SERDE[mover.player_type]})
        states[order] = ctx.bg_apply_moves(states[order], acts, rolls)  # EMPTY_MOVE -> skip_turn
        round_count += 1
        for i, g in enumerate(order):
            if acts[i] == empty:
                continue  # versus.rs:222-225: no winner / round-limit test after a skipped turn
            off = states["off"][g]
            w = -1 if off[0] == 15 else (1 if off[1] == 15 else None)
            if w is None and round_count >= round_limit:
                w = 0
            if w is not None:
                live[g] = False
                winners[g], rounds[g] = w, round_count
                wins_p1 += w == -1
                wins_p2 += w == 1
                if keep_games:  # versus.rs:236-246
                    games[g].winner = player1.player_type if w == -1 else (player2.player_type if w == 1 else Agent.Nobody)
                    retired.append(games[g])
    res = PlayResult(player1.player_type, player2.player_type, int(wins_p1), int(wins_p2), num_games, winners, rounds)
    if keep_games:
        res.games = retired
    return res


# ---------------------------------------------------------------- tictactoe (BASELINE configs[0])
def _ttt_winner(board):
    from .tictactoe import TicTacToe
    for a, b, c in TicTacToe._LINES:
        if board[a] != 0 and board[a] == board[b] == board[c]:
            return int(board[a])
    return 0 if all(v != 0 for v in board) else None


def play_tictactoe(player1, player2, mcts_config=None, temp=1.0, seed=0xD1EE, num_games=400, round_limit=400, ctx=None):
    ctx = ctx or _ffi.default_context()
    cfg = mcts_config or MctsConfig()
    states = np.zeros(num_games, dtype=_ffi.TTT_STATE)
    states["player"] = -1
    states["player"][num_games // 2:] = 1  # skip_turn for the second half (versus.rs:172-174)
    dummy = np.zeros(1, dtype=_ffi.TTT_STATE)
    dummy["board"][0] = (-1, -1, -1, 0, 0, 0, 0, 0, 0)
    dummy["player"] = 1
    live = np.ones(num_games, dtype=bool)
    winners = np.zeros(num_games, dtype=np.int8)
    rounds = np.zeros(num_games, dtype=np.int32)
    wins_p1 = wins_p2 = 0
    round_count = 0
    EMPTY = 10

    def actions(player, live_ids):
        out = np.full(len(live_ids), EMPTY, dtype=np.int64)
        if len(live_ids) == 0:
            return out
        if player.player_type is Agent.Random:
            for i, g in enumerate(live_ids):
                cells = [c for c in range(9) if states["board"][g][c] == 0]
                if cells:
                    o = _ffi.philox(seed, round_count, int(g), _ffi.STREAM_GAME, 0)
                    out[i] = cells[_ffi.index_of(o[2], len(cells))]
            return out
        if player.player_type is Agent.Mcts:
            dense = np.repeat(dummy, num_games)
            dense[live_ids] = states[live_ids]
            best, status, _ = ctx.mcts_search(_ffi.GAME_TICTACTOE, dense, dense["player"].copy(), cfg.record(), seed, 0, round_count)
            for g in live_ids:
                if status[g] != _ffi.OK:
                    raise _ffi.DieeError(int(status[g]), f"game {g}: search failed")
            return np.asarray(best, dtype=np.int64)[live_ids]
        raise ValueError("only Agent::Mcts and Agent::Random are built for TicTacToe (no 3x3 net geometry)")

    while live.any():
        ids = np.nonzero(live)[0]
        ids_p1 = ids[states["player"][ids] == -1]
        ids_p2 = ids[states["player"][ids] != -1]
        a1, a2 = actions(player1, ids_p1), actions(player2, ids_p2)
        round_count += 1
        for g, a in list(zip(ids_p1, a1)) + list(zip(ids_p2, a2)):
            if a == EMPTY:
                states["player"][g] = -states["player"][g]
                continue
            states["board"][g][a] = states["player"][g]     # apply_move (tictactoe/mod.rs:46-49)
            states["player"][g] = -states["player"][g]
            w = _ttt_winner(states["board"][g])
            if w is None and round_count >= round_limit:
                w = 0
            if w is not None:
                live[g] = False
                winners[g], rounds[g] = w, round_count
                wins_p1 += w == -1
                wins_p2 += w == 1
    return PlayResult(player1.player_type, player2.player_type, int(wins_p1), int(wins_p2), num_games, winners, rounds)


def play(game, player1, player2, mcts_config=None, temp=1.0, **kw):
    """`pub fn play<T: LearnableGame>(player1, player2, mcts_config, temp) -> PlayResult<T>` (versus.rs:160);
    `game` stands for the type parameter: "backgammon" | "tictactoe" (or the host classes)."""
    name = game if isinstance(game, str) else game.name()
    if name == "backgammon":
        return play_backgammon(player1, player2, mcts_config, temp, **kw)
    if name in ("tictactoe", "tic-tac-toe"):
        return play_tictactoe(player1, player2, mcts_config, temp, **kw)
    raise ValueError(f"unknown game {name!r}")
