"""`Backgammon`: the reference's LearnableGame impl (src/backgammon/backgammon_logic.rs:54-415) as a
thin host object over the CUDA env kernels.  Same method names, argument meaning and error
behaviour (the reference's asserts/panics surface as AssertionError / DieeError)."""
import numpy as np

from . import _ffi


def _move_rec(actions):
    m = np.zeros(1, dtype=_ffi.MOVE)
    m[0] = (_ffi.NONE,) * 4
    assert len(actions) <= 2, "encoding for actions > 2 is not implemented!"  # backgammon_logic.rs:263
    if len(actions) > 0:
        m["from1"], m["to1"] = actions[0]
    if len(actions) > 1:
        m["from2"], m["to2"] = actions[1]
    return m


def _move_list(rec):
    out = []
    if rec["from1"] != _ffi.NONE:
        out.append((int(rec["from1"]), int(rec["to1"])))
    if rec["from2"] != _ffi.NONE:
        out.append((int(rec["from2"]), int(rec["to2"])))
    return out


class Backgammon:
    """trait LearnableGame for Backgammon (base.rs:8-51)."""

    EMPTY_MOVE = []                 # backgammon_logic.rs:72
    IS_DETERMINISTIC = False
    ACTION_SPACE_SIZE = 1352        # :74
    N_INPUT_CHANNELS = 6
    CONV_OUTPUT_SIZE = 24
    N_FILTERS = 256
    N_RES_BLOCKS = 19

    def __init__(self, ctx=None, seed=0, game_id=0):
        # Backgammon::new :80-94
        self.s = np.zeros(1, dtype=_ffi.BG_STATE)
        self.s["pts"][0] = [2, 0, 0, 0, 0, -5, 0, -3, 0, 0, 0, 5, -5, 0, 0, 0, 3, 0, 5, 0, 0, 0, 0, -2]
        self.s["player"] = -1
        self.id = 0
        self._ctx = ctx
        # injected stream position (include/diee.h): the reference draws from thread_rng()
        self.seed, self.game_id, self._rolls = seed, game_id, 0

    @classmethod
    def new(cls, ctx=None, seed=0, game_id=0):
        return cls(ctx, seed, game_id)

    @classmethod
    def init_with_fields(cls, board, player, is_second_play, ctx=None):  # :419-427
        g = cls(ctx)
        g.board = board
        g.s["player"] = player
        g.s["second"] = 1 if is_second_play else 0
        return g

    @staticmethod
    def name():
        return "backgammon"

    @property
    def ctx(self):
        if self._ctx is None:
            self._ctx = _ffi.default_context()
        return self._ctx

    # ---- fields ----
    @property
    def board(self):
        r = self.s[0]
        return ([int(x) for x in r["pts"]], tuple(int(x) for x in r["bar"]), tuple(int(x) for x in r["off"]))

    @board.setter
    def board(self, b):
        self.s["pts"][0] = b[0]
        self.s["bar"][0] = b[1]
        self.s["off"][0] = b[2]

    @property
    def roll(self):
        return tuple(int(x) for x in self.s["roll"][0])

    @roll.setter
    def roll(self, r):
        self.s["roll"][0] = r

    @property
    def player(self):
        return int(self.s["player"][0])

    @property
    def is_second_play(self):
        return bool(self.s["second"][0])

    def copy(self):
        g = Backgammon(self._ctx, self.seed, self.game_id)
        g.s = self.s.copy()
        g.id, g._rolls = self.id, self._rolls
        return g

    # ---- LearnableGame ----
    def _next_dice(self):
        stream, idx = (_ffi.STREAM_INIT, 0) if self._rolls == 0 else (_ffi.STREAM_GAME, self._rolls - 1)
        w = _ffi.philox(self.seed, idx, self.game_id, stream, 0)
        self._rolls += 1
        return _ffi.die_of(w[0]), _ffi.die_of(w[1])

    def roll_die(self):  # :100-104
        self.roll = self._next_dice()
        return self.roll

    def get_valid_moves(self):  # :403-414
        assert self.roll != (0, 0), "die has not been rolled!"
        moves, counts = self.ctx.bg_valid_moves(self.s)
        if counts[0] < 0:
            raise _ffi.DieeError(int(counts[0]), "get_valid_moves failed")
        return [_move_list(moves[0, k]) for k in range(int(counts[0]))]

    def apply_move(self, actions):  # :176-186
        d = self._next_dice() if not (self.roll[0] == self.roll[1] and not self.is_second_play) else (0, 0)
        self.s = self.ctx.bg_apply_moves(self.s, _move_rec(actions), np.array(d, dtype=np.uint8))

    def skip_turn(self):  # :192-196
        self.s = self.ctx.bg_apply_moves(self.s, _move_rec([]), np.array(self._next_dice(), dtype=np.uint8))

    def get_player(self):
        return self.player

    def check_winner(self):  # :106-108, :527-534
        off = self.s["off"][0]
        return -1 if off[0] == 15 else (1 if off[1] == 15 else None)

    def as_tensor(self):  # :198-252 -> numpy f32 [1,6,4,6]
        assert self.roll != (0, 0), "die has not been rolled!"
        return self.ctx.bg_encode_states(self.s).reshape(1, 6, 4, 6)

    def encode(self, actions):  # :262-359
        return int(self.ctx.bg_encode_moves(self.s, _move_rec(actions))[0])

    def decode(self, action):  # :361-401
        return _move_list(self.ctx.bg_decode_moves(self.s, np.array([action], dtype=np.uint16))[0])

    def get_id(self):
        return self.id

    def set_id(self, new_id):
        self.id = new_id

    def to_pretty_str(self):  # :110-174 (layout only; not on the hot path)
        pts, bar, off = self.board
        who = "Player 1" if self.player == -1 else "Player 2"
        top = " ".join(f"{v:3d}" for v in pts[12:])
        bot = " ".join(f"{v:3d}" for v in reversed(pts[:12]))
        return (f"Current turn: {who}\tRoll: {self.roll}\nPlayer 1:\n\tBroken Pieces: {bar[0]}\n\tPieces Collected:{off[0]}\n"
                f"Player 2:\n\tBroken Pieces: {bar[1]}\n\tPieces Collected:{off[1]}\n{'=' * 60}\n{top}\n{bot}\n{'=' * 60}")
