"""`TicTacToe`: the reference's second LearnableGame (src/tictactoe/mod.rs:15-117).  The state is
nine cells; the env itself is host bookkeeping, the search over it runs on the GPU (mcts.py)."""
import numpy as np

from . import _ffi


class TicTacToe:
    EMPTY_MOVE = 10                 # tictactoe/mod.rs:18
    IS_DETERMINISTIC = True
    ACTION_SPACE_SIZE = 9
    CONV_OUTPUT_SIZE = 9
    N_INPUT_CHANNELS = 3
    N_FILTERS = 64
    N_RES_BLOCKS = 4
    _LINES = ((0, 1, 2), (3, 4, 5), (6, 7, 8), (0, 3, 6), (1, 4, 7), (2, 5, 8), (0, 4, 8), (2, 4, 6))

    def __init__(self):
        self.s = np.zeros(1, dtype=_ffi.TTT_STATE)
        self.s["player"] = -1
        self.id = 0

    @classmethod
    def new(cls):
        return cls()

    @staticmethod
    def name():
        return "tictactoe"

    @property
    def board(self):
        return [int(x) for x in self.s["board"][0]]

    @board.setter
    def board(self, b):
        self.s["board"][0] = b

    def copy(self):
        g = TicTacToe()
        g.s = self.s.copy()
        g.id = self.id
        return g

    def get_valid_moves(self):  # :36-44
        return [i for i, v in enumerate(self.board) if v == 0]

    def apply_move(self, action):  # :46-49
        self.s["board"][0][action] = self.s["player"][0]
        self.s["player"] = -self.s["player"]

    def roll_die(self):  # base.rs:32-37
        raise RuntimeError("roll_die called on deterministic game!")

    def skip_turn(self):  # :51-53
        self.s["player"] = -self.s["player"]

    def get_player(self):
        return int(self.s["player"][0])

    def check_winner(self):  # :59-79
        b = self.board
        for a, c, d in self._LINES:
            if b[a] != 0 and b[a] == b[c] == b[d]:
                return b[a]
        return 0 if all(v != 0 for v in b) else None

    def as_tensor(self):  # :81-92 -> [1,3,3,3]
        b = np.array(self.board).reshape(3, 3)
        return np.stack([b == -1, b == 0, b == 1]).astype(np.float32)[None]

    def decode(self, action):
        return int(action)

    def encode(self, action):
        return int(action)

    def get_id(self):
        return self.id

    def set_id(self, new_id):
        self.id = new_id

    def to_pretty_str(self):  # :108-116
        b = self.board
        return "\n".join("|".join(str(v) for v in b[r * 3:r * 3 + 3]) for r in range(3))
