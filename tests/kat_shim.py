"""Shim namespace that the transcribed reference tests (tests/golden/ref_*_kats.py) run against.

`make_shim(backend)` binds `Backgammon::...` / `TicTacToe::...` to a backend object:
tests/orc.py-backed OracleBackend (CPU oracle) or the CUDA path's backend (tests/test_gpu_*.py).
A backend may raise Unsupported for tree-level helpers it does not implement.
"""
import numpy as np


class Unsupported(Exception):
    pass


class RList(list):
    """Rust Vec / tuple stand-in (mutable, with the handful of methods the tests use)."""

    def is_empty(self):
        return len(self) == 0

    def len(self):
        return len(self)

    def contains(self, x):
        return any(norm(x) == norm(y) for y in self)

    def clone(self):
        return v(norm(self))

    def eq(self, other):
        return norm(self) == norm(other)

    def sum_i8(self):
        return sum(self)


class Node:
    def __init__(self, value, children):
        self.value = tuple(value)
        self.children = RList(children)

    def __eq__(self, other):
        return norm(self) == norm(other)

    def __repr__(self):
        return f"Node({self.value}, {list(self.children)})"


class Opt:
    def __init__(self, x):
        self.x = x

    def is_none(self):
        return self.x is None

    def unwrap(self):
        assert self.x is not None
        return self.x


def norm(x):
    if isinstance(x, Node):
        return ("node", tuple(x.value), norm(x.children))
    if isinstance(x, (list, tuple)):
        return [norm(y) for y in x]
    if isinstance(x, np.generic):
        return x.item()
    return x


def v(x):
    if isinstance(x, (list, tuple)) and not isinstance(x, Node):
        return RList(v(y) for y in x)
    return x


def assert_eq(a, b):
    assert norm(a) == norm(b), f"{norm(a)} != {norm(b)}"


def assert_(c):
    assert c


class _BgObj:
    def __init__(self, backend, board, player, second):
        self._b = backend
        self.board = v(board)
        self.roll = RList([0, 0])
        self.player = player
        self.is_second_play = second

    def get_valid_moves(self):
        return v(self._b.valid_moves(norm(self.board), norm(self.roll), self.player, self.is_second_play))


class _Backgammon:
    def __init__(self, backend):
        self._b = backend

    def get_initial_state(self):
        return v(self._b.initial_board())

    def new(self):
        return _BgObj(self._b, self._b.initial_board(), -1, False)

    def init_with_fields(self, board, player, second):
        return _BgObj(self._b, norm(board), player, second)

    def get_next_state(self, state, actions, player):
        return v(self._b.next_state(norm(state), norm(actions), player))

    def get_normal_moves(self, dice, state, player):
        return v([_tree(t) for t in self._b.normal_moves(norm(dice), norm(state), player)])

    def get_entry_moves(self, dice, state, player):
        return v([_tree(t) for t in self._b.entry_moves(norm(dice), norm(state), player)])

    def is_collectible(self, state, player):
        return self._b.is_collectible(norm(state), player)

    def check_win(self, state, player):
        return self._b.check_win(norm(state), player)

    def extract_sequences_list(self, trees):
        return v(self._b.extract_sequences_list([_untree(t) for t in trees]))

    def extract_sequences_node(self, tree):
        return v(self._b.extract_sequences_node(_untree(tree)))

    def remove_duplicate_states(self, state, sequences, player):
        return v(self._b.remove_duplicate_states(norm(state), norm(sequences), player))


def _tree(t):
    return Node(t[0], [_tree(c) for c in t[1]])


def _untree(n):
    return (tuple(n.value), [_untree(c) for c in n.children])


class _TttObj:
    def __init__(self, backend):
        self._b = backend
        self.board = RList([0] * 9)
        self.player = -1

    def get_player(self):
        return self.player

    def apply_move(self, m):
        board, player = self._b.ttt_apply(norm(self.board), self.player, m)
        self.board = v(board)
        self.player = player

    def get_valid_moves(self):
        return v(self._b.ttt_valid_moves(norm(self.board), self.player))

    def check_winner(self):
        return Opt(self._b.ttt_check_winner(norm(self.board), self.player))


class _TicTacToe:
    def __init__(self, backend):
        self._b = backend

    def new(self):
        return _TttObj(self._b)


class Shim:
    Node = Node
    v = staticmethod(v)
    assert_eq = staticmethod(assert_eq)
    assert_ = staticmethod(assert_)

    def __init__(self, backend):
        self.Backgammon = _Backgammon(backend)
        self.TicTacToe = _TicTacToe(backend)


def make_shim(backend):
    return Shim(backend)


class OracleBackend:
    """binds the shim to the CPU oracle (tests/orc.py)"""

    def __init__(self):
        import orc
        self.o = orc

    def initial_board(self):
        s = self.o.bg_new()
        return ([int(x) for x in s["pts"][0]], tuple(int(x) for x in s["bar"][0]), tuple(int(x) for x in s["off"][0]))

    def _board(self, b):
        return self.o.make_board(b[0], b[1], b[2])

    def valid_moves(self, board, roll, player, second):
        s = self.o.make_state(board[0], board[1], board[2], roll, player, second)
        return self.o.bg_valid_moves(s)

    def next_state(self, board, actions, player):
        return self.o.board_tuple(self.o.bg_next_state(self._board(board), actions, player))

    def normal_moves(self, dice, board, player):
        return self.o.bg_normal_moves(dice, self._board(board), player)[0]

    def entry_moves(self, dice, board, player):
        return self.o.bg_entry_moves(dice, self._board(board), player)[0]

    def is_collectible(self, board, player):
        return self.o.bg_is_collectible(self._board(board), player)

    def check_win(self, board, player):
        return self.o.bg_check_win(self._board(board), player)

    def _pool(self, trees):
        o = self.o
        nodes = []

        def alloc(sibs):
            first = len(nodes)
            for t in sibs:
                nodes.append([t[0][0], t[0][1], -1, 0])
            for i, t in enumerate(sibs):
                if t[1]:
                    fc = alloc(t[1])
                    nodes[first + i][2] = fc
                    nodes[first + i][3] = len(t[1])
            return first

        alloc(trees)
        pool = np.zeros(max(1, len(nodes)), dtype=o.ANODE)
        for i, (f, t, fc, nc) in enumerate(nodes):
            pool[i] = (f, t, 0, 0, fc, nc)
        return pool

    def extract_sequences_list(self, trees):
        return self.o.bg_extract_sequences_list(self._pool(trees), len(trees))

    def extract_sequences_node(self, tree):
        return self.o.bg_extract_sequences_node(self._pool([tree]), 0)

    def remove_duplicate_states(self, board, sequences, player):
        return self.o.bg_remove_duplicate_states(self._board(board), sequences, player)

    def ttt_apply(self, board, player, m):
        s = self.o.ttt_new()
        s["board"][0] = board
        s["player"][0] = player
        self.o.ttt_apply_move(s, m)
        return [int(x) for x in s["board"][0]], int(s["player"][0])

    def ttt_valid_moves(self, board, player):
        s = self.o.ttt_new()
        s["board"][0] = board
        s["player"][0] = player
        return self.o.ttt_valid_moves(s)

    def ttt_check_winner(self, board, player):
        s = self.o.ttt_new()
        s["board"][0] = board
        s["player"][0] = player
        return self.o.ttt_check_winner(s)
