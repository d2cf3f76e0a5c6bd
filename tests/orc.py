"""ctypes binding of the CPU oracle (oracle/liborc.so) for tests, smoke() and bench's CPU legs.

The oracle is test infrastructure; the product package (die_e_b200) never imports this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_DIR = os.path.join(os.path.dirname(_HERE), "oracle")
NONE = -2
NO_WINNER = 2
MAX_MOVES = 512

BG_STATE = np.dtype([("pts", "i1", (24,)), ("bar", "u1", (2,)), ("off", "u1", (2,)),
                     ("roll", "u1", (2,)), ("player", "i1"), ("second", "u1")])
assert BG_STATE.itemsize == 32
BOARD = np.dtype([("pts", "i1", (24,)), ("bar", "u1", (2,)), ("off", "u1", (2,))])
MOVE = np.dtype([("from1", "i1"), ("to1", "i1"), ("from2", "i1"), ("to2", "i1")])
TTT_STATE = np.dtype([("board", "i1", (9,)), ("player", "i1"), ("pad", "u1", (6,))])
assert TTT_STATE.itemsize == 16
SEQ = np.dtype([("from", "i1", (4,)), ("to", "i1", (4,)), ("len", "i4")])
ANODE = np.dtype([("from", "i1"), ("to", "i1"), ("die", "i1"), ("pad", "i1"),
                  ("first_child", "i4"), ("n_children", "i4")])
NODE_STATS = np.dtype([("parent", "i4"), ("visits", "f4"), ("value", "f4"), ("action", MOVE),
                       ("n_moves", "i4"), ("n_untried", "i4")])
MCTS_CFG = np.dtype([("iterations", "u4"), ("c", "f4"), ("simulate_round_limit", "u4"),
                     ("dirichlet_alpha", "f4"), ("dirichlet_epsilon", "f4"), ("mode_flags", "u4")])

MODE_ROLLOUT_CHECK_CURRENT = 1
MODE_PASS_CHILD = 2
ERR_NO_MOVES_PANIC = -3

STREAM_INIT, STREAM_GAME, STREAM_ROLLOUT, STREAM_EXPAND, STREAM_DIRICHLET, STREAM_SAMPLE = range(6)


def build():
    subprocess.run(["make", "-s", "-C", ORACLE_DIR], check=True)


_lib = None


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(ORACLE_DIR, "liborc.so")
        if not os.path.exists(path):
            build()
        _lib = C.CDLL(path)
        _lib.orc_ln_f32.restype = C.c_float
        _lib.orc_ln_f32.argtypes = [C.c_float]
        _lib.orc_bg_encode.restype = C.c_uint32
        _lib.orc_bg_decode.restype = C.c_uint32  # 4-byte struct returned in a register
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def philox(seed, c0, c1, c2, c3):
    out = np.zeros(4, dtype=np.uint32)
    lib().orc_philox(C.c_uint64(seed), C.c_uint32(c0), C.c_uint32(c1), C.c_uint32(c2), C.c_uint32(c3), _p(out))
    return out


def die(w):
    return 1 + ((int(w) * 6) >> 32)


def index(w, n):
    return (int(w) * n) >> 32


def bg_new():
    s = np.zeros(1, dtype=BG_STATE)
    lib().orc_bg_new(_p(s))
    return s


def make_board(pts, bar=(0, 0), off=(0, 0)):
    b = np.zeros(1, dtype=BOARD)
    b["pts"][0] = pts
    b["bar"][0] = bar
    b["off"][0] = off
    return b


def make_state(pts, bar=(0, 0), off=(0, 0), roll=(0, 0), player=-1, second=False):
    s = np.zeros(1, dtype=BG_STATE)
    s["pts"][0] = pts
    s["bar"][0] = bar
    s["off"][0] = off
    s["roll"][0] = roll
    s["player"][0] = player
    s["second"][0] = 1 if second else 0
    return s


def board_tuple(b):
    b = b.reshape(-1)[0]
    return (list(int(x) for x in b["pts"]), tuple(int(x) for x in b["bar"]), tuple(int(x) for x in b["off"]))


def bg_next_state(board, actions, player):
    b = board.copy()
    f = np.array([a[0] for a in actions], dtype=np.int8)
    t = np.array([a[1] for a in actions], dtype=np.int8)
    lib().orc_bg_next_state(_p(b), _p(f), _p(t), C.c_int(len(actions)), C.c_int(player))
    return b


def bg_is_collectible(board, player):
    return bool(lib().orc_bg_is_collectible(_p(board), C.c_int(player)))


def bg_check_win(board, player):
    return bool(lib().orc_bg_check_win(_p(board), C.c_int(player)))


def _trees(fn, dice, board, player):
    pool = np.zeros(8192, dtype=ANODE)
    n_pool = C.c_int(0)
    d = np.array(list(dice), dtype=np.uint8)
    n_roots = fn(_p(d), C.c_int(len(d)), _p(board), C.c_int(player), _p(pool), C.c_int(len(pool)), C.byref(n_pool))
    assert n_roots >= 0
    return pool[: n_pool.value], n_roots


def _to_tree(pool, i):
    nd = pool[i]
    ch = [_to_tree(pool, int(nd["first_child"]) + k) for k in range(int(nd["n_children"]))]
    return ((int(nd["from"]), int(nd["to"])), ch)


def bg_normal_moves(dice, board, player):
    pool, n = _trees(lib().orc_bg_normal_moves, dice, board, player)
    return [_to_tree(pool, i) for i in range(n)], pool, n


def bg_entry_moves(dice, board, player):
    pool, n = _trees(lib().orc_bg_entry_moves, dice, board, player)
    return [_to_tree(pool, i) for i in range(n)], pool, n


def _seqs_to_list(seqs, n):
    return [[(int(seqs[i]["from"][k]), int(seqs[i]["to"][k])) for k in range(int(seqs[i]["len"]))] for i in range(n)]


def _list_to_seqs(lst):
    seqs = np.zeros(max(1, len(lst)), dtype=SEQ)
    for i, s in enumerate(lst):
        seqs[i]["len"] = len(s)
        for k, (f, t) in enumerate(s):
            seqs[i]["from"][k] = f
            seqs[i]["to"][k] = t
    return seqs


def bg_extract_sequences_list(pool, n_roots):
    out = np.zeros(4096, dtype=SEQ)
    n = lib().orc_bg_extract_sequences_list(_p(pool), C.c_int(n_roots), _p(out), C.c_int(len(out)))
    assert n >= 0
    return _seqs_to_list(out, n)


def bg_extract_sequences_node(pool, node):
    out = np.zeros(4096, dtype=SEQ)
    n = lib().orc_bg_extract_sequences_node(_p(pool), C.c_int(node), _p(out), C.c_int(len(out)))
    assert n >= 0
    return _seqs_to_list(out, n)


def bg_remove_duplicate_states(board, sequences, player):
    seqs = _list_to_seqs(sequences)
    out = np.zeros(max(1, len(sequences)), dtype=SEQ)
    n = lib().orc_bg_remove_duplicate_states(_p(board), _p(seqs), C.c_int(len(sequences)), C.c_int(player), _p(out))
    return _seqs_to_list(out, n)


def moves_to_list(mv, n):
    out = []
    for i in range(n):
        m = mv[i]
        s = []
        if m["from1"] != NONE:
            s.append((int(m["from1"]), int(m["to1"])))
        if m["from2"] != NONE:
            s.append((int(m["from2"]), int(m["to2"])))
        out.append(s)
    return out


def list_to_move(seq):
    m = np.zeros(1, dtype=MOVE)
    m[0] = (NONE, NONE, NONE, NONE)
    if len(seq) > 0:
        m["from1"], m["to1"] = seq[0]
    if len(seq) > 1:
        m["from2"], m["to2"] = seq[1]
    return m


def bg_valid_moves_raw(state):
    """-> (MOVE array, n) for one state (shape-(1,) BG_STATE array or a record)"""
    s = np.ascontiguousarray(state).reshape(-1)[:1]
    mv = np.zeros(MAX_MOVES, dtype=MOVE)
    n = lib().orc_bg_valid_moves(_p(s), _p(mv), C.c_int(MAX_MOVES))
    if n == -2:
        raise AssertionError("die has not been rolled!")
    assert n >= 0
    return mv, n


def bg_valid_moves(state):
    mv, n = bg_valid_moves_raw(state)
    return moves_to_list(mv, n)


def _move_u32(m):
    return C.c_uint32(int(np.ascontiguousarray(m).reshape(-1)[:1].view(np.uint32)[0]))


def bg_apply_move(state, move, d0, d1):
    lib().orc_bg_apply_move(_p(state), _move_u32(move), C.c_uint8(d0), C.c_uint8(d1))


def bg_skip_turn(state, d0, d1):
    lib().orc_bg_skip_turn(_p(state), C.c_uint8(d0), C.c_uint8(d1))


def bg_check_winner(state):
    w = lib().orc_bg_check_winner(_p(np.ascontiguousarray(state).reshape(-1)[:1]))
    return None if w == NO_WINNER else w


def bg_encode(state, seq):
    return int(lib().orc_bg_encode(_p(state), _move_u32(list_to_move(seq))))


def bg_decode(state, action):
    r = lib().orc_bg_decode(_p(state), C.c_uint32(action))
    m = np.array([r], dtype=np.uint32).view(MOVE)
    return moves_to_list(m, 1)[0]


def bg_as_tensor(state):
    out = np.zeros(144, dtype=np.float32)
    rc = lib().orc_bg_as_tensor(_p(state), _p(out))
    if rc != 0:
        raise AssertionError("die has not been rolled!")
    return out.reshape(1, 6, 4, 6)


def bg_playout(state, seed, game_id, round_limit):
    s = state.copy()
    plies = C.c_int32(0)
    w = lib().orc_bg_playout(_p(s), C.c_uint64(seed), C.c_uint32(game_id), C.c_int(round_limit), C.byref(plies))
    return w, plies.value, s


def bg_random_ply(state, blk):
    b = np.asarray(blk, dtype=np.uint32)
    return lib().orc_bg_random_ply(_p(state), _p(b))


def mcts_cfg(iterations=100, c=2.0, limit=400, alpha=0.3, eps=0.25, mode=0):
    cfg = np.zeros(1, dtype=MCTS_CFG)
    cfg[0] = (iterations, c, limit, alpha, eps, mode)
    return cfg


def mcts_search_bg(state, player, cfg, seed, game_id, epoch, want_finals=False):
    it = int(cfg["iterations"][0])
    nodes = np.zeros(it + 1, dtype=NODE_STATS)
    states = np.zeros(it + 1, dtype=BG_STATE)
    finals = np.zeros(it, dtype=BG_STATE)
    best = np.zeros(1, dtype=MOVE)
    n = C.c_int32(0)
    rc = lib().orc_mcts_search_bg_ex(_p(np.ascontiguousarray(state).reshape(-1)[:1]), C.c_int(player), _p(cfg), C.c_uint64(seed),
                                     C.c_uint32(game_id), C.c_uint32(epoch), _p(best), _p(nodes), _p(states), C.byref(n),
                                     _p(finals))
    if want_finals:
        return rc, best, nodes[: n.value], states[: n.value], finals
    return rc, best, nodes[: n.value], states[: n.value]


def ttt_new():
    s = np.zeros(1, dtype=TTT_STATE)
    s["player"] = -1
    return s


def ttt_valid_moves(state):
    out = np.zeros(9, dtype=np.uint8)
    n = lib().orc_ttt_valid_moves(_p(state), _p(out))
    return [int(x) for x in out[:n]]


def ttt_apply_move(state, m):
    lib().orc_ttt_apply_move(_p(state), C.c_uint8(m))


def ttt_check_winner(state):
    w = lib().orc_ttt_check_winner(_p(state))
    return None if w == NO_WINNER else w


def ttt_as_tensor(state):
    out = np.zeros(27, dtype=np.float32)
    lib().orc_ttt_as_tensor(_p(state), _p(out))
    return out.reshape(1, 3, 3, 3)


def mcts_search_ttt(state, player, cfg, seed, game_id, epoch):
    it = int(cfg["iterations"][0])
    nodes = np.zeros(it + 1, dtype=NODE_STATS)
    states = np.zeros(it + 1, dtype=TTT_STATE)
    best = C.c_uint8(0)
    n = C.c_int32(0)
    rc = lib().orc_mcts_search_ttt(_p(np.ascontiguousarray(state).reshape(-1)[:1]), C.c_int(player), _p(cfg), C.c_uint64(seed),
                                   C.c_uint32(game_id), C.c_uint32(epoch), C.byref(best), _p(nodes), _p(states), C.byref(n))
    return rc, best.value, nodes[: n.value], states[: n.value]


def mcts_search_bg_batch(states, players, cfg, seed, first_game_id, epoch, nthreads):
    states = np.ascontiguousarray(states).reshape(-1)
    n = len(states)
    players = np.ascontiguousarray(players, dtype=np.int8)
    best = np.zeros(n, dtype=MOVE)
    status = np.zeros(n, dtype=np.int32)
    lib().orc_mcts_search_bg_batch(_p(states), C.c_int(n), _p(players), _p(cfg), C.c_uint64(seed), C.c_uint32(first_game_id),
                                   C.c_uint32(epoch), _p(best), _p(status), C.c_int(nthreads))
    return best, status


def bg_playout_batch(states, seed, first_game_id, round_limit, nthreads):
    states = np.ascontiguousarray(states).reshape(-1)
    n = len(states)
    winners = np.zeros(n, dtype=np.int8)
    plies = np.zeros(n, dtype=np.int32)
    lib().orc_bg_playout_batch(_p(states), C.c_int(n), C.c_uint64(seed), C.c_uint32(first_game_id), C.c_int(round_limit),
                               _p(winners), _p(plies), C.c_int(nthreads))
    return winners, plies


# ---- AlphaZero search / self-play oracle ----
ANODE_A = np.dtype([("parent", "i4"), ("first_child", "i4"), ("n_children", "i4"), ("visits", "f4"), ("value", "f4"),
                    ("prior", "f4"), ("action", MOVE), ("state", BG_STATE)])
TRAJ = np.dtype([("state", BG_STATE), ("game_id", "u4"), ("ply", "u2"), ("outcome", "i1"), ("pad", "u1"),
                 ("n_pi", "u2"), ("pad2", "u2"), ("pi_offset", "u4")])
EVAL_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p)


def make_eval(fn):
    """fn(states: BG_STATE array) -> (policy [n,1352] f32, value [n] f32); returns a C callback"""
    def cb(states_p, n, policy_p, value_p, _user):
        st = np.ctypeslib.as_array(C.cast(states_p, C.POINTER(C.c_uint8)), shape=(n * 32,)).view(BG_STATE).copy()
        p, v = fn(st)
        np.ctypeslib.as_array(C.cast(policy_p, C.POINTER(C.c_float)), shape=(n * 1352,))[:] = np.asarray(p, dtype=np.float32).reshape(-1)
        np.ctypeslib.as_array(C.cast(value_p, C.POINTER(C.c_float)), shape=(n,))[:] = np.asarray(v, dtype=np.float32).reshape(-1)
    return EVAL_FN(cb)


def dirichlet(seed, epoch, alpha, n=1352):
    out = np.zeros(n, dtype=np.float32)
    lib().orc_dirichlet(C.c_uint64(seed), C.c_uint32(epoch), C.c_float(alpha), C.c_int(n), _p(out))
    return out


def alpha_mcts_parallel(states, game_ids, cfg, seed, epoch, eval_cb, max_nodes):
    states = np.ascontiguousarray(states).reshape(-1)
    n = len(states)
    game_ids = np.ascontiguousarray(game_ids, dtype=np.uint32)
    nodes = np.zeros((n, max_nodes), dtype=ANODE_A)
    n_nodes = np.zeros(n, dtype=np.int32)
    status = np.zeros(n, dtype=np.int32)
    lib().orc_alpha_mcts_parallel(_p(states), C.c_int(n), _p(game_ids), _p(cfg), C.c_uint64(seed), C.c_uint32(epoch), eval_cb,
                                  C.c_void_p(0), C.c_int(max_nodes), _p(nodes), _p(n_nodes), _p(status))
    return nodes, n_nodes, status


def root_pi(nodes_row, temperature_inv):
    """get_prob_tensor_parallel row + pow(1/T) for one game's tree (utils.rs:42-58, alpha_parallel.rs:164-166)"""
    row = np.ascontiguousarray(nodes_row)
    ids = np.zeros(MAX_MOVES, dtype=np.uint16)
    pi = np.zeros(MAX_MOVES, dtype=np.float32)
    n = lib().orc_root_pi(_p(row), C.c_float(temperature_inv), _p(ids), _p(pi))
    return ids[:n].copy(), pi[:n].copy()


def weighted_select(ids, pi, seed, game_id, ply):
    ids = np.ascontiguousarray(ids, dtype=np.uint16)
    pi = np.ascontiguousarray(pi, dtype=np.float32)
    return int(lib().orc_weighted_select(_p(ids), _p(pi), C.c_int(len(ids)), C.c_uint64(seed), C.c_uint32(game_id), C.c_uint32(ply)))


def self_play(n_games, cfg, temperature, seed, first_game_id, eval_cb, max_nodes):
    limit = int(cfg["simulate_round_limit"][0])
    rec_cap = n_games * (2 * limit + 4)
    pi_cap = rec_cap * 48
    rec = np.zeros(rec_cap, dtype=TRAJ)
    pi_ids = np.zeros(pi_cap, dtype=np.uint16)
    pi_vals = np.zeros(pi_cap, dtype=np.float32)
    n_rec, n_pi, n_waves = C.c_int(0), C.c_int(0), C.c_int(0)
    rc = lib().orc_self_play(C.c_int(n_games), _p(cfg), C.c_float(temperature), C.c_uint64(seed), C.c_uint32(first_game_id),
                             eval_cb, C.c_void_p(0), C.c_int(max_nodes), _p(rec), C.c_int(rec_cap), _p(pi_ids), _p(pi_vals),
                             C.c_int(pi_cap), C.byref(n_rec), C.byref(n_pi), C.byref(n_waves))
    assert rc == 0, rc
    return rec[: n_rec.value], pi_ids[: n_pi.value], pi_vals[: n_pi.value], n_waves.value
