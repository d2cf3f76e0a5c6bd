"""ctypes binding of libdiee_cuda.so (include/diee.h).  No CPU fallback: a missing library or a
missing CUDA device raises."""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libdiee_cuda.so")

OK = 0
ERR_INVALID, ERR_CUDA, ERR_NO_MOVES_PANIC, ERR_OVERFLOW, ERR_NOT_ROLLED = -1, -2, -3, -4, -5
MAX_MOVES = 256
NONE = -2
ACTION_SPACE = 1352
GAME_BACKGAMMON, GAME_TICTACTOE = 0, 1
STREAM_INIT, STREAM_GAME, STREAM_ROLLOUT, STREAM_EXPAND, STREAM_DIRICHLET, STREAM_SAMPLE = range(6)
MODE_ROLLOUT_CHECK_CURRENT = 1
MODE_PASS_CHILD = 2

BG_STATE = np.dtype([("pts", "i1", (24,)), ("bar", "u1", (2,)), ("off", "u1", (2,)),
                     ("roll", "u1", (2,)), ("player", "i1"), ("second", "u1")])
MOVE = np.dtype([("from1", "i1"), ("to1", "i1"), ("from2", "i1"), ("to2", "i1")])
TTT_STATE = np.dtype([("board", "i1", (9,)), ("player", "i1"), ("pad", "u1", (6,))])
MCTS_CFG = np.dtype([("iterations", "u4"), ("c", "f4"), ("simulate_round_limit", "u4"),
                     ("dirichlet_alpha", "f4"), ("dirichlet_epsilon", "f4"), ("mode_flags", "u4")])
NODE = np.dtype([("parent", "i4"), ("visits", "f4"), ("value", "f4"), ("action", MOVE),
                 ("n_moves", "i4"), ("n_untried", "i4")])
SEARCH_STATS = np.dtype([("rollout_plies", "u8"), ("select_levels", "u4"), ("select_children", "u4"),
                         ("expansions", "u4"), ("terminal_leaves", "u4")])
ANODE = np.dtype([("parent", "i4"), ("first_child", "i4"), ("n_children", "i4"), ("visits", "f4"), ("value", "f4"),
                  ("prior", "f4"), ("action", MOVE), ("state", BG_STATE)])
TRAJ = np.dtype([("state", BG_STATE), ("game_id", "u4"), ("ply", "u2"), ("outcome", "i1"), ("pad", "u1"),
                 ("n_pi", "u2"), ("pad2", "u2"), ("pi_offset", "u4")])
SELFPLAY_OPTS = np.dtype([("flags", "u4"), ("max_waves", "i4"), ("leaves_per_game", "i4"), ("target_games", "i4"),
                          ("virtual_loss", "f4"), ("pad", "u4")])
SELFPLAY_REPORT = np.dtype([("waves", "i4"), ("games_finished", "i4"), ("games_cut", "i4"), ("pad", "i4"), ("game_moves", "u8")])
SP_REFILL = 1
assert ANODE.itemsize == 60 and TRAJ.itemsize == 48 and SELFPLAY_OPTS.itemsize == 24 and SELFPLAY_REPORT.itemsize == 24
assert SEARCH_STATS.itemsize == 24
assert BG_STATE.itemsize == 32 and MOVE.itemsize == 4 and TTT_STATE.itemsize == 16 and NODE.itemsize == 24

# every symbol include/diee.h declares (tests check the .so exports each one)
SYMBOLS = [
    "diee_ctx_create", "diee_ctx_destroy", "diee_ctx_set_stream", "diee_sync", "diee_last_error", "diee_version",
    "diee_launch_count", "diee_dev_alloc", "diee_dev_free", "diee_dev_upload", "diee_dev_download", "diee_philox",
    "diee_bg_valid_moves", "diee_bg_valid_moves_dev", "diee_bg_apply_moves", "diee_bg_apply_moves_dev",
    "diee_bg_playout", "diee_bg_playout_dev", "diee_bg_encode_moves", "diee_bg_decode_moves",
    "diee_bg_encode_states", "diee_bg_encode_states_dev", "diee_mcts_search", "diee_mcts_search_dev",
    "diee_net_create", "diee_net_destroy", "diee_net_set_precision", "diee_net_param_count", "diee_net_forward", "diee_net_forward_dev",
    "diee_search_timing", "diee_search_work", "diee_comm_unique_id", "diee_comm_init", "diee_comm_destroy", "diee_traj_allgather",
    "diee_net_broadcast", "diee_dirichlet", "diee_alpha_search", "diee_alpha_search_dev", "diee_alpha_search_vl", "diee_selfplay_run", "diee_selfplay_run_ex", "diee_net_eval_count",
    "diee_arena_create", "diee_arena_round", "diee_arena_read", "diee_arena_destroy",
]


class DieeError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"diee error {code}: {msg}")
        self.code = code


_lib = None


def lib():
    """loads libdiee_cuda.so; raises if it has not been built (python -m die_e_b200.build)"""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: build it with `python die_e_b200/build.py` "
                              "(there is no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        L.diee_last_error.restype = C.c_char_p
        L.diee_last_error.argtypes = [C.c_void_p]
        L.diee_version.restype = C.c_char_p
        L.diee_launch_count.restype = C.c_int64
        L.diee_launch_count.argtypes = [C.c_void_p]
        L.diee_net_param_count.restype = C.c_int64
        L.diee_net_param_count.argtypes = [C.c_void_p]
        L.diee_net_eval_count.restype = C.c_uint64
        L.diee_net_eval_count.argtypes = [C.c_void_p]
        _lib = L
    return _lib


def _p(a):
    if a is None:
        return C.c_void_p(0)
    if isinstance(a, np.ndarray):
        assert a.flags["C_CONTIGUOUS"]
        return a.ctypes.data_as(C.c_void_p)
    return C.c_void_p(int(a))  # raw device pointer


def philox(seed, c0, c1, c2, c3):
    out = np.zeros(4, dtype=np.uint32)
    lib().diee_philox(C.c_uint64(seed), C.c_uint32(c0), C.c_uint32(c1), C.c_uint32(c2), C.c_uint32(c3), _p(out))
    return out


def dirichlet(seed, epoch, alpha, n=ACTION_SPACE):
    """the shared root-noise vector of one search wave (mcts/noise.rs:27-34); host-side, no GPU needed"""
    out = np.zeros(n, dtype=np.float32)
    rc = lib().diee_dirichlet(C.c_uint64(seed), C.c_uint32(epoch), C.c_float(alpha), C.c_int32(n), _p(out))
    if rc != OK:
        raise DieeError(rc, "diee_dirichlet: bad argument")
    return out


COMM_ID_BYTES = 128


def _prefer_bundled_nccl():
    """points DIEE_NCCL_LIB at the NCCL wheel next to torch (if there is one and nothing was chosen): whichever
    libnccl.so.2 is mapped first is the one a later `import torch` gets, and the system's may be older than torch needs"""
    if os.environ.get("DIEE_NCCL_LIB"):
        return
    import importlib.util
    try:
        spec = importlib.util.find_spec("nvidia.nccl")
    except (ImportError, ValueError):
        spec = None
    for base in (list(spec.submodule_search_locations) if spec and spec.submodule_search_locations else []):
        cand = os.path.join(base, "lib", "libnccl.so.2")
        if os.path.exists(cand):
            os.environ["DIEE_NCCL_LIB"] = cand
            return


def comm_unique_id():
    """the NCCL rendezvous id: rank 0 draws it and ships it to the other ranks (diee_comm_init)"""
    _prefer_bundled_nccl()
    out = np.zeros(COMM_ID_BYTES, dtype=np.uint8)
    rc = lib().diee_comm_unique_id(_p(out))
    if rc != OK:
        raise DieeError(rc, "diee_comm_unique_id: NCCL is not available")
    return out.tobytes()


def die_of(w):
    return 1 + ((int(w) * 6) >> 32)


def index_of(w, n):
    return (int(w) * n) >> 32


class Context:
    """one diee_ctx (one GPU)"""

    def __init__(self, device=0):
        self._h = C.c_void_p(0)
        rc = lib().diee_ctx_create(C.c_int32(device), C.byref(self._h))
        if rc != OK:
            raise DieeError(rc, "diee_ctx_create failed: no usable CUDA device (there is no CPU fallback)")

    def close(self):
        if self._h:
            lib().diee_ctx_destroy(self._h)
            self._h = C.c_void_p(0)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _chk(self, rc):
        if rc != OK:
            raise DieeError(rc, lib().diee_last_error(self._h).decode())

    def set_stream(self, cuda_stream):
        self._chk(lib().diee_ctx_set_stream(self._h, C.c_void_p(int(cuda_stream) if cuda_stream else 0)))

    def sync(self):
        self._chk(lib().diee_sync(self._h))

    def launch_count(self):
        return int(lib().diee_launch_count(self._h))

    def search_timing(self):
        """(tree_ms, rollout_ms) of the last reference-exact backgammon search, from CUDA events on its stream"""
        a, b = C.c_float(0), C.c_float(0)
        self._chk(lib().diee_search_timing(self._h, C.byref(a), C.byref(b)))
        return float(a.value), float(b.value)

    def search_work(self):
        """plies the rollouts of the last backgammon search actually played (closed-form tails excluded)"""
        n = C.c_uint64(0)
        self._chk(lib().diee_search_work(self._h, C.byref(n)))
        return int(n.value)

    # ---- env, host buffers ----
    def bg_valid_moves(self, states, want_ids=False):
        states = np.ascontiguousarray(states, dtype=BG_STATE).reshape(-1)
        n = len(states)
        moves = np.zeros((n, MAX_MOVES), dtype=MOVE)
        counts = np.zeros(n, dtype=np.int32)
        ids = np.zeros((n, MAX_MOVES), dtype=np.uint16) if want_ids else None
        self._chk(lib().diee_bg_valid_moves(self._h, _p(states), C.c_int32(n), _p(moves), _p(counts), _p(ids)))
        return (moves, counts, ids) if want_ids else (moves, counts)

    def bg_apply_moves(self, states, moves, next_rolls):
        states = np.ascontiguousarray(states, dtype=BG_STATE).reshape(-1).copy()
        moves = np.ascontiguousarray(moves, dtype=MOVE).reshape(-1)
        next_rolls = np.ascontiguousarray(next_rolls, dtype=np.uint8).reshape(-1)
        assert len(moves) == len(states) and len(next_rolls) == 2 * len(states)
        self._chk(lib().diee_bg_apply_moves(self._h, _p(states), _p(moves), _p(next_rolls), C.c_int32(len(states))))
        return states

    def bg_playout(self, starts, seed, first_game_id=0, round_limit=400, want_finals=False):
        starts = np.ascontiguousarray(starts, dtype=BG_STATE).reshape(-1)
        n = len(starts)
        winners = np.zeros(n, dtype=np.int8)
        plies = np.zeros(n, dtype=np.int32)
        finals = np.zeros(n, dtype=BG_STATE) if want_finals else None
        self._chk(lib().diee_bg_playout(self._h, _p(starts), C.c_int32(n), C.c_uint64(seed), C.c_uint32(first_game_id),
                                        C.c_int32(round_limit), _p(winners), _p(plies), _p(finals)))
        return (winners, plies, finals) if want_finals else (winners, plies)

    def bg_encode_moves(self, states, moves):
        states = np.ascontiguousarray(states, dtype=BG_STATE).reshape(-1)
        moves = np.ascontiguousarray(moves, dtype=MOVE).reshape(-1)
        ids = np.zeros(len(states), dtype=np.uint16)
        self._chk(lib().diee_bg_encode_moves(self._h, _p(states), _p(moves), C.c_int32(len(states)), _p(ids)))
        return ids

    def bg_decode_moves(self, states, ids):
        states = np.ascontiguousarray(states, dtype=BG_STATE).reshape(-1)
        ids = np.ascontiguousarray(ids, dtype=np.uint16).reshape(-1)
        moves = np.zeros(len(states), dtype=MOVE)
        self._chk(lib().diee_bg_decode_moves(self._h, _p(states), _p(ids), C.c_int32(len(states)), _p(moves)))
        return moves

    def bg_encode_states(self, states):
        states = np.ascontiguousarray(states, dtype=BG_STATE).reshape(-1)
        out = np.zeros((len(states), 6, 4, 6), dtype=np.float32)
        self._chk(lib().diee_bg_encode_states(self._h, _p(states), C.c_int32(len(states)), _p(out)))
        return out

    # ---- pure MCTS, host buffers ----
    def mcts_search(self, game_kind, states, players, cfg, seed, first_game_id=0, epoch=0, dump=False):
        sdt = BG_STATE if game_kind == GAME_BACKGAMMON else TTT_STATE
        states = np.ascontiguousarray(states, dtype=sdt).reshape(-1)
        n = len(states)
        players = np.ascontiguousarray(players, dtype=np.int8).reshape(-1)
        assert len(players) == n
        cfg = np.ascontiguousarray(cfg, dtype=MCTS_CFG).reshape(-1)[:1]
        best = np.zeros(n, dtype=MOVE) if game_kind == GAME_BACKGAMMON else np.zeros(n, dtype=np.uint8)
        status = np.zeros(n, dtype=np.int32)
        stats = np.zeros(n, dtype=SEARCH_STATS)
        cap = int(cfg["iterations"][0]) + 1
        nodes = np.zeros((n, cap), dtype=NODE) if dump else None
        nstates = np.zeros((n, cap), dtype=sdt) if dump else None
        n_nodes = np.zeros(n, dtype=np.int32) if dump else None
        finals = np.zeros((n, cap - 1), dtype=sdt) if dump else None
        self._chk(lib().diee_mcts_search(self._h, C.c_int32(game_kind), _p(states), C.c_int32(n), _p(players), _p(cfg),
                                         C.c_uint64(seed), C.c_uint32(first_game_id), C.c_uint32(epoch), _p(best),
                                         _p(status), _p(nodes), _p(nstates), _p(n_nodes), _p(stats), _p(finals)))
        if dump:
            return best, status, stats, nodes, nstates, n_nodes, finals
        return best, status, stats

    # ---- AlphaZero search / self-play ----
    def net_eval_count(self):
        return int(lib().diee_net_eval_count(self._h))

    def alpha_search(self, net, states, game_ids, cfg, seed, epoch=0, max_nodes=0, dump=False):
        states = np.ascontiguousarray(states, dtype=BG_STATE).reshape(-1)
        n = len(states)
        game_ids = np.ascontiguousarray(game_ids, dtype=np.uint32).reshape(-1)
        assert len(game_ids) == n
        cfg = np.ascontiguousarray(cfg, dtype=MCTS_CFG).reshape(-1)[:1]
        if max_nodes <= 0:
            max_nodes = 1 + (int(cfg["iterations"][0]) + 1) * 128
        ids = np.zeros((n, MAX_MOVES), dtype=np.uint16)
        moves = np.zeros((n, MAX_MOVES), dtype=MOVE)
        visits = np.zeros((n, MAX_MOVES), dtype=np.float32)
        counts = np.zeros(n, dtype=np.int32)
        status = np.zeros(n, dtype=np.int32)
        nodes = np.zeros((n, max_nodes), dtype=ANODE) if dump else None
        n_nodes = np.zeros(n, dtype=np.int32) if dump else None
        self._chk(lib().diee_alpha_search(self._h, net._h, _p(states), C.c_int32(n), _p(game_ids), _p(cfg), C.c_uint64(seed),
                                          C.c_uint32(epoch), C.c_int32(max_nodes), _p(ids), _p(moves), _p(visits), _p(counts),
                                          _p(status), _p(nodes), _p(n_nodes)))
        if dump:
            return ids, moves, visits, counts, status, nodes, n_nodes
        return ids, moves, visits, counts, status

    def alpha_search_vl(self, net, states, game_ids, cfg, seed, epoch=0, max_nodes=0, leaves_per_game=4, virtual_loss=1.0):
        """NON-PARITY search: several leaves per game and step with virtual loss (diee_alpha_search_vl)"""
        states = np.ascontiguousarray(states, dtype=BG_STATE).reshape(-1)
        n = len(states)
        game_ids = np.ascontiguousarray(game_ids, dtype=np.uint32).reshape(-1)
        cfg = np.ascontiguousarray(cfg, dtype=MCTS_CFG).reshape(-1)[:1]
        ids = np.zeros((n, MAX_MOVES), dtype=np.uint16)
        moves = np.zeros((n, MAX_MOVES), dtype=MOVE)
        visits = np.zeros((n, MAX_MOVES), dtype=np.float32)
        counts = np.zeros(n, dtype=np.int32)
        status = np.zeros(n, dtype=np.int32)
        self._chk(lib().diee_alpha_search_vl(self._h, net._h, _p(states), C.c_int32(n), _p(game_ids), _p(cfg), C.c_uint64(seed),
                                             C.c_uint32(epoch), C.c_int32(max_nodes), C.c_int32(leaves_per_game),
                                             C.c_float(virtual_loss), _p(ids), _p(moves), _p(visits), _p(counts), _p(status)))
        return ids, moves, visits, counts, status

    def alpha_search_dev(self, net, d_states, n, d_ids, cfg, seed, epoch, max_nodes, d_root_ids, d_root_moves, d_root_visits,
                         d_root_counts, d_status):
        cfg = np.ascontiguousarray(cfg, dtype=MCTS_CFG).reshape(-1)[:1]
        self._chk(lib().diee_alpha_search_dev(self._h, net._h, _p(d_states), C.c_int32(n), _p(d_ids), _p(cfg), C.c_uint64(seed),
                                              C.c_uint32(epoch), C.c_int32(max_nodes), _p(d_root_ids), _p(d_root_moves),
                                              _p(d_root_visits), _p(d_root_counts), _p(d_status)))

    def selfplay_run_ex(self, net, n_games, cfg, temperature, seed, first_game_id=0, max_nodes=0, rec_cap=None, pi_cap=None,
                        max_waves=0, flags=0, leaves_per_game=0, target_games=0, virtual_loss=0.0):
        """diee_selfplay_run_ex: (records, pi ids, pi values, report); see include/diee.h for the options"""
        cfg = np.ascontiguousarray(cfg, dtype=MCTS_CFG).reshape(-1)[:1]
        limit = int(cfg["simulate_round_limit"][0])
        opts = np.zeros(1, dtype=SELFPLAY_OPTS)
        opts[0] = (flags, max_waves, leaves_per_game, target_games, virtual_loss, 0)
        if not rec_cap:
            if flags & SP_REFILL:
                rec_cap = (target_games + n_games) * 300
            elif max_waves:
                rec_cap = n_games * (max_waves + 2)   # a time-boxed run records at most one fragment per game and wave
            else:
                rec_cap = n_games * (2 * limit + 4)
        pi_cap = pi_cap or rec_cap * 48
        rec = np.zeros(rec_cap, dtype=TRAJ)
        pi_ids = np.zeros(pi_cap, dtype=np.uint16)
        pi_vals = np.zeros(pi_cap, dtype=np.float32)
        rep = np.zeros(1, dtype=SELFPLAY_REPORT)
        n_rec, n_pi, n_waves = C.c_int32(0), C.c_int32(0), C.c_int32(0)
        self._chk(lib().diee_selfplay_run_ex(self._h, net._h, C.c_int32(n_games), _p(cfg), C.c_float(temperature), C.c_uint64(seed),
                                             C.c_uint32(first_game_id), C.c_int32(max_nodes), _p(opts), _p(rec), C.c_int32(rec_cap),
                                             _p(pi_ids), _p(pi_vals), C.c_int32(pi_cap), C.byref(n_rec), C.byref(n_pi),
                                             C.byref(n_waves), _p(rep)))
        return rec[: n_rec.value], pi_ids[: n_pi.value], pi_vals[: n_pi.value], rep[0]

    def selfplay_run(self, net, n_games, cfg, temperature, seed, first_game_id=0, max_nodes=0, rec_cap=None, pi_cap=None):
        cfg = np.ascontiguousarray(cfg, dtype=MCTS_CFG).reshape(-1)[:1]
        limit = int(cfg["simulate_round_limit"][0])
        rec_cap = rec_cap or n_games * (2 * limit + 4)
        pi_cap = pi_cap or rec_cap * 48
        rec = np.zeros(rec_cap, dtype=TRAJ)
        pi_ids = np.zeros(pi_cap, dtype=np.uint16)
        pi_vals = np.zeros(pi_cap, dtype=np.float32)
        n_rec, n_pi, n_waves = C.c_int32(0), C.c_int32(0), C.c_int32(0)
        self._chk(lib().diee_selfplay_run(self._h, net._h, C.c_int32(n_games), _p(cfg), C.c_float(temperature), C.c_uint64(seed),
                                          C.c_uint32(first_game_id), C.c_int32(max_nodes), _p(rec), C.c_int32(rec_cap), _p(pi_ids),
                                          _p(pi_vals), C.c_int32(pi_cap), C.byref(n_rec), C.byref(n_pi), C.byref(n_waves)))
        return rec[: n_rec.value], pi_ids[: n_pi.value], pi_vals[: n_pi.value], n_waves.value

    # ---- multi-GPU exchange over NCCL (one rank per context) ----
    def comm_init(self, nranks, rank, unique_id):
        _prefer_bundled_nccl()
        uid = np.frombuffer(bytes(unique_id), dtype=np.uint8).copy()
        assert len(uid) == COMM_ID_BYTES
        self._chk(lib().diee_comm_init(self._h, C.c_int32(nranks), C.c_int32(rank), _p(uid)))

    def comm_destroy(self):
        self._chk(lib().diee_comm_destroy(self._h))

    def traj_allgather(self, rec, pi_ids, pi_vals, rec_cap, pi_cap):
        """every rank's packed records -> all of them on every rank, rank-major, pi_offset rebased"""
        rec = np.ascontiguousarray(rec, dtype=TRAJ).reshape(-1)
        pi_ids = np.ascontiguousarray(pi_ids, dtype=np.uint16).reshape(-1)
        pi_vals = np.ascontiguousarray(pi_vals, dtype=np.float32).reshape(-1)
        assert len(pi_ids) == len(pi_vals)
        rec_out = np.zeros(max(1, rec_cap), dtype=TRAJ)
        ids_out = np.zeros(max(1, pi_cap), dtype=np.uint16)
        vals_out = np.zeros(max(1, pi_cap), dtype=np.float32)
        n_rec, n_pi = C.c_int32(0), C.c_int32(0)
        self._chk(lib().diee_traj_allgather(self._h, _p(rec), C.c_int32(len(rec)), _p(pi_ids), _p(pi_vals), C.c_int32(len(pi_ids)),
                                            _p(rec_out), C.c_int32(rec_cap), _p(ids_out), _p(vals_out), C.c_int32(pi_cap),
                                            C.byref(n_rec), C.byref(n_pi)))
        return rec_out[: n_rec.value], ids_out[: n_pi.value], vals_out[: n_pi.value]

    def net_broadcast(self, tensors, root=0):
        """in place: every rank's float32 arrays take rank `root`'s values"""
        for t in tensors:
            assert t.dtype == np.float32 and t.flags["C_CONTIGUOUS"]
        ptrs = (C.c_void_p * len(tensors))(*[t.ctypes.data for t in tensors])
        numels = np.array([t.size for t in tensors], dtype=np.int64)
        self._chk(lib().diee_net_broadcast(self._h, ptrs, _p(numels), C.c_int32(len(tensors)), C.c_int32(root)))

    # ---- device-pointer forms (ints = raw device addresses, e.g. torch tensor .data_ptr()) ----
    def bg_playout_dev(self, d_starts, n, seed, first_game_id, round_limit, d_winners, d_plies, d_finals=0):
        self._chk(lib().diee_bg_playout_dev(self._h, _p(d_starts), C.c_int32(n), C.c_uint64(seed), C.c_uint32(first_game_id),
                                            C.c_int32(round_limit), _p(d_winners), _p(d_plies), _p(d_finals)))

    def bg_valid_moves_dev(self, d_states, n, d_moves, d_counts, d_ids=0):
        self._chk(lib().diee_bg_valid_moves_dev(self._h, _p(d_states), C.c_int32(n), _p(d_moves), _p(d_counts), _p(d_ids)))

    def bg_encode_states_dev(self, d_states, n, d_out):
        self._chk(lib().diee_bg_encode_states_dev(self._h, _p(d_states), C.c_int32(n), _p(d_out)))

    def mcts_search_dev(self, game_kind, d_states, n, d_players, cfg, seed, first_game_id, epoch, d_best, d_status, d_stats=0):
        cfg = np.ascontiguousarray(cfg, dtype=MCTS_CFG).reshape(-1)[:1]
        self._chk(lib().diee_mcts_search_dev(self._h, C.c_int32(game_kind), _p(d_states), C.c_int32(n), _p(d_players), _p(cfg),
                                             C.c_uint64(seed), C.c_uint32(first_game_id), C.c_uint32(epoch), _p(d_best),
                                             _p(d_status), _p(d_stats)))


AGENT_RANDOM, AGENT_MCTS = 0, 1


class Arena:
    """a diee_arena handle: versus::play with the games resident on the device (include/diee.h)"""

    def __init__(self, ctx, n_games, seed, round_limit):
        self.ctx, self.n = ctx, n_games
        self._h = C.c_void_p(0)
        ctx._chk(lib().diee_arena_create(ctx._h, C.c_int32(n_games), C.c_uint64(seed), C.c_int32(round_limit), C.byref(self._h)))

    def round(self, agent_p1, agent_p2, cfg=None):
        """one round of the reference's loop -> (retired so far, wins p1, wins p2, draws)"""
        cfg = np.ascontiguousarray(cfg, dtype=MCTS_CFG).reshape(-1)[:1] if cfg is not None else None
        out = np.zeros(5, dtype=np.int32)
        self.ctx._chk(lib().diee_arena_round(self.ctx._h, self._h, C.c_int32(agent_p1), C.c_int32(agent_p2), _p(cfg), _p(out)))
        return int(out[0]), int(out[1]), int(out[2]), int(out[3])

    def read(self):
        states = np.zeros(self.n, dtype=BG_STATE)
        winners = np.zeros(self.n, dtype=np.int8)
        rounds = np.zeros(self.n, dtype=np.int32)
        self.ctx._chk(lib().diee_arena_read(self.ctx._h, self._h, _p(states), _p(winners), _p(rounds)))
        return states, winners, rounds

    def close(self):
        if self._h and self.ctx._h:
            lib().diee_arena_destroy(self.ctx._h, self._h)
        self._h = C.c_void_p(0)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


NET_BF16, NET_SPLIT3, NET_FP32 = 0, 1, 2


class Net:
    """a diee_net handle (policy/value ResNet on one ctx)"""

    def __init__(self, ctx, tensors, game_kind=GAME_BACKGAMMON):
        self.ctx = ctx
        arrs = [np.ascontiguousarray(t, dtype=np.float32) for t in tensors]
        ptrs = (C.c_void_p * len(arrs))(*[a.ctypes.data for a in arrs])
        numels = np.array([a.size for a in arrs], dtype=np.int64)
        self._h = C.c_void_p(0)
        ctx._chk(lib().diee_net_create(ctx._h, C.c_int32(game_kind), ptrs, _p(numels), C.c_int32(len(arrs)), C.byref(self._h)))

    def close(self):
        if self._h and self.ctx._h:
            lib().diee_net_destroy(self.ctx._h, self._h)
        self._h = C.c_void_p(0)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def param_count(self):
        return int(lib().diee_net_param_count(self._h))

    def set_precision(self, precision):
        """NET_BF16 (fast path), NET_SPLIT3 (tensor cores, 24-bit operands) or NET_FP32 (parity mode, CUDA-core fp32)"""
        self.ctx._chk(lib().diee_net_set_precision(self.ctx._h, self._h, C.c_int32(precision)))

    def forward(self, states):
        states = np.ascontiguousarray(states, dtype=BG_STATE).reshape(-1)
        n = len(states)
        policy = np.zeros((n, ACTION_SPACE), dtype=np.float32)
        value = np.zeros(n, dtype=np.float32)
        self.ctx._chk(lib().diee_net_forward(self.ctx._h, self._h, _p(states), C.c_int32(n), _p(policy), _p(value)))
        return policy, value

    def forward_dev(self, d_states, n, d_policy, d_value):
        self.ctx._chk(lib().diee_net_forward_dev(self.ctx._h, self._h, _p(d_states), C.c_int32(n), _p(d_policy), _p(d_value)))


_default_ctx = None


def default_context():
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = Context(0)
    return _default_ctx
