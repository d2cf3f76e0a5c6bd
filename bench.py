#!/usr/bin/env python3
"""bench.py -- the die-e hot path on B200: MCTS simulations/sec and self-play throughput.

Default workload = BASELINE.json configs[2]: backgammon pure MCTS (random rollouts),
iterations=100, exploration_const=2, simulate_round_limit=400, 1,024 games batched per GPU.
A "step" is one `mct_search` for every game of the batch (the call `versus.rs:303-306` makes each
arena round) = games x iterations simulations.  The default line also carries, under `detail`,
sub-records of the other configurations the metric names: the same search with rollouts that test the
rolled-out state (`check_current`), the AlphaZero search and a time-boxed self-play (configs[3]) per net
precision, and -- under --gpus N -- the one exchange step (all-gather of the trajectories through the C ABI).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload mcts|playout|alpha|selfplay] [--games G] [--rollout ref_exact|check_current]
                    [--precision bf16|split3|fp32] [--no-subrecords]

N > 1: launched by torchrun, one rank per GPU; games are sharded (rank r owns global game ids
[r*G, (r+1)*G)), no data-path collective (weak scaling); time = max over ranks.

--impl reference: the reference's own CPU path for the same config -- the C oracle restating
`mct_search` (the Rust reference cannot be built here: no cargo/rustc), one task per game over
all host threads like rayon's par_iter, on a bounded sample of the SAME inputs (same seed, same games).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SEED = 0xD1EE
OPENING = [2, 0, 0, 0, 0, -5, 0, -3, 0, 0, 0, 5, -5, 0, 0, 0, 3, 0, 5, 0, 0, 0, 0, -2]
NET_FLOP_PER_EVAL = 1.0825e9  # dense-tap 2*MAC count of one forward (256 filters, 19 blocks)
PRECISIONS = {"bf16": 0, "split3": 1, "fp32": 2}
PRECISION_DTYPE = {"bf16": "bf16 net operands, fp32 accumulate (NOT the reference's fp32: 4e-2 off on the value head)",
                   "split3": "3 bf16 planes per operand on tcgen05 (integer leading digit, exact big products): within the fp32 tolerance",
                   "fp32": "fp32 FMA on CUDA cores (the reference's arithmetic)"}


def peaks():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def tensor_peaks():
    """(burst, sustained) dense bf16 TFLOP/s"""
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(p["bf16_tflops"]), float(p["bf16_tflops_sustained"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 1590.0, 1400.0, "fallback (B200_PROFILING.md)"


def workload_config(args):
    """`config` of the JSON line: names the workload and its inputs.  BOTH arms print exactly this dict."""
    base = {"seed": hex(SEED), "games_per_gpu": args.games, "simulate_round_limit": args.round_limit,
            "inputs": "synthetic: game g (global id) = the opening position, first roll from its Philox INIT stream, advanced "
                      "10*(g % 9) plies of uniform random play on its GAME stream (SURVEY 8(d) M-inputs)"}
    if args.workload == "playout":
        base["inputs"] = "synthetic: game g (global id) = the opening position, first roll from its Philox INIT stream"
        base["workload"] = ("backgammon random-vs-random playouts (BASELINE configs[1]): env step + legal-move generation, "
                            "to a winner or the round limit")
        return base
    base.update(iterations=args.iterations, c=2.0)
    if args.workload == "mcts":
        base["workload"] = "backgammon pure MCTS, random rollouts (BASELINE configs[2]); one step = one mct_search per game"
        base["rollout"] = args.rollout
        base["no_move_nodes"] = "PASS_CHILD"
    else:
        base["workload"] = ("backgammon AlphaZero search, policy/value ResNet 256x19 with synthetic weights (BASELINE configs[3]); "
                            + ("one step = one alpha_mcts_parallel (one game-move of every game)" if args.workload == "alpha"
                               else "one step = one self_play_parallel from the opening to a winner"))
        base.update(dirichlet_alpha=0.3, dirichlet_epsilon=0.25, temperature=1.25)
    return base


class ClockSampler:
    """samples nvidia-smi clocks / throttle reasons during the timed region"""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.samples, self._stop, self._t = gpu_index, [], threading.Event(), None

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.idx), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self._stop.wait(0.2)

    def start(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()

    def stop(self):
        self._stop.set()
        if self._t:
            self._t.join(timeout=6)
        sm = [float(s[0]) for s in self.samples if s[0].replace(".", "").isdigit()]
        mx = [float(s[1]) for s in self.samples if s[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for s in self.samples for i in range(4) if len(s) >= 6 and s[2 + i].lower() == "active"})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.samples)}


# ---------------------------------------------------------------- inputs
def initial_states(ffi, first_gid, n):
    s = np.zeros(n, dtype=ffi.BG_STATE)
    s["pts"][:] = OPENING
    s["player"] = -1
    for g in range(n):
        w = ffi.philox(SEED, 0, first_gid + g, ffi.STREAM_INIT, 0)
        s["roll"][g] = (ffi.die_of(w[0]), ffi.die_of(w[1]))
    return s


def midgame_states(ctx, ffi, first_gid, n):
    """SURVEY 8(d) M-inputs: game g advanced k plies of random play, k = 10*(g % 9) in 0..80 (made with
    the product's own playout kernel); the few games per thousand that are already over by then keep their opening roll.
    host_states_for_reference() makes the same states with the oracle; parity_check compares the two."""
    s = initial_states(ffi, first_gid, n)
    out = s.copy()
    for grp in range(1, 9):
        idx = np.arange(n)[(np.arange(n) + first_gid) % 9 == grp]
        if len(idx) == 0:
            continue
        # one launch over the whole batch keeps every game on its own id-keyed stream
        _, _, fin_all = ctx.bg_playout(s, seed=SEED, first_game_id=first_gid, round_limit=10 * grp, want_finals=True)
        out[idx] = fin_all[idx]
    alive = (out["off"][:, 0] < 15) & (out["off"][:, 1] < 15)
    out[~alive] = s[~alive]
    return out


def host_states_for_reference(args, first_gid=0, n=None, midgame=None):
    """the same synthetic inputs as the GPU arm, generated with the oracle (no GPU needed)"""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import orc
    orc.build()
    G = args.games if n is None else n
    s = np.zeros(G, dtype=orc.BG_STATE)
    s["pts"][:] = OPENING
    s["player"] = -1
    for g in range(G):
        w = orc.philox(SEED, 0, first_gid + g, orc.STREAM_INIT, 0)
        s["roll"][g] = (orc.die(w[0]), orc.die(w[1]))
    if (args.workload in ("mcts", "alpha", "selfplay")) if midgame is None else midgame:
        start = s.copy()
        for g in range(G):
            for ply in range(10 * ((first_gid + g) % 9)):
                orc.bg_random_ply(s[g:g + 1], orc.philox(SEED, ply, first_gid + g, orc.STREAM_GAME, 0))
                if orc.bg_check_winner(s[g:g + 1]) is not None:
                    s[g] = start[g]  # a game that is over this early (a few per thousand) searches from its opening roll instead
                    break
    return s


# ---------------------------------------------------------------- roofline helpers
def ncu_traffic(summary_name):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` summary of the
    dominant kernel (profiles/), or None.  The capture is of the default configuration of that workload."""
    path = os.path.join(ROOT, "profiles", summary_name)
    if not os.path.exists(path):
        return None
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    total, seen = 0.0, 0
    for line in open(path):
        if line.startswith("dram__bytes_read.sum") or line.startswith("dram__bytes_write.sum"):
            unit = line.split("[")[1].split("]")[0]
            total += float(line.split("=")[1].replace(",", "")) * scale.get(unit, 1.0)
            seen += 1
    return total if seen == 2 else None


def first_existing(*names):
    for n in names:
        if os.path.exists(os.path.join(ROOT, "profiles", n)):
            return n
    return names[-1]


def rollout_kernel_name(n_items, playout=False):
    """which lane kernel the library picks for a job (lane_kernels.cu launch_lane_job): the packed one from 640 rollouts
    (400 playouts) per SM on, as ONE wave on two CTAs per SM up to 1,024 items per SM"""
    env = os.environ.get("DIEE_LANE_PACK", "1")
    sms = 148
    packed = env == "2" or (env != "0" and n_items >= sms * (400 if playout else 640))
    if not packed:
        return "lane_run_kernel (one lane per %s, warp-vote scheduling)" % ("game" if playout else "rollout")
    one_wave = n_items <= 2 * sms * 512
    return "lane_pack_kernel (games queued by the code of their next ply; %s)" % ("one wave, two CTAs per SM" if one_wave else "three CTAs per SM, up to 512 resident games each, refilled from the job queue")


def issue_record(name):
    """the issue-side reading of the dominant kernel: profiles/<name> is written by tools/ncu_issue.py from a committed
    ncu --set full export (never measured live: ncu replays kernels), or None"""
    path = os.path.join(ROOT, "profiles", name)
    try:
        return json.load(open(path))
    except Exception:
        return None


def bsim_bytes(stats, n_sims):
    """algorithmic HBM bytes per simulation of the tree kernel over the SoA pool (DESIGN.md section 4 derives each term):
    select: per level 8 B (move counts + visits of the node) + 12 B per child (parent, visits, value);
    expand: 32 B parent state + 4 B count update + 32 B child state + 20 B child fields = 88 B;
    backprop: 20 B per node on the path (visits+value read-modify-write = 16 B, parent link 4 B)."""
    lv = float(stats["select_levels"].sum())
    ch = float(stats["select_children"].sum())
    ex = float(stats["expansions"].sum())
    path_nodes = lv + n_sims  # a leaf at depth d has d+1 nodes on its path
    total = 8.0 * lv + 12.0 * ch + 88.0 * ex + 20.0 * path_nodes
    return total / max(1, n_sims)


# ---------------------------------------------------------------- parity checks inside the bench
def load_oracle():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import orc
    orc.build()
    return orc


def parity_check_mcts(ctx, ffi, args, cfg, h_states, h_players, first_gid, epoch, d_best_host, n_games=16):
    """What was TIMED is checked at the size it was timed: the whole batch is searched once more through the host call
    with the pool dumped (same seed / ids / epoch as the last timed step => the same search), its best moves must equal
    the timed step's, and `n_games` sampled games are compared with the oracle: best move, every node's visits / value
    bit-for-bit, and the end state of every rollout."""
    orc = load_oracle()
    G = len(h_states)
    best, status, stats, nodes, nstates, n_nodes, finals = ctx.mcts_search(ffi.GAME_BACKGAMMON, h_states, h_players, cfg, SEED,
                                                                          first_gid, epoch, dump=True)
    ok = bool((status == 0).all()) and best.view(np.int32).tobytes() == d_best_host.tobytes()
    sample = np.linspace(0, G - 1, n_games).astype(int)
    ocfg = orc.mcts_cfg(int(cfg["iterations"][0]), float(cfg["c"][0]), int(cfg["simulate_round_limit"][0]), 0.3, 0.25,
                        int(cfg["mode_flags"][0]))
    rollouts = 0
    ref_inputs = host_states_for_reference(args, first_gid, G) if G <= 2048 else None
    inputs_same = ref_inputs is not None and ref_inputs.tobytes() == np.ascontiguousarray(h_states).tobytes()
    for i in sample:
        rc, obest, onodes, ostates, ofin = orc.mcts_search_bg(h_states[i:i + 1], int(h_players[i]), ocfg, SEED, first_gid + int(i),
                                                              epoch, want_finals=True)
        k = len(onodes)
        good = (status[i] == rc and n_nodes[i] == k and best[i:i + 1].tobytes() == obest.tobytes()
                and nodes[i, :k]["visits"].tobytes() == onodes["visits"].tobytes()
                and nodes[i, :k]["value"].tobytes() == onodes["value"].tobytes()
                and nodes[i, :k]["parent"].tobytes() == onodes["parent"].tobytes()
                and nstates[i, :k].tobytes() == ostates.tobytes()
                and finals[i].tobytes() == ofin.tobytes())
        rollouts += int((np.frombuffer(ofin.tobytes(), dtype=np.uint8).reshape(-1, 32).any(axis=1)).sum())
        ok = ok and bool(good)
    return {"ok": bool(ok), "games": int(len(sample)), "of": int(G), "rollouts": int(rollouts),
            "timed_step_best_moves_equal_dump": best.view(np.int32).tobytes() == d_best_host.tobytes(),
            "inputs_identical_to_reference_arm": bool(inputs_same),
            "what": "best move, per-node visits/value/parent/state and every rollout's end state, bit-exact vs the C oracle"}


def parity_check_playout(ffi, args, h_states, first_gid, seed, winners, plies, finals, n_games=64):
    orc = load_oracle()
    G = len(h_states)
    sample = np.linspace(0, G - 1, n_games).astype(int)
    ok = True
    for i in sample:
        w, p, fin = orc.bg_playout(h_states[i:i + 1], seed, first_gid + int(i), args.round_limit)
        ok = ok and (int(winners[i]), int(plies[i])) == (w, p) and finals[i:i + 1].tobytes() == fin.tobytes()
    ref_inputs = host_states_for_reference(args, first_gid, min(G, 4096))
    return {"ok": bool(ok), "games": int(len(sample)), "of": int(G),
            "inputs_identical_to_reference_arm": ref_inputs.tobytes() == np.ascontiguousarray(h_states[:len(ref_inputs)]).tobytes(),
            "what": "winner, ply count and final state of sampled games, bit-exact vs the C oracle"}


# ---------------------------------------------------------------- the GPU arm
def run_ours(args, rank, world):
    import torch
    import torch.distributed as dist
    from die_e_b200 import _ffi as ffi

    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    ctx = ffi.Context(local)
    # an explicit (non-default) torch stream: the ctx launches on it, the timing events are recorded on it
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    ctx.set_stream(stream.cuda_stream)
    hbm_peak, peak_src = peaks()
    G = args.games
    first_gid = rank * G

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce_max(x):
        if world > 1:
            t = torch.tensor([float(x)], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return float(x)

    def reduce_sum(x):
        if world > 1:
            t = torch.tensor([float(x)], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            return float(t.item())
        return float(x)

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    net = None

    def make_net():
        from die_e_b200 import nnet
        nt = ffi.Net(ctx, nnet.synthetic_tensors(seed=SEED, filters=256, blocks=19, bn_stats="identity"))
        nt.set_precision(PRECISIONS[args.precision])
        return nt

    played_plies = []
    if args.workload == "mcts":
        mode = ffi.MODE_PASS_CHILD | (ffi.MODE_ROLLOUT_CHECK_CURRENT if args.rollout == "check_current" else 0)
        cfg = np.zeros(1, dtype=ffi.MCTS_CFG)
        cfg[0] = (args.iterations, 2.0, args.round_limit, 0.3, 0.25, mode)
        h_states = midgame_states(ctx, ffi, first_gid, G)
        h_players = h_states["player"].copy()
        d_states = torch.from_numpy(h_states.view(np.uint8).reshape(G, 32)).to(dev)
        d_players = torch.from_numpy(h_players).to(dev)
        d_best = torch.zeros(G, dtype=torch.int32, device=dev)
        d_status = torch.zeros(G, dtype=torch.int32, device=dev)
        d_stats = torch.zeros(G, 24, dtype=torch.uint8, device=dev)
        units_per_step = G * args.iterations

        def step_dev(i):
            ctx.mcts_search_dev(ffi.GAME_BACKGAMMON, d_states.data_ptr(), G, d_players.data_ptr(), cfg, SEED, first_gid,
                                i & 0xFFFF, d_best.data_ptr(), d_status.data_ptr(), d_stats.data_ptr())

        # e2e: the public host-buffer call (what versus.rs's Agent::Mcts arm maps to) from pinned memory
        p_states = torch.from_numpy(h_states.view(np.uint8).reshape(G, 32).copy()).pin_memory()
        p_players = torch.from_numpy(h_players.copy()).pin_memory()
        np_states, np_players = p_states.numpy().view(ffi.BG_STATE).reshape(-1), p_players.numpy()
        h2d, d2h = G * 32 + G, G * 4 + G * 4 + G * 24

        def step_e2e(i):
            best, status, stats = ctx.mcts_search(ffi.GAME_BACKGAMMON, np_states, np_players, cfg, SEED, first_gid, i & 0xFFFF)
            return stats

        metric, unit = "mcts_simulations_per_sec", "simulations/s"
        dtype = "int8 boards / f32 UCB"
    elif args.workload in ("alpha", "selfplay"):
        cfg = np.zeros(1, dtype=ffi.MCTS_CFG)
        cfg[0] = (args.iterations, 2.0, args.round_limit, 0.3, 0.25, 0)
        net = make_net()
        h_states = midgame_states(ctx, ffi, first_gid, G)
        h_ids = np.arange(first_gid, first_gid + G, dtype=np.uint32)
        d_states = torch.from_numpy(h_states.view(np.uint8).reshape(G, 32)).to(dev)
        d_ids = torch.from_numpy(h_ids.view(np.int32)).to(dev)
        d_rids = torch.zeros(G, ffi.MAX_MOVES, dtype=torch.int16, device=dev)
        d_rmoves = torch.zeros(G, ffi.MAX_MOVES, dtype=torch.int32, device=dev)
        d_rvis = torch.zeros(G, ffi.MAX_MOVES, dtype=torch.float32, device=dev)
        d_rcnt = torch.zeros(G, dtype=torch.int32, device=dev)
        d_status = torch.zeros(G, dtype=torch.int32, device=dev)
        units_per_step = G * args.iterations
        dtype = PRECISION_DTYPE[args.precision] + " / f32 PUCT"
        if args.workload == "alpha":
            def step_dev(i):
                ctx.alpha_search_dev(net, d_states.data_ptr(), G, d_ids.data_ptr(), cfg, SEED, i & 0xFFFF, 0, d_rids.data_ptr(),
                                     d_rmoves.data_ptr(), d_rvis.data_ptr(), d_rcnt.data_ptr(), d_status.data_ptr())
            p_states = torch.from_numpy(h_states.view(np.uint8).reshape(G, 32).copy()).pin_memory()
            np_states = p_states.numpy().view(ffi.BG_STATE).reshape(-1)
            h2d, d2h = G * 36, G * (ffi.MAX_MOVES * 10 + 8)

            def step_e2e(i):
                return ctx.alpha_search(net, np_states, h_ids, cfg, SEED, i & 0xFFFF)
            metric, unit = "alphazero_simulations_per_sec", "simulations/s"
        else:
            units_per_step = G  # games
            sp_info = {}

            def step_dev(i):
                t0 = time.perf_counter()
                e0 = ctx.net_eval_count()
                rec, pi_ids, pi_vals, rep = ctx.selfplay_run_ex(net, G, cfg, 1.25, SEED + i, first_gid, max_waves=args.max_waves,
                                                               flags=ffi.SP_REFILL if args.refill else 0, target_games=args.refill,
                                                               leaves_per_game=args.leaves, virtual_loss=1.0 if args.leaves > 1 else 0.0)
                sp_info.update(records=len(rec), waves=int(rep["waves"]), evals=ctx.net_eval_count() - e0, wall=time.perf_counter() - t0,
                               rec=(rec, pi_ids, pi_vals), finished=int(rep["games_finished"]), cut=int(rep["games_cut"]),
                               moves=int(rep["game_moves"]))
            h2d, d2h = 0, 0
            step_e2e = None
            time_boxed = args.max_waves > 0 and not args.refill
            metric, unit = ("selfplay_game_moves_per_sec", "game-moves/s") if time_boxed else ("selfplay_games_per_sec", "games/s")
    else:
        h_states = initial_states(ffi, first_gid, G)
        d_states = torch.from_numpy(h_states.view(np.uint8).reshape(G, 32)).to(dev)
        d_winners = torch.zeros(G, dtype=torch.int8, device=dev)
        d_plies = torch.zeros(G, dtype=torch.int32, device=dev)
        d_finals = torch.zeros(G, 32, dtype=torch.uint8, device=dev)

        def step_dev(i):
            ctx.bg_playout_dev(d_states.data_ptr(), G, SEED + i, first_gid, args.round_limit, d_winners.data_ptr(),
                               d_plies.data_ptr(), d_finals.data_ptr())

        p_states = torch.from_numpy(h_states.view(np.uint8).reshape(G, 32).copy()).pin_memory()
        np_states = p_states.numpy().view(ffi.BG_STATE).reshape(-1)
        h2d, d2h = G * 32, G * (1 + 4 + 32)

        def step_e2e(i):
            return ctx.bg_playout(np_states, SEED + i, first_gid, args.round_limit, want_finals=True)

        units_per_step = None  # plies: read back per step
        metric, unit = "playout_plies_per_sec", "plies/s"
        dtype = "int8 boards"

    # ---- warm-up ----
    for i in range(args.warmup):
        step_dev(i)
        torch.cuda.synchronize()
    # W short steps are a few milliseconds -- less than the SM clocks take to come up from idle, and under torchrun the
    # ranks do not start together: keep stepping (untimed) until 60 ms have passed, so that every rank times steady state
    extra_warm = 0
    if args.workload in ("mcts", "playout"):
        t_w = time.perf_counter()
        while time.perf_counter() - t_w < 0.06 and extra_warm < 256:
            step_dev(args.warmup + args.steps + 1000 + extra_warm)
            torch.cuda.synchronize()
            extra_warm += 1

    # ---- timed: device-resident (`value`) ----
    launches0 = ctx.launch_count()
    sampler = ClockSampler(local)
    sampler.start()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    units = 0.0
    stats_acc = None
    split_ms = []  # (tree kernel ms, rollout kernel ms) per timed step, from CUDA events inside the library
    barrier()
    t_wall0 = time.perf_counter()
    for i in range(args.steps):
        flush.fill_(i & 0xFF)  # L2 flush between timed iterations (outside the event bracket)
        evs[i][0].record(stream)
        step_dev(args.warmup + i)
        evs[i][1].record(stream)
        if args.workload == "playout":
            evs[i][1].synchronize()
            units += float(d_plies.sum().item())
        elif args.workload == "mcts":
            evs[i][1].synchronize()
            if args.rollout == "ref_exact":
                played_plies.append(ctx.search_work())
            st = d_stats.cpu().numpy().view(ffi.SEARCH_STATS).reshape(-1)
            stats_acc = st if stats_acc is None else np.concatenate([stats_acc, st])
            units += units_per_step
        else:
            evs[i][1].synchronize()
            if args.workload == "selfplay":  # games that reached a winner or the cap; a time-boxed run counts game-moves instead
                units += sp_info["moves"] if time_boxed else sp_info["finished"]
            else:
                units += units_per_step
    barrier()
    t_wall = time.perf_counter() - t_wall0
    launches = ctx.launch_count() - launches0
    last_epoch = (args.warmup + args.steps - 1) & 0xFFFF
    timed_best = d_best.cpu().numpy().copy() if args.workload == "mcts" else None
    timed_playout = (d_winners.cpu().numpy().copy(), d_plies.cpu().numpy().copy(),
                     d_finals.cpu().numpy().copy().view(ffi.BG_STATE).reshape(-1)) if args.workload == "playout" else None
    if args.workload == "mcts" and args.rollout == "ref_exact":
        # the two kernels of the search timed alone (untimed extra searches, one launch each back to back on the stream): the
        # timed steps may run them sliced and overlapped on an SM partition, where a per-kernel duration means nothing
        keep = os.environ.get("DIEE_SEARCH_SLICES")
        os.environ["DIEE_SEARCH_SLICES"] = "1"
        for i in range(3):
            flush.fill_(i)
            step_dev(args.warmup + args.steps + 100 + i)
            torch.cuda.synchronize()
            split_ms.append(ctx.search_timing())
        if keep is None:
            del os.environ["DIEE_SEARCH_SLICES"]
        else:
            os.environ["DIEE_SEARCH_SLICES"] = keep
    # The timed region of the short workloads is a few milliseconds, one nvidia-smi call takes longer: keep the SAME
    # step running (untimed, uncounted) until the sampler has seen the GPU under this load a few times.
    if args.workload != "selfplay":
        t_load = time.perf_counter()
        j = 0
        while len(sampler.samples) < 4 and time.perf_counter() - t_load < 3.0:
            step_dev(args.warmup + args.steps + j)
            torch.cuda.synchronize()
            j += 1
    clocks = sampler.stop()
    kern_ms = [a.elapsed_time(b) for a, b in evs]
    dev_ms = float(sum(kern_ms))
    if args.workload in ("mcts", "alpha"):
        assert int(d_status.abs().sum().item()) == 0

    # ---- what was timed is checked against the oracle, at the size it was timed (rank 0's shard) ----
    parity = None
    if rank == 0 and not args.no_parity_check:
        if args.workload == "mcts":
            parity = parity_check_mcts(ctx, ffi, args, cfg, h_states, h_players, first_gid, last_epoch, timed_best)
        elif args.workload == "playout":
            parity = parity_check_playout(ffi, args, h_states, first_gid, SEED + args.warmup + args.steps - 1, *timed_playout)
        else:
            parity = {"ok": None, "what": "the AlphaZero search couples the games of a batch (quirk Q9), so a sample of games cannot be "
                      "checked alone; the whole 1,024-game / 100-iteration / 256x19 search is compared with the oracle in "
                      "tests/test_gpu_baseline_sizes.py"}
        if parity.get("ok") is False:
            raise SystemExit("parity_check FAILED: " + json.dumps(parity))

    # ---- timed: end to end through the host-buffer C-ABI call ----
    exchange = None
    if args.workload == "selfplay":
        # diee_selfplay_run IS the public host call (opening positions are made inside, records land in host
        # buffers); what is added here is the one exchange step: all-gather of the trajectories over NCCL
        e2e_units, e2e_s = units, t_wall
        if world > 1:
            exchange = exchange_step(ctx, ffi, dist, dev, rank, world, sp_info["rec"], barrier)
            e2e_s += exchange["seconds"]
        d2h = int(sp_info["records"]) * 48
    else:
        step_e2e(0)
        barrier()
        t0 = time.perf_counter()
        e2e_units = 0.0
        for i in range(args.steps):
            r = step_e2e(args.warmup + i)
            e2e_units += float(r[1].sum()) if args.workload == "playout" else units_per_step
        barrier()
        e2e_s = time.perf_counter() - t0

    # ---- sub-records of the default line (every rank runs them; time = max over ranks, units summed) ----
    detail = {}
    default_line = (args.workload == "mcts" and args.rollout == "ref_exact" and G == 1024 and args.iterations == 100
                    and not args.no_subrecords)
    if default_line:
        # (1) the same search at a throughput batch: 8,192 games per GPU ("thousands of concurrent games", north_star).
        # Not the headline (`value` is BASELINE configs[2], 1,024 games per GPU); it shows how far the 1,024-game figure
        # is bound by the latency of one game's 100 sequential iterations rather than by issue slots.
        detail["large_batch"] = sub_mcts(ctx, ffi, torch, dev, stream, rank, world, 8192, cfg, reduce_max)
        detail["large_batch_32768"] = sub_mcts(ctx, ffi, torch, dev, stream, rank, world, 32768, cfg, reduce_max, reps=2)
        detail["large_batch_65536"] = sub_mcts(ctx, ffi, torch, dev, stream, rank, world, 65536, cfg, reduce_max, reps=2)
        # (2) rollouts that test the rolled-out state (the evident intent of node.rs:181, quirk Q5): the mode in which the
        # rollouts decide the search
        cc = np.zeros(1, dtype=ffi.MCTS_CFG)
        cc[0] = (args.iterations, 2.0, args.round_limit, 0.3, 0.25, ffi.MODE_PASS_CHILD | ffi.MODE_ROLLOUT_CHECK_CURRENT)
        detail["check_current"] = {"1024_games": sub_mcts(ctx, ffi, torch, dev, stream, rank, world, 1024, cc, reduce_max, reps=2),
                                   "16384_games": sub_mcts(ctx, ffi, torch, dev, stream, rank, world, 16384, cc, reduce_max, reps=1),
                                   "65536_games": sub_mcts(ctx, ffi, torch, dev, stream, rank, world, 65536, cc, reduce_max, reps=1),
                                   "note": "1,024 games: the whole search fused in one launch (warp per game); 16,384 and 65,536 games: "
                                           "lock-step on the lane engine, one tree launch + one rollout launch per iteration"}
        # (3) configs[3]: AlphaZero search + time-boxed self-play, per net precision that fits the time budget
        detail["alphazero"] = sub_alpha(ctx, ffi, torch, dev, stream, rank, world, args, h_states, reduce_max, reduce_sum, barrier,
                                        dist if world > 1 else None)

    # ---- reduce over ranks: max time, summed units ----
    if world > 1 and args.workload == "selfplay":
        for k in ("records", "evals", "finished", "cut", "moves"):
            sp_info[k] = int(reduce_sum(sp_info[k]))
    if world > 1:
        t = torch.tensor([dev_ms, e2e_s, t_wall], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        u = torch.tensor([units, e2e_units, float(launches)], device=dev, dtype=torch.float64)
        dist.all_reduce(u, op=dist.ReduceOp.SUM)
        dev_ms, e2e_s, t_wall = [float(x) for x in t.tolist()]
        units, e2e_units, launches = [float(x) for x in u.tolist()]

    out = None
    if rank == 0:
        value = units / (dev_ms / 1e3)
        dominant = None
        roof_extra = {}
        if args.workload == "mcts":
            n_sims_local = G * args.iterations * args.steps
            b_sim = bsim_bytes(stats_acc, n_sims_local)
            plies_per_sim = float(stats_acc["rollout_plies"].sum()) / n_sims_local
            detail.update(tree_bytes_per_simulation=round(b_sim, 1), rollout_plies_per_simulation=round(plies_per_sim, 2),
                          mean_select_depth=round(float(stats_acc["select_levels"].sum()) / n_sims_local, 3))
            if split_ms:
                tree_ms = float(np.mean([a for a, _ in split_ms]))
                roll_ms = float(np.mean([b for _, b in split_ms]))
                played = float(np.mean(played_plies))
                detail.update(tree_kernel_ms=round(tree_ms, 4), rollout_kernel_ms=round(roll_ms, 4),
                              rollout_plies_played_per_search=int(played),
                              rollout_plies_resolved_in_closed_form_per_search=int(plies_per_sim * G * args.iterations - played),
                              rollout_plies_played_per_sec=round(played / (roll_ms / 1e3), 1))
                # dominant kernel = the rollouts: SURVEY 8d(1) 64 B per ply x the plies one launch EXECUTES
                alg_bytes = 64.0 * played
                nominal = 64.0 * plies_per_sim * G * args.iterations
                roof_extra = {"achieved_counting_closed_form_plies": round(nominal / (roll_ms / 1e3) / 1e9, 2),
                              "frac_counting_closed_form_plies": round(nominal / (roll_ms / 1e3) / 1e9 / hbm_peak, 4)}
                dominant = (rollout_kernel_name(G * args.iterations) + ": all rollouts of one search", roll_ms)
            else:
                alg_bytes = b_sim * G * args.iterations  # per launch
        elif args.workload == "playout":
            alg_bytes = 64.0 * units / args.steps / world  # 32 B read + 32 B write per ply
            detail.update(bytes_per_ply=64, games_per_sec=round(G * world * args.steps / (dev_ms / 1e3), 1),
                          mean_plies_per_game=round(units / (G * world * args.steps), 2))
        else:
            alg_bytes = 0.0
        traffic, issue = None, None
        if args.workload == "mcts" and dominant and G == 1024 and args.iterations == 100 and args.round_limit == 400:
            if rollout_kernel_name(G * args.iterations).startswith("lane_pack"):
                traffic = ncu_traffic(first_existing("r02_lane_pack_headline_ncu_full_summary.txt"))
                issue = issue_record("r02_lane_pack_headline_issue.json")
            else:
                traffic = ncu_traffic(first_existing("r02_lane_run_rollouts_ncu_full_summary.txt", "r01_lane_run_rollouts_v8_ncu_full_summary.txt"))
                issue = issue_record("r02_lane_run_rollouts_issue.json")
        if args.workload == "playout" and G == 65536 and args.round_limit == 400:
            if rollout_kernel_name(G, playout=True).startswith("lane_pack"):
                traffic = ncu_traffic(first_existing("r02_lane_pack_playouts_ncu_full_summary.txt"))
                issue = issue_record("r02_lane_pack_playouts_issue.json")
            else:
                traffic = ncu_traffic(first_existing("r02_lane_run_playouts_ncu_full_summary.txt", "r01_lane_run_playouts_v8_ncu_full_summary.txt"))
                issue = issue_record("r02_lane_run_playouts_issue.json")
        avg_launch_ms = float(np.mean(kern_ms))
        achieved = alg_bytes / ((dominant[1] if dominant else avg_launch_ms) / 1e3) / 1e9
        roof = None
        if args.workload in ("alpha", "selfplay"):
            burst, sustained, tsrc = tensor_peaks()
            step_s = avg_launch_ms / 1e3
            if args.workload == "alpha":
                evals = G * (args.iterations + 1)
                detail.update(net_evals_per_step=evals, net_evals_per_sec=round(evals / step_s, 1))
            else:
                evals = sp_info["evals"]
                detail.update(net_evals_per_step=int(evals), waves_per_step=sp_info["waves"], records_per_step=sp_info["records"],
                              simulations_per_sec=round(evals / step_s, 1), games_finished_per_step=sp_info["finished"],
                              games_cut_per_step=sp_info["cut"], game_moves_per_sec=round(sp_info["moves"] / step_s, 1),
                              games_per_sec_at_111_moves_per_game=round(sp_info["moves"] / step_s / 111.0, 2),
                              concurrent_games=G * world,
                              mode=("NON-PARITY: " + ", ".join(x for x in (f"slot refill to {args.refill} games" if args.refill else "",
                                                                          f"{args.leaves} leaves per game and step, virtual loss 1" if args.leaves > 1 else "",
                                                                          f"time box {args.max_waves} waves" if args.max_waves else "") if x))
                              if (args.refill or args.leaves > 1 or args.max_waves) else "reference: self_play_parallel, record for record")
            tf = NET_FLOP_PER_EVAL * evals / step_s / 1e12 / (world if args.workload == "selfplay" else 1)  # per GPU
            roof = {"bound": "tensor", "achieved": round(tf, 2), "peak": sustained, "unit": "TFLOP/s", "frac": round(tf / sustained, 4),
                    "traffic": None, "peak_source": tsrc + " (sustained figure: the kernel runs inside a long step)",
                    "kernel": "conv3x3_tc_kernel (tcgen05, 38 of the 42 launches of one forward)",
                    "precision": args.precision,
                    "note": "achieved = 1.0825 GFLOP of USEFUL work per evaluated position (the fp32 model's 2*MAC count) x positions "
                            "/ step time, i.e. the WHOLE step (tree kernels, heads, host sampling included) charged against the "
                            "bf16 tensor peak; split3 spends six bf16 MMAs per product, so its ceiling is 1/6 of that peak"}
        out = {
            "metric": metric, "value": round(value, 1), "unit": unit, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(dev_ms / args.steps, 4), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": dtype, "data": "synthetic",
            "config": workload_config(args),
            "detail": {"timing": "CUDA events per step on the launch stream, max over ranks; L2 flushed (256 MiB fill) between "
                                 "timed steps; after the W warm-up steps the loop keeps stepping untimed until 60 ms have "
                                 "passed (clock ramp, rank start skew)", "extra_warmup_steps": extra_warm,
                       "wall_s": round(t_wall, 3),
                       "parallelism": f"games sharded over {world} GPU(s) by global game id, no data-path collective; "
                                      "the one exchange step (trajectory all-gather, C ABI over NCCL) is timed under "
                                      "detail.alphazero.exchange", **detail},
            "roofline": roof or {"bound": "hbm", "achieved": round(achieved, 4), "peak": hbm_peak, "unit": "GB/s",
                         "frac": round(achieved / hbm_peak, 8), "traffic": traffic, "peak_source": peak_src,
                         "kernel": (dominant[0] if dominant else "mcts_search_kernel<BgGame> (fused rollouts)") if args.workload == "mcts"
                         else rollout_kernel_name(G, playout=True) + ": whole games",
                         **roof_extra, "issue": issue,
                         "note": "achieved = 64 B per ply (32 B state in + 32 B out, SURVEY 8d(1)) x the plies the launch EXECUTES / that "
                                 "kernel's CUDA-event time (plies of a reference-exact rollout after both sides have collected "
                                 "everything are forced passes resolved in closed form: not executed, not counted).  The path is "
                                 "integer-issue bound, not HBM bound: a game lives on the chip (a lane's registers, or shared memory) from its "
                                 "first ply to its last, so the real traffic is 64 B per GAME (`traffic`, ncu); `issue` (from the committed ncu "
                                 "export, tools/ncu_issue.py) is the reading that tracks kernel quality"},
            "e2e": {"value": round(e2e_units / e2e_s, 1), "unit": unit, "h2d_bytes_per_step": h2d * world,
                    "d2h_bytes_per_step": d2h * world},
            "parity_check": parity,
            "gpu_launches": int(launches), "clocks": clocks,
        }
        if exchange:
            out["detail"]["exchange"] = exchange
    if net is not None:
        net.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return out, h_states


def sub_mcts(ctx, ffi, torch, dev, stream, rank, world, G2, cfg, reduce_max, reps=3):
    """one more batch size / mode of the pure-MCTS search, device-resident, CUDA events; value over all ranks"""
    h2 = midgame_states(ctx, ffi, rank * G2, G2)
    d2_states = torch.from_numpy(h2.view(np.uint8).reshape(G2, 32)).to(dev)
    d2_players = torch.from_numpy(h2["player"].copy()).to(dev)
    d2_best = torch.zeros(G2, dtype=torch.int32, device=dev)
    d2_status = torch.zeros(G2, dtype=torch.int32, device=dev)

    def go(i):
        ctx.mcts_search_dev(ffi.GAME_BACKGAMMON, d2_states.data_ptr(), G2, d2_players.data_ptr(), cfg, SEED, rank * G2, i,
                            d2_best.data_ptr(), d2_status.data_ptr(), 0)
    go(0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for i in range(reps):
        go(1 + i)
    e1.record(stream)
    e1.synchronize()
    ms2 = reduce_max(e0.elapsed_time(e1))
    assert int(d2_status.abs().sum().item()) == 0
    its = int(cfg["iterations"][0])
    out = {"games_per_gpu": G2, "value": round(reps * G2 * world * its / (ms2 / 1e3), 1), "unit": "simulations/s",
           "ms_per_search": round(ms2 / reps, 4)}
    if not (int(cfg["mode_flags"][0]) & ffi.MODE_ROLLOUT_CHECK_CURRENT):
        # which rollout kernel ran (lane_kernels.cu: jobs of >= 640 rollouts per SM go to the packed kernel), its share of
        # the search, and that the other kernel finds the same moves on this very batch
        try:
            tree_ms, roll_ms = ctx.search_timing()
            plies = ctx.search_work()
            packed = rollout_kernel_name(G2 * its).startswith("lane_pack")
            best_timed = d2_best.cpu().numpy().tobytes()
            old_env = os.environ.get("DIEE_LANE_PACK")
            os.environ["DIEE_LANE_PACK"] = "0" if packed else "2"
            go(reps)
            torch.cuda.synchronize()
            same = d2_best.cpu().numpy().tobytes() == best_timed
            if old_env is None:
                del os.environ["DIEE_LANE_PACK"]
            else:
                os.environ["DIEE_LANE_PACK"] = old_env
            out.update({"rollout_kernel": "lane_pack_kernel" if packed else "lane_run_kernel", "tree_ms": round(tree_ms, 4),
                        "rollout_ms": round(roll_ms, 4), "rollout_plies_played": plies,
                        "rollout_gplies_per_s": round(plies / roll_ms / 1e6, 2),
                        "same_best_moves_with_the_other_rollout_kernel": bool(same)})
            if packed and G2 == 8192:  # the committed ncu export of this very job (tools/pack_stats.py 8192, tools/ncu_issue.py)
                out["issue"] = issue_record("r02_lane_pack_rollouts_issue.json")
            if not same:
                raise SystemExit("bench: the packed and the lane-resident rollout kernels disagree at %d games" % G2)
        except ffi.DieeError:
            pass
    return out


def exchange_step(ctx, ffi, dist, dev, rank, world, recs, barrier):
    """the one exchange step of the path: every rank's finished trajectories all-gathered into every
    rank's replay buffer through the C ABI (diee_comm_init / diee_traj_allgather, NCCL over NVLink)"""
    import torch
    uid = [ffi.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    ctx.comm_init(world, rank, uid[0])
    rec, pi_ids, pi_vals = recs
    n = torch.tensor([len(rec), len(pi_ids)], device=dev, dtype=torch.int64)
    dist.all_reduce(n, op=dist.ReduceOp.SUM)
    tot_rec, tot_pi = int(n[0].item()), int(n[1].item())
    ctx.traj_allgather(rec, pi_ids, pi_vals, tot_rec + 8, tot_pi + 8)  # warm-up (communicator set-up, buffers)
    barrier()
    t0 = time.perf_counter()
    g_rec, g_ids, g_vals = ctx.traj_allgather(rec, pi_ids, pi_vals, tot_rec + 8, tot_pi + 8)
    barrier()
    dt = time.perf_counter() - t0
    ctx.comm_destroy()
    nbytes = tot_rec * 48 + tot_pi * 6
    # rank-major: this rank's own records sit behind those of the lower ranks, bit for bit (pi_offset is rebased)
    cnt = torch.zeros(world, device=dev, dtype=torch.int64)
    cnt[rank] = len(rec)
    dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    off = int(cnt[:rank].sum().item())
    mine = g_rec[off:off + len(rec)]
    ok = (len(g_rec) == tot_rec and len(g_ids) == tot_pi and mine["state"].tobytes() == rec["state"].tobytes()
          and (mine["game_id"] == rec["game_id"]).all() and (mine["outcome"] == rec["outcome"]).all())
    return {"collective": "ncclAllGather x3 (counts, status, padded slabs) via diee_traj_allgather", "ranks": world,
            "records": tot_rec, "pi_entries": tot_pi, "bytes_received_per_rank": nbytes, "seconds": round(dt, 6),
            "gb_per_s_per_rank": round(nbytes / dt / 1e9, 3), "ok": bool(ok),
            "note": "host buffers in, host buffers out: the time includes staging H2D, the collective and D2H"}


def sub_alpha(ctx, ffi, torch, dev, stream, rank, world, args, h_states, reduce_max, reduce_sum, barrier, dist):
    """BASELINE configs[3] inside the default line: one alpha_mcts_parallel of the 1,024 games per precision, and a
    time-boxed self_play_parallel (diee_selfplay_run_ex, max_waves) whose records go through the exchange step"""
    from die_e_b200 import nnet
    G = len(h_states)
    first_gid = rank * G
    burst, sustained, tsrc = tensor_peaks()
    cfg = np.zeros(1, dtype=ffi.MCTS_CFG)
    cfg[0] = (args.iterations, 2.0, args.round_limit, 0.3, 0.25, 0)
    net = ffi.Net(ctx, nnet.synthetic_tensors(seed=SEED, filters=256, blocks=19, bn_stats="identity"))
    h_ids = np.arange(first_gid, first_gid + G, dtype=np.uint32)
    d_states = torch.from_numpy(h_states.view(np.uint8).reshape(G, 32)).to(dev)
    d_ids = torch.from_numpy(h_ids.view(np.int32)).to(dev)
    d_rids = torch.zeros(G, ffi.MAX_MOVES, dtype=torch.int16, device=dev)
    d_rmoves = torch.zeros(G, ffi.MAX_MOVES, dtype=torch.int32, device=dev)
    d_rvis = torch.zeros(G, ffi.MAX_MOVES, dtype=torch.float32, device=dev)
    d_rcnt = torch.zeros(G, dtype=torch.int32, device=dev)
    d_status = torch.zeros(G, dtype=torch.int32, device=dev)
    out = {"net": "ResNet 256 filters x 19 blocks, synthetic weights", "games_per_gpu": G, "search": {}, "selfplay": {}}
    for prec in ("split3", "bf16"):
        net.set_precision(PRECISIONS[prec])

        def go(i):
            ctx.alpha_search_dev(net, d_states.data_ptr(), G, d_ids.data_ptr(), cfg, SEED, i & 0xFFFF, 0, d_rids.data_ptr(),
                                 d_rmoves.data_ptr(), d_rvis.data_ptr(), d_rcnt.data_ptr(), d_status.data_ptr())
        go(0)
        torch.cuda.synchronize()
        reps = 2 if prec == "bf16" else 1
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for i in range(reps):
            go(1 + i)
        e1.record(stream)
        e1.synchronize()
        ms = reduce_max(e0.elapsed_time(e1)) / reps
        assert int(d_status.abs().sum().item()) == 0
        evals = G * (args.iterations + 1) * world
        tf = NET_FLOP_PER_EVAL * evals / (ms / 1e3) / 1e12
        out["search"][prec] = {"value": round(G * world * args.iterations / (ms / 1e3), 1), "unit": "simulations/s",
                               "ms_per_search": round(ms, 3), "net_evals_per_sec": round(evals / (ms / 1e3), 1),
                               "tensor_tflops_useful": round(tf, 2), "frac_of_sustained_bf16_peak": round(tf / sustained / world, 4),
                               "dtype": PRECISION_DTYPE[prec]}
        # time-boxed self-play from the opening: `waves` game-moves of every game
        waves = 6 if prec == "bf16" else 2
        barrier()
        e0n = ctx.net_eval_count()
        t0 = time.perf_counter()
        rec, pi_ids, pi_vals, rep = ctx.selfplay_run_ex(net, G, cfg, 1.25, SEED, first_gid, max_waves=waves)
        dt = reduce_max(time.perf_counter() - t0)
        ev = reduce_sum(ctx.net_eval_count() - e0n)
        moves = reduce_sum(int(rep["game_moves"]))
        tf = NET_FLOP_PER_EVAL * ev / dt / 1e12
        out["selfplay"][prec] = {"time_box_waves": waves, "game_moves_per_sec": round(moves / dt, 1),
                                 "simulations_per_sec": round(moves * args.iterations / dt, 1), "seconds": round(dt, 3),
                                 "records": int(reduce_sum(len(rec))), "tensor_tflops_useful": round(tf, 2),
                                 "frac_of_sustained_bf16_peak": round(tf / sustained / world, 4),
                                 "games_per_sec_at_111_moves_per_game": round(moves / dt / 111.0, 2),
                                 "note": "time-boxed sample of self_play_parallel through the host call (diee_selfplay_run_ex, "
                                         "max_waves): whole-run games/s are in profiles/ (bench.py --workload selfplay); the "
                                         "estimate divides game-moves/s by the mean game length of those runs"}
        if prec == "split3" and world > 1:
            out["exchange"] = exchange_step(ctx, ffi, dist, dev, rank, world, (rec, pi_ids, pi_vals), barrier)
    out["tensor_peak"] = {"sustained_tflops": sustained, "source": tsrc}
    out["precision_note"] = ("the reference computes in fp32 (lib.rs:20): bf16 is the fast NON-parity mode; split3 is the "
                             "tensor-core mode inside the fp32 tolerance and the library's default (error vs fp64 at 19 blocks: 4.9e-6 "
                             "value / 6.7e-6 policy, torch fp32 itself 8.1e-6 / 2.5e-6, tests/test_gpu_net.py)")
    net.close()
    return out


# ---------------------------------------------------------------- the CPU arm (the oracle = the reference's algorithm in C)
def cpu_alpha_leg(args, h_states, threads):
    """the reference's AlphaZero search on the host: the C oracle's single-threaded lock-step loop
    (alpha_mcts.rs:153) around a PyTorch-CPU fp32 forward of the same ResNet (all cores)"""
    import orc
    import net_oracle
    import torch
    from die_e_b200 import nnet
    torch.set_num_threads(threads)
    tens = nnet.synthetic_tensors(seed=SEED, filters=256, blocks=19, bn_stats="identity")

    def ev(st):
        x = np.concatenate([orc.bg_as_tensor(st[i:i + 1]) for i in range(len(st))])
        p, v = net_oracle.forward(tens, x, 19, dtype=torch.float32)
        return p.astype(np.float32), v.astype(np.float32)
    cb = orc.make_eval(ev)
    n = min(len(h_states), 16)
    iters = min(args.iterations, 20)  # bounded sample: the per-simulation cost does not depend on the iteration count
    cfg = orc.mcts_cfg(iters, 2.0, args.round_limit, 0.3, 0.25, 0)
    t0 = time.perf_counter()
    orc.alpha_mcts_parallel(h_states[:n], np.arange(n), cfg, SEED, 0, cb, 1 + (iters + 1) * 128)
    dt = time.perf_counter() - t0
    return n * iters / dt, f"games 0..{n - 1} of the {len(h_states)} x {iters} of the {args.iterations} iterations, {dt:.1f} s", dt, n


def cpu_sample_size(args, h_states, seconds_target, threads, mode):
    """how many games (a prefix of the GPU arm's batch) fill about `seconds_target` on `threads` host threads"""
    import orc
    if args.workload == "mcts":
        cfg = orc.mcts_cfg(args.iterations, 2.0, args.round_limit, 0.3, 0.25, mode)
        n0 = min(len(h_states), 2 * threads)
        t0 = time.perf_counter()
        orc.mcts_search_bg_batch(h_states[:n0], h_states["player"][:n0].copy(), cfg, SEED, 0, 0, threads)
    else:
        n0 = min(len(h_states), 64 * threads)
        t0 = time.perf_counter()
        orc.bg_playout_batch(h_states[:n0], SEED, 0, args.round_limit, threads)
    dt = time.perf_counter() - t0
    return int(max(threads, min(len(h_states), n0 * seconds_target / max(dt, 1e-3))))


def cpu_run(args, h_states, n, threads, mode):
    """one timed pass of the oracle over games 0..n-1 of the batch, one task per game over `threads` host threads"""
    import orc
    t0 = time.perf_counter()
    if args.workload == "mcts":
        cfg = orc.mcts_cfg(args.iterations, 2.0, args.round_limit, 0.3, 0.25, mode)
        orc.mcts_search_bg_batch(h_states[:n], h_states["player"][:n].copy(), cfg, SEED, 0, 0, threads)
        units = n * args.iterations
    else:
        _, plies = orc.bg_playout_batch(h_states[:n], SEED, 0, args.round_limit, threads)
        units = float(plies.sum())
    dt = time.perf_counter() - t0
    return units / dt, dt


def cpu_leg(args, h_states, seconds_target, threads, mode, reps=3):
    """the oracle (C restatement of the reference's CPU path) on a bounded sample of the GPU arm's own inputs:
    `reps` repetitions at all host threads and at half of them (the reference's default pool, main.rs:100-106);
    best and median of each"""
    load_oracle()
    if args.workload in ("alpha", "selfplay"):
        v, sample, dt, n = cpu_alpha_leg(args, h_states, threads)
        return {"value": round(v, 2), "median": round(v, 2), "cores": threads, "kind": "port", "sample": sample, "reps": [round(v, 2)],
                "seconds": round(dt, 1)}
    per = max(0.5, seconds_target / (2.0 * reps))
    out = {}
    t_all = time.perf_counter()
    for label, th in (("full", threads), ("half", max(1, threads // 2))):
        n = cpu_sample_size(args, h_states, per, th, mode)
        vals = [cpu_run(args, h_states, n, th, mode)[0] for _ in range(reps)]
        out[label] = (th, n, vals)
    th, n, vals = out["full"]
    th2, n2, vals2 = out["half"]
    what = "games x %d iterations" % args.iterations if args.workload == "mcts" else "games"
    return {"value": round(max(vals), 1), "median": round(float(np.median(vals)), 1), "cores": th, "kind": "port",
            "sample": f"games 0..{n - 1} of the GPU arm's own {len(h_states)} {what} (same seed, ids, inputs), {reps} repetitions",
            "reps": [round(v, 1) for v in vals],
            "half_cores": {"cores": th2, "value": round(max(vals2), 1), "median": round(float(np.median(vals2)), 1),
                           "games": n2, "reps": [round(v, 1) for v in vals2]},
            "seconds": round(time.perf_counter() - t_all, 1),
            "note": "C oracle restating the reference's mct_search, one task per game over a pthread pool (= rayon par_iter, "
                    "versus.rs:303-306); the Rust reference cannot be built here; the C port omits its deep clones and "
                    "per-candidate allocations, so it is a stronger baseline than the real reference"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="mcts", choices=["mcts", "playout", "alpha", "selfplay"])
    ap.add_argument("--games", type=int, default=None)
    ap.add_argument("--iterations", type=int, default=100)
    ap.add_argument("--round-limit", type=int, default=400)
    ap.add_argument("--rollout", default="ref_exact", choices=["ref_exact", "check_current"])
    ap.add_argument("--precision", default="split3", choices=list(PRECISIONS))
    ap.add_argument("--cpu-seconds", type=float, default=18.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-subrecords", "--no-large-batch", dest="no_subrecords", action="store_true")
    ap.add_argument("--no-parity-check", action="store_true")
    ap.add_argument("--refill", type=int, default=0, help="selfplay: NON-PARITY slot refill, stop after this many finished games")
    ap.add_argument("--leaves", type=int, default=0, help="selfplay/alpha: NON-PARITY leaves per game and step (virtual loss 1)")
    ap.add_argument("--max-waves", type=int, default=0, help="selfplay: time box (game-move waves)")
    args = ap.parse_args()
    if args.games is None:
        args.games = 65536 if args.workload == "playout" else 1024
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    threads = os.cpu_count() or 1
    mode = 2 | (1 if args.rollout == "check_current" else 0)

    if args.impl == "reference":
        if rank != 0:
            return
        load_oracle()
        # the same inputs as rank 0 of the GPU arm; a bounded prefix of them per step
        h_states = host_states_for_reference(args, 0, min(args.games, 1024 if args.workload != "playout" else 16384))
        per_step = max(1.0, min(20.0, 90.0 / max(1, args.steps + args.warmup)))
        unit = {"mcts": "simulations/s", "alpha": "simulations/s", "selfplay": "simulations/s", "playout": "plies/s"}[args.workload]
        t_all0 = time.perf_counter()
        vals = []
        if args.workload in ("alpha", "selfplay"):
            for i in range(args.warmup + args.steps):
                v, sample, dt, n = cpu_alpha_leg(args, h_states, threads)
                if i >= args.warmup:
                    vals.append((v, dt))
        else:
            n = cpu_sample_size(args, h_states, per_step, threads, mode)
            what = "games x %d iterations" % args.iterations if args.workload == "mcts" else "games"
            sample = f"games 0..{n - 1} of the GPU arm's own {args.games} {what} per step (same seed, ids, inputs)"
            for i in range(args.warmup + args.steps):
                v, dt = cpu_run(args, h_states, n, threads, mode)
                if i >= args.warmup:
                    vals.append((v, dt))
        tot_units = sum(v * dt for v, dt in vals)
        tot_s = sum(dt for _, dt in vals)
        value = tot_units / tot_s
        rates = [v for v, _ in vals]
        line = {
            "impl": "reference", "metric": {"mcts": "mcts_simulations_per_sec", "alpha": "alphazero_simulations_per_sec",
                                            "selfplay": "alphazero_simulations_per_sec", "playout": "playout_plies_per_sec"}[args.workload],
            "value": round(value, 1), "unit": unit, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(1e3 * tot_s / max(1, args.steps), 2), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int8 boards / f32 UCB" if args.workload in ("mcts", "playout") else "fp32 net (torch CPU) / f32 PUCT",
            "data": "synthetic", "config": workload_config(args),
            "detail": {"wall_s": round(time.perf_counter() - t_all0, 1), "each_step": "one pass of the oracle over a bounded prefix of the "
                       "GPU arm's batch on the host CPU"},
            "cpu_baseline": {"value": round(value, 1), "best": round(max(rates), 1), "median": round(float(np.median(rates)), 1),
                             "unit": unit, "cores": threads, "kind": "port", "sample": sample,
                             "note": "C oracle restating the reference's mct_search, one task per game over a pthread pool "
                                     "(= rayon par_iter, versus.rs:303-306); the Rust reference cannot be built here; the C "
                                     "port omits its deep clones and per-candidate allocations, so it is a stronger baseline"},
            "e2e": {"value": round(value, 1), "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
        }
        print(json.dumps(line), flush=True)
        return

    out, h_states = run_ours(args, rank, world)
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            cb = cpu_leg(args, h_states, args.cpu_seconds, threads, mode)
            if args.workload == "selfplay":  # games/s extrapolated from the measured simulations/s
                sims_per_game = args.iterations * out["detail"]["waves_per_step"]
                cb["sample"] += f"; {cb['value']:.1f} simulations/s extrapolated to games/s with {sims_per_game} simulations per game-slot"
                cb["value"] = round(cb["value"] / sims_per_game, 6)
                cb["median"] = round(cb["median"] / sims_per_game, 6)
            cb["unit"] = out["unit"]
            out["cpu_baseline"] = cb
        else:
            out["cpu_baseline"] = None
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
