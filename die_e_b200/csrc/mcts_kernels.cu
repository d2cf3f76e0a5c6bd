// mcts_kernels.cu -- pure MCTS: one warp runs one game's whole search.
//
// Reference: src/mcts/simple_mcts.rs (mct_search :10-39, select_ucb :41-52, select_leaf_node
// :88-94, backpropagate :96-103, select_most_visits :71-86), src/mcts/node.rs (ucb :86-96,
// expand :118-137, simulate :176-196), src/mcts/node_store.rs (arena).
//
// Node pool: structure-of-arrays in HBM, one contiguous slab of (iterations+1) nodes per game
// (states / parent / visits / value / action / move counts), so a warp's select pass reads
// coalesced runs of parent[], visits[], value[].  A node's children are exactly the nodes whose
// parent[] equals it, in creation order, so UCB select is a warp scan over the slab with a
// "later index wins ties" arg-max (Rust's max_by keeps the LAST maximum).  UCB is evaluated in
// IEEE f32 with explicit round-to-nearest intrinsics (no FMA contraction, no fast division) and
// ln(parent visits) comes from a host-built table of (float)log((double)n): visits are
// integer-valued, so this is bit-identical to the oracle.
#include <cstdlib>
#include <type_traits>

#include "bg_device.cuh"
#include "bg_lane.cuh"
#include "lane_pack.cuh"
#include "launchers.h"

namespace diee {

constexpr int MCTS_WARPS_PER_CTA = 4;
constexpr int NO_WINNER = 2;
#ifndef DIEE_TREE_MINB
#define DIEE_TREE_MINB 5  // resident CTAs per SM the many-wave instance of the tree kernel is compiled for (see mcts_search_kernel)
#endif
constexpr int NODE_PLAYS = 4;                 // plays kept per node from the call that counted them (the ones expand pops first)
constexpr int ROOT_PLAYS = 128;               // cached plays of the root (a backgammon position has at most ~130)
constexpr uint32_t NM_UNKNOWN = 0xFFFFFFFFu;  // a node whose legal moves have not been counted yet

// ---------------- game policies ----------------
// The warp's board as the lane engine's bit planes (warp-uniform: every lane holds the whole board).
__device__ __forceinline__ void bg_to_planes(const BgWarp &g, lane::LaneBoard &b) {
    const int a = abs(g.v);
    uint32_t neg[4], pos[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        neg[k] = __ballot_sync(FULL, g.v < 0 && ((a >> k) & 1)) & M24;
        pos[k] = __ballot_sync(FULL, g.v > 0 && ((a >> k) & 1)) & M24;
    }
    if (g.player < 0) {
#pragma unroll
        for (int k = 0; k < 4; ++k) { b.own[k] = neg[k]; b.opp[k] = pos[k]; }
        b.bar_own = g.bar0; b.bar_opp = g.bar1; b.off_own = g.off0; b.off_opp = g.off1;
    } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) { b.own[k] = lane::l_rev24(pos[k]); b.opp[k] = lane::l_rev24(neg[k]); }
        b.bar_own = g.bar1; b.bar_opp = g.bar0; b.off_own = g.off1; b.off_opp = g.off0;
    }
    b.roll0 = g.roll0; b.roll1 = g.roll1; b.player = g.player; b.second = g.second;
}

// Number of legal plays of the warp's game and (k >= 0) the k-th of them, as (overflow << 63 | U << 32 | play).
// Contact play and bar entries are counted in closed form by the lane engine (bg_lane.cuh) -- ~10x fewer
// instructions than building the list -- redundantly on every lane; positions in the bear-off regime build
// the list cooperatively.  One out-of-line copy (the tree kernel calls it three times per iteration and is
// instruction-fetch bound); everything travels BY VALUE in registers: a reference to the caller's board
// would force it into local memory and every call would start with a round trip through it.
// k >= 0: the k-th play; k == -1: count only; k == -2: the play at index_of(w, U) (a rollout's uniform choice);
// k <= -3: the play -k-2 from the end (-3 = the last one).  k may differ from lane to lane.
__device__ __noinline__ unsigned long long bg_count_and_kth(int v, uint32_t scal, WarpSlab *slab, int lane, int k, uint32_t w,
                                                            const uint32_t *pb_index, const uint16_t *pb_plays) {
    BgWarp g;
    g.v = v;
    g.bar0 = (int)(scal & 15u); g.bar1 = (int)((scal >> 4) & 15u); g.off0 = (int)((scal >> 8) & 15u); g.off1 = (int)((scal >> 12) & 15u);
    g.roll0 = (int)((scal >> 16) & 15u); g.roll1 = (int)((scal >> 20) & 15u);
    g.player = ((scal >> 24) & 1u) ? 1 : -1;
    g.second = (int)((scal >> 25) & 1u);
    lane::LaneBoard b;
    bg_to_planes(g, b);
    const int hi = max(g.roll0, g.roll1), lo = min(g.roll0, g.roll1);
    lane::LaneMasks m;
    const bool closed = lane::l_closed_applies(b, m, lo, hi);
    uint32_t seq = SEQ_EMPTY;
    if (closed || b.bar_own > 0 || m.own1 == 0) {
        int U = 0;
        lane::LanePlay pl;
        pl.n = 0; pl.x1 = pl.t1 = pl.x2 = pl.t2 = 0;
        if (b.bar_own > 0 || m.own1 != 0) U = lane::l_contact_select(b, m, lo, hi, k, w, pl);
        if (pl.n > 0) seq = lane::l_play_to_seq(pl, g.player);
        return ((unsigned long long)(uint32_t)U << 32) | seq;
    }
    if (lane::l_pure_bearoff(b)) {  // every checker home, no opposing checker there: the play table
        const uint32_t e = __ldg(pb_index + lane::l_pb_key(b));
        const int U = (int)(e & 255u);
        if (k == -2 && U > 0) k = (int)index_of(w, (uint32_t)U);
        else if (k < -2) k = k + U + 2 >= 0 ? k + U + 2 : -1;
        if (k >= 0 && k < U) seq = lane::l_play_to_seq(lane::l_pb_unpack(__ldg(pb_plays + (e >> 8) + k)), g.player);
        return ((unsigned long long)(uint32_t)U << 32) | seq;
    }
    bool ovf = false;
    const int U = bg_movegen(g, *slab, lane, ovf);
    if (k == -2 && U > 0) k = (int)index_of(w, (uint32_t)U);
        else if (k < -2) k = k + U + 2 >= 0 ? k + U + 2 : -1;
    if (k >= 0 && k < U) seq = slab->raw[k];
    __syncwarp();
    return ((unsigned long long)ovf << 63) | ((unsigned long long)(uint32_t)U << 32) | seq;
}

struct BgGame {
    using State = diee_bg_state;
    BgWarp g;
    const uint32_t *pb_index;
    const uint16_t *pb_plays;
    __device__ __forceinline__ void attach(const PbTable &t) { pb_index = t.index; pb_plays = t.plays; }
    __device__ __forceinline__ int count_and_kth(WarpSlab &slab, int lane, bool &ovf, int k, uint32_t &seq) const {
        const uint32_t scal = (uint32_t)g.bar0 | ((uint32_t)g.bar1 << 4) | ((uint32_t)g.off0 << 8) | ((uint32_t)g.off1 << 12) |
                              ((uint32_t)g.roll0 << 16) | ((uint32_t)g.roll1 << 20) | ((g.player > 0 ? 1u : 0u) << 24) |
                              ((uint32_t)(g.second ? 1 : 0) << 25);
        const unsigned long long r = bg_count_and_kth(g.v, scal, &slab, lane, k, 0u, pb_index, pb_plays);
        if (r >> 63) ovf = true;
        if (k >= 0 || k < -2) seq = (uint32_t)r;
        return (int)((r >> 32) & 0x7FFFFFFFu);
    }
    // the play a rollout makes: uniform over the legal plays with the word w (SEQ_EMPTY when there is none)
    __device__ __forceinline__ uint32_t random_play(WarpSlab &slab, int lane, bool &ovf, uint32_t w) const {
        const uint32_t scal = (uint32_t)g.bar0 | ((uint32_t)g.bar1 << 4) | ((uint32_t)g.off0 << 8) | ((uint32_t)g.off1 << 12) |
                              ((uint32_t)g.roll0 << 16) | ((uint32_t)g.roll1 << 20) | ((g.player > 0 ? 1u : 0u) << 24) |
                              ((uint32_t)(g.second ? 1 : 0) << 25);
        const unsigned long long r = bg_count_and_kth(g.v, scal, &slab, lane, -2, w, pb_index, pb_plays);
        if (r >> 63) ovf = true;
        return (uint32_t)r;
    }
    __device__ __forceinline__ void load(const State *s, int lane) { bg_load(g, s, lane); }
    __device__ __forceinline__ void store(State *s, int lane) const { bg_store(g, s, lane); }
    __device__ __forceinline__ int winner() const { const int w = bg_winner(g); return w == 0 ? NO_WINNER : w; }
    __device__ __forceinline__ int movegen(WarpSlab &slab, int lane, bool &ovf) const { return bg_movegen(g, slab, lane, ovf); }
    __device__ __forceinline__ uint32_t move_at(const WarpSlab &slab, int k) const { return slab.raw[k]; }
    __device__ __forceinline__ void step(uint32_t seq, int d0, int d1, int lane) { bg_step(g, seq, d0, d1, lane); }
};

struct TttGame {  // tictactoe/mod.rs; the whole state is warp-uniform
    using State = diee_ttt_state;
    __device__ __forceinline__ void attach(const PbTable &) {}
    uint32_t xm, om;  // cells held by -1 / +1
    int player;
    __device__ __forceinline__ void load(const State *s, int lane) {
        const int b = lane < 9 ? (int)s->board[lane] : 0;
        xm = __ballot_sync(FULL, b == -1);
        om = __ballot_sync(FULL, b == 1);
        player = s->player;
    }
    __device__ __forceinline__ void store(State *s, int lane) const {
        if (lane < 9) s->board[lane] = ((xm >> lane) & 1u) ? -1 : (((om >> lane) & 1u) ? 1 : 0);
        if (lane == 9) s->player = (int8_t)player;
        if (lane >= 10 && lane < 16) s->pad[lane - 10] = 0;
    }
    // check_winner :59-79: a cell belongs to one side, so at most one side has a line unless the
    // position is unreachable; the table order (rows, cols, diags) decides then.
    __device__ __forceinline__ int winner() const {
        const uint32_t L[8] = {0x007u, 0x038u, 0x1C0u, 0x049u, 0x092u, 0x124u, 0x111u, 0x054u};
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            if ((xm & L[k]) == L[k]) return -1;
            if ((om & L[k]) == L[k]) return 1;
        }
        return ((xm | om) & 0x1FFu) == 0x1FFu ? 0 : NO_WINNER;
    }
    __device__ __forceinline__ int movegen(WarpSlab &, int, bool &) const { return __popc(~(xm | om) & 0x1FFu); }
    __device__ __forceinline__ int count_and_kth(WarpSlab &slab, int, bool &, int k, uint32_t &seq) const {
        const int U = __popc(~(xm | om) & 0x1FFu);
        if (k < -2) k += U + 2;
        if (k >= 0 && k < U) seq = move_at(slab, k);
        return U;
    }
    __device__ __forceinline__ uint32_t random_play(WarpSlab &slab, int, bool &, uint32_t w) const {
        const int U = __popc(~(xm | om) & 0x1FFu);
        return U > 0 ? move_at(slab, (int)index_of(w, (uint32_t)U)) : SEQ_EMPTY;
    }
    __device__ __forceinline__ uint32_t move_at(const WarpSlab &, int k) const {  // k-th empty cell ascending :36-44
        uint32_t e = ~(xm | om) & 0x1FFu;
        for (int i = 0; i < k; ++i) e &= e - 1;
        return (uint32_t)(__ffs(e) - 1) | 0xFEFEFE00u;
    }
    __device__ __forceinline__ void step(uint32_t seq, int, int, int) {
        if (seq != SEQ_EMPTY) {  // apply_move :46-49
            const uint32_t bit = 1u << (seq & 0xFFu);
            if (player < 0) xm |= bit; else om |= bit;
        }
        player = -player;  // skip_turn :51-53
    }
};

struct Pool {
    void *states;
    int32_t *parent;
    float *visits, *value;
    uint32_t *action;
    uint32_t *nmoves;  // n_moves << 16 | n_untried
    int32_t *n_nodes;  // per game
    int32_t *sim_node; // [game][iteration]: node whose rollout the rollout kernel still has to run, or -1
    void *finals;      // [game][iteration]: state each simulation's rollout ended in (zero = no rollout)
    PbTable pb;        // pure bear-off play table (backgammon)
    float *roll_result; // [game]: result of the rollout the game is waiting for (lock-step search)
};

// lock-step search: did this iteration defer its rollout (warp-uniform; sim_node[it] was just written by lane 0)
__device__ __forceinline__ bool sim_node_deferred(const int32_t *sim_node, uint32_t it, int lane) {
    __syncwarp();
    int v = 0;
    if (lane == 0) v = sim_node[it];
    return __shfl_sync(FULL, v, 0) >= 0;
}

__device__ __forceinline__ float outcome(int winner, int player) {  // simple_mcts.rs:26-28
    return winner == player ? 1.0f : (winner == -player ? -1.0f : 0.0f);
}

// The rollout body shared by the fused and the split forms: `limit` plies of Node::simulate's loop
// (node.rs:180-194) from the state held in `game`.  Returns the result; counts plies.
template <class G>
__device__ __forceinline__ float rollout_plies(G &game, WarpSlab &slab, int lane, bool &ovf, const diee_mcts_cfg &cfg,
                                               bool check_current, int player, uint64_t seed, uint32_t gid, uint32_t c3,
                                               unsigned long long &plies) {
    PhiloxLanes rng;
    for (uint32_t k = 0; k < cfg.simulate_round_limit; ++k) {
        if (check_current) {
            const int wc = game.winner();
            if (wc != NO_WINNER) return outcome(wc, player);
        }
        if ((k & 31u) == 0) rng.fill(seed, k, gid, DIEE_STREAM_ROLLOUT, c3, lane);
        const int src = (int)(k & 31u);
        const int d0 = die_of(__shfl_sync(FULL, rng.w0, src));
        const int d1 = die_of(__shfl_sync(FULL, rng.w1, src));
        const uint32_t w2 = __shfl_sync(FULL, rng.w2, src);
        const uint32_t sq = game.random_play(slab, lane, ovf, w2);
        game.step(sq, d0, d1, lane);
        ++plies;
    }
    return 0.f;
}

// One game's share of a launch: iterations [it_begin, it_end) of its search (see the kernel below for the forms).  `mine` is
// the game's node slab in shared memory, or null when the search works on the pool in HBM / L2 directly.  A device function
// so that the persistent check-current search (cc_search_kernel) can run tree steps from inside its own loop.
template <class G, bool SPLIT, bool LOCK>
__device__ __forceinline__ void mcts_game_body(const int gidx, const int lane, WarpSlab &slab, uint32_t *root_plays, unsigned char *mine,
                                               const typename G::State *__restrict__ roots, const int8_t *__restrict__ players,
                                               const diee_mcts_cfg &cfg, uint32_t it_begin, uint32_t it_end, uint64_t seed,
                                               uint32_t first_game_id, uint32_t epoch, const Pool &pool, const float *__restrict__ ln_table,
                                               uint32_t *__restrict__ best_out, int32_t *__restrict__ status_out,
                                               diee_search_stats *__restrict__ stats_out, bool fill_counts) {
    const bool slab_in_smem = mine != nullptr;
    const int cap = (int)cfg.iterations + 1;
    const size_t base = (size_t)gidx * cap;
    // The game's node slab.  While the kernel runs it lives in SHARED memory when it fits (the search is a
    // chain of dependent reads of parent / visits / value / move counts / states -- a few dozen cycles each
    // from shared memory, several hundred from L2) and is written back to the HBM pool, coalesced, at the
    // end; otherwise the kernel works on the pool directly.
    typename G::State *gst = reinterpret_cast<typename G::State *>(pool.states) + base;
    int32_t *gparent = pool.parent + base;
    float *gvisits = pool.visits + base, *gvalue = pool.value + base;
    uint32_t *gnm = pool.nmoves + base;
    typename G::State *st = gst;
    int32_t *parent = gparent;
    float *visits = gvisits, *value = gvalue;
    uint32_t *nm = gnm;
    // first | last << 16 child index of every node (shared-memory slabs only): select_ucb then scans just the
    // 32-node chunks that can hold children of the node instead of the whole slab
    uint32_t *crange = nullptr;
    // value / visits of every node, refreshed where the two change (back-propagation): select_ucb then needs one
    // IEEE division per child instead of two (same operands, same rounding, so the same bits)
    float *qv = nullptr;
    // the last NODE_PLAYS plays of every node counted by this launch: a node of a 100-iteration search is expanded
    // a handful of times, and each time would otherwise regenerate its move set to pop one play
    uint32_t *nplays = nullptr;
    if (slab_in_smem) {
        st = reinterpret_cast<typename G::State *>(mine);
        parent = reinterpret_cast<int32_t *>(mine + (size_t)cap * sizeof(typename G::State));
        visits = reinterpret_cast<float *>(parent + cap);
        value = visits + cap;
        nm = reinterpret_cast<uint32_t *>(value + cap);
        crange = nm + cap;
        qv = reinterpret_cast<float *>(crange + cap);
        nplays = reinterpret_cast<uint32_t *>(qv + cap);
    }
    uint32_t *action = pool.action + base;
    int32_t *sim_node = pool.sim_node + (size_t)gidx * cfg.iterations;
    typename G::State *finals = reinterpret_cast<typename G::State *>(pool.finals) + (size_t)gidx * cfg.iterations;
    const int player = players[gidx];
    const uint32_t gid = first_game_id + (uint32_t)gidx;
    const bool check_current = cfg.mode_flags & DIEE_MODE_ROLLOUT_CHECK_CURRENT;
    const bool pass_child = cfg.mode_flags & DIEE_MODE_PASS_CHILD;

    // The search may be cut into slices of iterations [it_begin, it_end) (one launch each, so that the
    // rollouts of a slice can run beside the tree work of the next): everything a later slice needs is in
    // the pool; n_nodes == 0 marks a terminal root.
    G game;
    game.attach(pool.pb);
    uint32_t best = SEQ_EMPTY;
    int status = DIEE_OK;
    int n_nodes = 0;
    unsigned long long plies = 0;
    uint32_t sel_levels = 0, sel_children = 0, terminal_leaves = 0;
    bool ovf = false;
    const bool last_slice = LOCK ? it_begin >= cfg.iterations : it_end >= cfg.iterations;

    if (it_begin == 0) {
        game.load(roots + gidx, lane);
        if (game.winner() == NO_WINNER) {  // simple_mcts.rs:12-14
            // root = add_node(state)  (Node::new computes the legal moves eagerly, node.rs:50)
            game.store(st, lane);
            uint32_t play = SEQ_EMPTY;
            int U = game.count_and_kth(slab, lane, ovf, lane, play);
            root_plays[lane] = play;
            if (U == 0 && pass_child) U = 1;
            if (lane == 0) {
                parent[0] = -1; visits[0] = 0.f; value[0] = 0.f; action[0] = SEQ_EMPTY; nm[0] = ((uint32_t)U << 16) | (uint32_t)U;
                if (crange) crange[0] = 0xFFFFu;
            }
            n_nodes = 1;
            __syncwarp();
        }
    } else {
        n_nodes = pool.n_nodes[gidx];
        status = status_out[gidx];
        if (slab_in_smem) {
            for (int i = lane; i < n_nodes; i += 32) { parent[i] = gparent[i]; visits[i] = gvisits[i]; value[i] = gvalue[i]; nm[i] = gnm[i]; }
            const int words = n_nodes * (int)(sizeof(typename G::State) / 4);
            for (int i = lane; i < words; i += 32) reinterpret_cast<uint32_t *>(st)[i] = reinterpret_cast<const uint32_t *>(gst)[i];
            __syncwarp();
            for (int i = lane; i < n_nodes; i += 32) qv[i] = __fdiv_rn(value[i], visits[i]);
            if (lane == 0) {  // rebuild the child ranges (children are created in index order)
                for (int i = 0; i < n_nodes; ++i) crange[i] = 0xFFFFu;
                for (int i = 1; i < n_nodes; ++i) {
                    const int p = parent[i];
                    const uint32_t r = crange[p];
                    crange[p] = ((r & 0xFFFFu) == 0xFFFFu ? (uint32_t)i : (r & 0xFFFFu)) | ((uint32_t)i << 16);
                }
            }
            __syncwarp();
        }
        if (stats_out) {
            const diee_search_stats ss = stats_out[gidx];
            plies = ss.rollout_plies; sel_levels = ss.select_levels; sel_children = ss.select_children; terminal_leaves = ss.terminal_leaves;
        }
    }

    if (LOCK && it_begin > 0 && n_nodes > 0 && status == DIEE_OK) {
        // ---- backpropagate :96-103 for the rollout the previous launch deferred ----
        const int leaf = sim_node[it_begin - 1];
        if (leaf >= 0) {
            if (lane == 0) {
                const float result = pool.roll_result[gidx];
                for (int i = leaf; i >= 0; i = parent[i]) {
                    visits[i] = __fadd_rn(visits[i], 1.0f);
                    value[i] = __fadd_rn(value[i], result);
                }
            }
            __syncwarp();
        }
    }
    const int first_cached = it_begin == 0 ? 1 : n_nodes;  // nodes created from here on are counted (and cached) by this launch
    const bool root_cache = !LOCK;  // a one-iteration launch pops at most one play of the root
    if (root_cache && n_nodes > 0 && status == DIEE_OK) {
        const int rootU = (int)(nm[0] >> 16), todo = min((int)(nm[0] & 0xFFFFu), ROOT_PLAYS);  // only untried plays are read
        if (todo > (it_begin == 0 ? 32 : 0)) {
            game.load(st, lane);
            for (int c0 = it_begin == 0 ? 32 : 0; c0 < todo; c0 += 32) {
                uint32_t play = SEQ_EMPTY;
                game.count_and_kth(slab, lane, ovf, c0 + lane < rootU ? c0 + lane : -1, play);
                root_plays[c0 + lane] = play;
            }
        }
        __syncwarp();
    }

    if (n_nodes > 0) {
        if (status == DIEE_OK)
        for (uint32_t it = it_begin; it < it_end; ++it) {
            int counted_node = -1;            // the node whose plays were counted during this descent ...
            uint32_t counted_last = SEQ_EMPTY;  // ... and its last play, which expand pops first
            // ---- select_leaf_node :88-94 ----
            int cur = 0;
            uint32_t nmv;
            int my_path_node = -1, depth = 0;  // lane d remembers the node of level d: the path back-propagation walks
            for (;;) {
                if (lane == depth) my_path_node = cur;
                ++depth;
                nmv = nm[cur];
                if (nmv == NM_UNKNOWN) {
                    // Node::new counts a node's legal moves eagerly (node.rs:50); nothing reads the count before
                    // the search comes back to the node, so it is taken here, on first arrival -- most nodes of a
                    // 100-iteration search are never reached again and never need it (a pool dump fills them in).
                    game.load(st + cur, lane);
                    uint32_t play = SEQ_EMPTY;  // lane j asks for the play j from the end
                    int Uc = game.count_and_kth(slab, lane, ovf, lane < NODE_PLAYS ? -3 - lane : -1, play);
                    if (nplays && lane < NODE_PLAYS) nplays[cur * NODE_PLAYS + lane] = play;
                    counted_last = __shfl_sync(FULL, play, 0);
                    counted_node = cur;
                    if (Uc == 0 && pass_child) Uc = 1;
                    nmv = ((uint32_t)Uc << 16) | (uint32_t)Uc;
                    if (lane == 0) nm[cur] = nmv;
                    __syncwarp();
                }
                const int nmoves = (int)(nmv >> 16), nunt = (int)(nmv & 0xFFFFu);
                if (nunt != 0 || nmoves == 0) break;  // has untried moves, or no children at all
                ++sel_levels;
                sel_children += (uint32_t)nmoves;
                // select_ucb :41-52 over the children of cur (creation order = index order)
                const float pvis = visits[cur];
                const float lnp = ln_table[(int)pvis];
                const float cl = __fmul_rn(cfg.c, lnp);
                float bs = -INFINITY;
                int bi = -1;
                int scan0 = 0, scan1 = n_nodes;
                if (crange) {
                    const uint32_t r = crange[cur];
                    scan0 = (int)(r & 0xFFFFu) & ~31;
                    scan1 = (int)(r >> 16) + 1;
                }
                for (int c0 = scan0; c0 < scan1; c0 += 32) {
                    const int idx = c0 + lane;
                    if (idx < scan1 && parent[idx] == cur) {
                        const float vi = visits[idx];
                        // Node::ucb node.rs:86-96: value/visits + sqrt(c*ln(parent.visits)/visits)
                        const float q = qv ? qv[idx] : __fdiv_rn(value[idx], vi);
                        const float s = __fadd_rn(q, __fsqrt_rn(__fdiv_rn(cl, vi)));
                        if (!(bs > s)) { bs = s; bi = idx; }  // within a lane idx only grows
                    }
                }
                // arg-max over the lanes, the LATER index among equal scores (max_by keeps the last maximum): two warp
                // reductions -- the best score as an order-preserving integer, then the largest index that has it
                // (scores are never NaN here: every child has visits >= 1, and never -0: values are sums from +0)
                const uint32_t sb = __float_as_uint(bs);
                const uint32_t key = bi >= 0 ? ((sb & 0x80000000u) ? ~sb : (sb | 0x80000000u)) : 0u;
                const uint32_t kmax = __reduce_max_sync(FULL, key);
                bi = __reduce_max_sync(FULL, (bi >= 0 && key == kmax) ? bi : -1);
                if (bi < 0) { status = DIEE_ERR_OVERFLOW; break; }  // cannot happen: a fully expanded node has children
                cur = bi;
            }
            if (status != DIEE_OK) break;
            const int nmoves = (int)(nmv >> 16), nunt = (int)(nmv & 0xFFFFu);
            game.load(st + cur, lane);
            const int w = game.winner();
            int leaf = cur;
            float result;
            if (w != NO_WINNER) {
                result = outcome(w, player);  // :25-30
                ++terminal_leaves;
                if (lane == 0) sim_node[it] = -1;
            } else {
                if (nunt == 0) { status = DIEE_ERR_NO_MOVES_PANIC; break; }  // node.rs:119-121 (Q6)
                // ---- Node::expand node.rs:118-137: pop the LAST untried move ----
                uint32_t seq = SEQ_EMPTY;  // stays EMPTY_MOVE for the pass child of a no-move node
                if (root_cache && cur == 0 && nunt <= ROOT_PLAYS) seq = root_plays[nunt - 1];
                else if (cur == counted_node) seq = counted_last;
                else if (nplays && cur >= first_cached && nmoves - nunt < NODE_PLAYS) seq = nplays[cur * NODE_PLAYS + (nmoves - nunt)];
                else game.count_and_kth(slab, lane, ovf, nunt - 1, seq);
                const int child = n_nodes;
                uint32_t blk[4];
                philox4x32_10((uint32_t)seed, (uint32_t)(seed >> 32), (uint32_t)child, gid, DIEE_STREAM_EXPAND, epoch, blk);
                game.step(seq, die_of(blk[0]), die_of(blk[1]), lane);
                game.store(st + child, lane);
                if (lane == 0) {
                    nm[cur] = ((uint32_t)nmoves << 16) | (uint32_t)(nunt - 1);
                    parent[child] = cur; visits[child] = 0.f; value[child] = 0.f; action[child] = seq;
                    nm[child] = NM_UNKNOWN;  // counted on first arrival (see select)
                    if (crange) {
                        const uint32_t r = crange[cur];
                        crange[cur] = ((r & 0xFFFFu) == 0xFFFFu ? (uint32_t)child : (r & 0xFFFFu)) | ((uint32_t)child << 16);
                        crange[child] = 0xFFFFu;
                    }
                }
                n_nodes = child + 1;
                leaf = child;
                // ---- Node::simulate node.rs:176-196 ----
                result = 0.f;
                const int w0 = game.winner();  // winner of the START state (Q5)
                bool deferred = false;
                if (cfg.simulate_round_limit > 0 && w0 != NO_WINNER) {
                    result = outcome(w0, player);
                } else if (SPLIT || LOCK) {
                    deferred = cfg.simulate_round_limit > 0;
                } else {
                    result = rollout_plies<G>(game, slab, lane, ovf, cfg, check_current, player, seed, gid,
                                              (epoch << 16) | (it & 0xFFFFu), plies);
                }
                if (lane == 0) sim_node[it] = deferred ? child : -1;
                if (!deferred) game.store(finals + it, lane);  // where the rollout ended
            }
            if (LOCK && sim_node_deferred(sim_node, it, lane)) { __syncwarp(); continue; }  // back-propagated by the next launch
            // ---- backpropagate :96-103 (no sign flip) ----
            // the path root .. cur was recorded during select, one node per lane; the new child (if any) joins it
            if (leaf != cur) { if (lane == depth) my_path_node = leaf; ++depth; }
            if (depth <= 32) {
                if (my_path_node >= 0) {
                    const float nv = __fadd_rn(visits[my_path_node], 1.0f), nw = __fadd_rn(value[my_path_node], result);
                    visits[my_path_node] = nv;
                    value[my_path_node] = nw;
                    if (qv) qv[my_path_node] = __fdiv_rn(nw, nv);
                }
            } else if (lane == 0) {  // a path longer than a warp: walk the parent links
                for (int i = leaf; i >= 0; i = parent[i]) {
                    const float nv = __fadd_rn(visits[i], 1.0f), nw = __fadd_rn(value[i], result);
                    visits[i] = nv;
                    value[i] = nw;
                    if (qv) qv[i] = __fdiv_rn(nw, nv);
                }
            }
            __syncwarp();
        }

        if (status == DIEE_OK && last_slice) {  // select_most_visits :71-86 (last maximum)
            float bv = -INFINITY;
            int bi = -1;
            for (int c0 = 1; c0 < n_nodes; c0 += 32) {
                const int idx = c0 + lane;
                if (idx < n_nodes && parent[idx] == 0) {
                    const float vi = visits[idx];
                    if (!(bv > vi)) { bv = vi; bi = idx; }
                }
            }
#pragma unroll
            for (int d = 16; d >= 1; d >>= 1) {
                const float os = __shfl_xor_sync(FULL, bv, d);
                const int oi = __shfl_xor_sync(FULL, bi, d);
                const bool take = oi >= 0 && (bi < 0 || (oi > bi ? !(bv > os) : (os > bv)));
                if (take) { bv = os; bi = oi; }
            }
            if (bi >= 0) best = action[bi];
        }
    }
    if (fill_counts && last_slice) {  // a pool dump shows every node as Node::new leaves it
        __syncwarp();
        for (int i = 1; i < n_nodes; ++i) {
            if (nm[i] != NM_UNKNOWN) continue;
            game.load(st + i, lane);
            uint32_t unused = SEQ_EMPTY;
            int Uc = game.count_and_kth(slab, lane, ovf, -1, unused);
            if (Uc == 0 && pass_child) Uc = 1;
            if (lane == 0) nm[i] = ((uint32_t)Uc << 16) | (uint32_t)Uc;
            __syncwarp();
        }
    }
    if (slab_in_smem) {
        __syncwarp();
        for (int i = lane; i < n_nodes; i += 32) { gparent[i] = parent[i]; gvisits[i] = visits[i]; gvalue[i] = value[i]; gnm[i] = nm[i]; }
        const int words = n_nodes * (int)(sizeof(typename G::State) / 4);
        for (int i = lane; i < words; i += 32) reinterpret_cast<uint32_t *>(gst)[i] = reinterpret_cast<const uint32_t *>(st)[i];
    }
    if (ovf && status == DIEE_OK) status = DIEE_ERR_OVERFLOW;
    if (lane == 0) {
        if (last_slice) best_out[gidx] = best;
        status_out[gidx] = status;
        pool.n_nodes[gidx] = n_nodes;
        if (stats_out) {
            diee_search_stats ss;
            ss.rollout_plies = plies; ss.select_levels = sel_levels; ss.select_children = sel_children;
            ss.expansions = n_nodes > 0 ? (uint32_t)(n_nodes - 1) : 0u; ss.terminal_leaves = terminal_leaves;
            stats_out[gidx] = ss;
        }
    }
}

// SPLIT = reference-exact rollouts only: Node::simulate tests the winner of the START state
// (node.rs:181, quirk Q5), so a rollout from a non-terminal node returns 0 whatever it plays and the
// tree never depends on it.  The tree kernel then only records which node each simulation rolls
// out from, and rollout_kernel runs all games x iterations rollouts concurrently (same stream
// coordinates, so every rollout plays the same plies as in the fused form).
//
// LOCK = lock-step search (DIEE_MODE_ROLLOUT_CHECK_CURRENT on the lane engine): the rollout DOES feed the tree, so the
// search runs as one launch per iteration for all games -- this kernel does back-propagation of the previous
// iteration's rollout result (pool.roll_result, written by lane_run_kernel<LANE_ROLLOUT_CC>), select and expand, and
// defers the new rollout to the lane kernel that follows it on the stream; a last launch with it_begin == it_end ==
// iterations back-propagates the last result and picks the move.  The pool stays in HBM / L2 between launches.
// MINB = resident CTAs per SM the register allocation must allow.  The kernel is latency-bound, so a batch of many
// waves gains from 5 CTAs per SM (102 registers, a few spilled words: 8,192 games 1.25 -> 1.13 ms, 32,768 games 4.08 ->
// 3.62 ms), while the one-wave BASELINE batch is a little faster with the full 126 (0.466 vs 0.478 ms).
template <class G, bool SPLIT, bool LOCK = false, int MINB = 1>
__global__ void __launch_bounds__(MCTS_WARPS_PER_CTA * 32, MINB)
mcts_search_kernel(const typename G::State *__restrict__ roots, int g0, int n, const int8_t *__restrict__ players,
                   diee_mcts_cfg cfg, uint32_t it_begin, uint32_t it_end, uint64_t seed, uint32_t first_game_id, uint32_t epoch,
                   Pool pool, const float *__restrict__ ln_table, uint32_t *__restrict__ best_out,
                   int32_t *__restrict__ status_out, diee_search_stats *__restrict__ stats_out, bool slab_in_smem,
                   bool fill_counts) {
    __shared__ WarpSlab slabs[MCTS_WARPS_PER_CTA];
    // the root's plays, generated once, 32 per call (every lane asks for a different one): Node::expand pops them one
    // by one over the first iterations of the search, and each pop would otherwise regenerate the root's move set
    __shared__ uint32_t root_plays_all[MCTS_WARPS_PER_CTA][ROOT_PLAYS];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int gidx = g0 + blockIdx.x * MCTS_WARPS_PER_CTA + wib;  // this launch covers games [g0, n)
    if (gidx >= n) return;
    WarpSlab &slab = slabs[wib];
    uint32_t *root_plays = root_plays_all[wib];
    extern __shared__ __align__(16) unsigned char slab_mem[];
    unsigned char *mine = slab_in_smem ? slab_mem + (size_t)wib * ((size_t)cfg.iterations + 1) * (sizeof(typename G::State) + 24 + 4 * NODE_PLAYS) : nullptr;
    mcts_game_body<G, SPLIT, LOCK>(gidx, lane, slab, root_plays, mine, roots, players, cfg, it_begin, it_end, seed, first_game_id, epoch, pool,
                                   ln_table, best_out, status_out, stats_out, fill_counts);
}

// ---------------- check-current search, persistent form ----------------
// With DIEE_MODE_ROLLOUT_CHECK_CURRENT a rollout's result feeds the tree, so iteration i+1 of a game waits for its rollout
// i -- but only for ITS rollout.  The lock-step form (one tree launch + one rollout launch per iteration) makes every game
// wait for the slowest rollout of its group, iteration after iteration.  Here every game runs at its own pace inside ONE
// launch: the games of a CTA are resident in shared memory exactly as in lane_pack_kernel (lane_kernels.cu) and queue up
// by the code their next step needs -- the five kinds of ply, or "tree step".  A warp that takes tree steps off the queue
// runs them one game after the other with the warp-per-game code of the tree kernel (mcts_game_body, LOCK form: back-
// propagate the rollout that just ended, select, expand), then each of its lanes loads the new leaf and the game goes
// back into the ply queues.  Same pool, same stream coordinates, same float operations in the same order per game as
// the lock-step form, so the results are bit-identical.  A game is bound to one slot of one CTA for the whole search.
struct CcSmem {
    PackSmem pk;
    WarpSlab slabs[PK_T / 32];
    uint32_t root_plays[PK_T / 32][32];  // (the LOCK form writes the root's first plays here and never reads them)
};

__device__ __forceinline__ void cc_load_state(LaneBoard &g, const diee_bg_state *s) {
    // written earlier in this launch by another warp of the CTA: L2, not the read-only path
    const uint4 a = __ldcg(reinterpret_cast<const uint4 *>(s));
    const uint4 b = __ldcg(reinterpret_cast<const uint4 *>(s) + 1);
    const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    l_load(g, w);
}

__global__ void __launch_bounds__(PK_T, 2)
cc_search_kernel(const diee_bg_state *__restrict__ roots, int n, const int8_t *__restrict__ players, diee_mcts_cfg cfg, uint64_t seed,
                 uint32_t first_game_id, uint32_t epoch, Pool pool, const float *__restrict__ ln_table, uint32_t *__restrict__ best_out,
                 int32_t *__restrict__ status_out, diee_search_stats *__restrict__ stats_out, bool fill_counts, int my_slots,
                 unsigned long long *__restrict__ work) {
    extern __shared__ __align__(16) unsigned char cc_smem_raw[];
    CcSmem &cs = *reinterpret_cast<CcSmem *>(cc_smem_raw);
    PackSmem &sm = cs.pk;
    const int tid = threadIdx.x, lane = tid & 31, wib = tid >> 5;
    // slot s of this CTA holds game s * gridDim.x + blockIdx.x for the whole search
    int n_mine = (n - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    n_mine = n_mine < 0 ? 0 : (n_mine > my_slots ? my_slots : n_mine);
    for (int s = tid; s < PK_S; s += PK_T) { sm.st[8][s] = 0; sm.st[10][s] = 0; sm.st[11][s] = (uint32_t)(s * (int)gridDim.x + (int)blockIdx.x); }
    for (int i = tid; i < PC_LISTS * PK_RING; i += PK_T) (&sm.ring[0][0])[i] = PK_EMPTY;
    __syncthreads();
    for (int s = tid; s < n_mine; s += PK_T) sm.ring[PC_TURN][s] = (unsigned)s;  // every game starts with a tree step
    if (tid < 8) { sm.head[tid] = 0; sm.tail[tid] = tid == PC_TURN ? (unsigned)n_mine : 0u; }
    if (tid < 4) sm.area_lock[tid] = 0;
    if (tid == 0) { sm.n_dead = PK_S - n_mine; sm.n_avail = n_mine; sm.drain = 1; }
    __syncthreads();
    volatile unsigned *vhead = sm.head, *vtail = sm.tail;
    volatile int *vlock = sm.area_lock;
    volatile int *vdead = &sm.n_dead, *vavail = &sm.n_avail;
    const size_t cap = (size_t)cfg.iterations + 1;
    for (;;) {
        unsigned ctl = 0;  // (one lane reads the CTA's counters, the warp agrees on them)
        if (lane == 0) ctl = (unsigned)*vdead | ((*vavail > 0 ? 1u : 0u) << 16);
        ctl = __shfl_sync(FULL, ctl, 0);
        if ((ctl & 0xFFFFu) >= (unsigned)PK_S) break;
        if (!((ctl >> 16) & 1u)) { __nanosleep(40u); continue; }
        // ---- the longest queue; whatever it holds is taken at once (every game of the search is resident: nothing to wait for)
        unsigned cnt = 0;
        if (lane < PC_LISTS) {
            const unsigned h0 = vhead[lane];
            cnt = vtail[lane] - h0;
            if (lane == PC_WALK && cnt && vlock[0] && vlock[1] && vlock[2]) cnt = 0;
            if (cnt > 1023u) cnt = 1023u;
        }
        const unsigned best = __reduce_max_sync(FULL, (cnt << 3) | (unsigned)(lane & 7));
        const int c = (int)(best & 7u);
        const int avail = (int)(best >> 3);
        const int take = avail >= 32 ? 32 : avail;
        if (take == 0) { __nanosleep(40u); continue; }
        int ok = 0, area = -1;
        unsigned h = 0;
        if (lane == 0) {
            if (c == PC_WALK)
                for (int a = 0; a < PK_AREAS && area < 0; ++a)
                    if (atomicCAS(&sm.area_lock[a], 0, 1) == 0) area = a;
            if (c != PC_WALK || area >= 0) {
                h = vhead[c];
                const unsigned t = vtail[c];
                if ((int)(t - h) >= take) ok = atomicCAS(&sm.head[c], h, h + (unsigned)take) == h;
                if (ok) atomicSub(&sm.n_avail, take);
                if (!ok && area >= 0) { atomicExch(&sm.area_lock[area], 0); area = -1; }
            }
        }
        ok = __shfl_sync(FULL, ok, 0);
        if (!ok) continue;
        h = __shfl_sync(FULL, h, 0);
        area = __shfl_sync(FULL, area, 0);
        const bool act = lane < take;
        int slot = 0;
        if (act) {
            slot = (int)ring_take(&sm.ring[c][(h + (unsigned)lane) & (PK_RING - 1)]);
        }
        __threadfence_block();
        __syncwarp();
        const uint32_t act_mask = __ballot_sync(FULL, act);
        int newc = -1;
        LaneBoard g;
        uint32_t k = 0, it = 0, gm = 0;
        bool keep = false;
        if (c == PC_TURN) {
            // ---- tree steps ----
            uint32_t misc = 0;
            if (act) {
                misc = sm.st[8][slot]; k = sm.st[9][slot]; it = sm.st[10][slot]; gm = sm.st[11][slot];
                if (misc & PK_HAS_GAME) {  // a rollout has ended: node.rs:181-185 on the rolled-out state (see lane_run_kernel)
#pragma unroll
                    for (int w = 0; w < 4; ++w) { g.own[w] = sm.st[w][slot]; g.opp[w] = sm.st[4 + w][slot]; }
                    unpack_misc(g, misc);
                    const int w = k < cfg.simulate_round_limit ? l_winner(g) : 0, pl = players[gm];
                    pool.roll_result[gm] = w == 0 ? 0.f : (w == pl ? 1.f : (w == -pl ? -1.f : 0.f));
                    if (stats_out) stats_out[gm].rollout_plies += k;
                    lane_store_state(g, reinterpret_cast<diee_bg_state *>(pool.finals) + (size_t)gm * cfg.iterations + it);
                    it += 1;  // the game's next iteration (== iterations: only the last back-propagation and the move are left)
                }
            }
            {
                const uint32_t sum = __reduce_add_sync(FULL, (act && (misc & PK_HAS_GAME)) ? k : 0u);
                if (lane == 0 && sum) atomicAdd(work + 1, (unsigned long long)sum);
            }
            __syncwarp();
            int my_node = -1;
            bool my_over = false;
            for (uint32_t todo = act_mask; todo; todo &= todo - 1u) {
                const int j = __ffs(todo) - 1;
                const int gj = (int)__shfl_sync(FULL, gm, j);
                uint32_t itj = __shfl_sync(FULL, it, j);
                int node = -1;
                bool over = false;
                for (;;) {  // tree steps until one defers a rollout, or the game's search is over
                    mcts_game_body<BgGame, false, true>(gj, lane, cs.slabs[wib], cs.root_plays[wib], nullptr, roots, players, cfg, itj,
                                                        itj < cfg.iterations ? itj + 1u : itj, seed, first_game_id, epoch, pool, ln_table,
                                                        best_out, status_out, stats_out, fill_counts);
                    __syncwarp();
                    if (itj >= cfg.iterations) { over = true; break; }
                    int v = 0;
                    if (lane == 0) v = __ldcg(pool.sim_node + (size_t)gj * cfg.iterations + itj);
                    v = __shfl_sync(FULL, v, 0);
                    if (v >= 0) { node = v; break; }
                    ++itj;
                }
                if (lane == j) { it = itj; my_node = node; my_over = over; }
            }
            __syncwarp();
            if (act) {
                if (my_over) {
                    newc = PC_DEAD;
                } else {
                    cc_load_state(g, reinterpret_cast<const diee_bg_state *>(pool.states) + (size_t)gm * cap + my_node);
                    k = 0;
                    keep = true;
                    newc = l_winner(g) != 0 ? PC_TURN : pack_class(g);
                    sm.st[10][slot] = it;
                }
            }
        } else if (act) {
            // ---- one ply ----
            const uint32_t misc = sm.st[8][slot];
#pragma unroll
            for (int w = 0; w < 4; ++w) { g.own[w] = sm.st[w][slot]; g.opp[w] = sm.st[4 + w][slot]; }
            unpack_misc(g, misc);
            k = sm.st[9][slot]; it = sm.st[10][slot]; gm = sm.st[11][slot];
            uint32_t o[4];
            l_philox((uint32_t)seed, (uint32_t)(seed >> 32), k, first_game_id + gm, DIEE_STREAM_ROLLOUT, (epoch << 16) | (it & 0xFFFFu), o);
            LanePlay pl;
            pl.n = 0; pl.x1 = pl.t1 = pl.x2 = pl.t2 = 0;
            if (c != PC_WALK) {
                const int hi = max(g.roll0, g.roll1), lo = min(g.roll0, g.roll1);
                LaneMasks m;
                l_closed_applies(g, m, lo, hi);
                if (g.bar_own == 0 && m.own1 != 0 && (m.own1 & ~0x3Fu) == 0) {
                    const uint32_t e = __ldg(pool.pb.index + l_pb_key(g));
                    const uint32_t U = e & 255u;
                    if (U > 0) pl = l_pb_unpack(__ldg(pool.pb.plays + (e >> 8) + l_index(o[2], U)));
                } else {
                    if (m.own1 != 0 || g.bar_own > 0) l_contact_select(g, m, lo, hi, -2, o[2], pl);
                }
            } else {
                LaneGen gen;
                uint32_t *scr = &sm.scr[area][0][lane];
                l_movegen_walk_t<true>(g, gen, scr, 32);
                if (gen.U > 0) pl = l_pick_walk(gen, scr, 32, (int)l_index(o[2], (uint32_t)gen.U));
            }
            l_step(g, pl, l_die(o[0]), l_die(o[1]));
            ++k;
            keep = true;
            newc = (k == cfg.simulate_round_limit || l_winner(g) != 0) ? PC_TURN : pack_class(g);
        }
        if (keep) {
#pragma unroll
            for (int w = 0; w < 4; ++w) { sm.st[w][slot] = g.own[w]; sm.st[4 + w][slot] = g.opp[w]; }
            sm.st[8][slot] = pack_misc(g);
            sm.st[9][slot] = k;
        }
        __threadfence_block();
        __syncwarp();
        if (area >= 0 && lane == 0) atomicExch(&sm.area_lock[area], 0);
        {
            const uint32_t same = __match_any_sync(FULL, newc);
            const int leader = __ffs(same) - 1;
            unsigned base = 0;
            if (lane == leader) {
                if (newc >= 0 && newc < PC_LISTS) { base = atomicAdd(&sm.tail[newc], (unsigned)__popc(same)); atomicAdd(&sm.n_avail, __popc(same)); }
                else if (newc == PC_DEAD) atomicAdd(&sm.n_dead, __popc(same));
            }
            base = __shfl_sync(FULL, base, leader);
            if (newc >= 0 && newc < PC_LISTS) {
                ring_put(&sm.ring[newc][(base + (unsigned)__popc(same & ((1u << lane) - 1u))) & (PK_RING - 1)], (unsigned)slot);
            }
        }
    }
}

// all deferred rollouts of a split search: one warp per (game, iteration)
template <class G>
__global__ void __launch_bounds__(MCTS_WARPS_PER_CTA * 32)
rollout_kernel(int n_games, diee_mcts_cfg cfg, uint64_t seed, uint32_t first_game_id, uint32_t epoch, Pool pool,
               const int8_t *__restrict__ players, int32_t *__restrict__ status_out, diee_search_stats *__restrict__ stats_out) {
    __shared__ WarpSlab slabs[MCTS_WARPS_PER_CTA];
    // the root's plays, generated once, 32 per call (every lane asks for a different one): Node::expand pops them one
    // by one over the first iterations of the search, and each pop would otherwise regenerate the root's move set
    __shared__ uint32_t root_plays_all[MCTS_WARPS_PER_CTA][ROOT_PLAYS];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const long long pair = (long long)blockIdx.x * MCTS_WARPS_PER_CTA + wib;
    if (pair >= (long long)n_games * cfg.iterations) return;
    const int g = (int)(pair / cfg.iterations);
    const uint32_t it = (uint32_t)(pair - (long long)g * cfg.iterations);
    const int node = pool.sim_node[pair];
    if (node < 0) return;
    const size_t cap = (size_t)cfg.iterations + 1;
    G game;
    game.attach(pool.pb);
    game.load(reinterpret_cast<const typename G::State *>(pool.states) + (size_t)g * cap + node, lane);
    bool ovf = false;
    unsigned long long plies = 0;
    rollout_plies<G>(game, slabs[wib], lane, ovf, cfg, false, players[g], seed, first_game_id + (uint32_t)g,
                     (epoch << 16) | (it & 0xFFFFu), plies);
    game.store(reinterpret_cast<typename G::State *>(pool.finals) + pair, lane);
    if (lane == 0) {
        if (stats_out) atomicAdd(reinterpret_cast<unsigned long long *>(&stats_out[g].rollout_plies), plies);
        if (ovf) atomicCAS(&status_out[g], DIEE_OK, DIEE_ERR_OVERFLOW);
    }
}

#ifdef DIEE_TRACE
// experiment build only: a timeline of the sliced search (events around every tree slice and rollout slice)
static cudaEvent_t g_trace_ev[2 + 4 * SEARCH_SLICES];
static int g_trace_slices = 0;
static void trace_init() {
    static bool done = false;
    if (!done) { for (auto &e : g_trace_ev) cudaEventCreate(&e); done = true; }
}
extern "C" int diee_debug_search_trace(float *out, int cap) {  // ms since the search began: per slice tree begin/end, rollouts begin/end
    cudaDeviceSynchronize();
    int k = 0;
    for (int s = 0; s < g_trace_slices && k + 4 <= cap; ++s)
        for (int j = 0; j < 4; ++j) cudaEventElapsedTime(&out[k++], g_trace_ev[0], g_trace_ev[2 + 4 * s + j]);
    return k;
}
#define TRACE(idx, stream) cudaEventRecord(g_trace_ev[idx], stream)
#else
#define TRACE(idx, stream) ((void)0)
#endif

static inline int mcts_grid(int n) { return (n + MCTS_WARPS_PER_CTA - 1) / MCTS_WARPS_PER_CTA; }

template <class G>
static cudaError_t launch_typed(cudaStream_t st, const void *roots, int n, const int8_t *players, const diee_mcts_cfg &cfg,
                                uint64_t seed, uint32_t first_game_id, uint32_t epoch, const Pool &pool, const PoolPtrs &pp,
                                const SearchPipe &pipe, const float *ln_table, uint32_t *best_out, int32_t *status_out,
                                diee_search_stats *stats_out, bool dump, int *launches) {
    const bool split = !(cfg.mode_flags & DIEE_MODE_ROLLOUT_CHECK_CURRENT);
    const typename G::State *r = static_cast<const typename G::State *>(roots);
    cudaError_t e;
    // node slabs of the CTA's games in shared memory when they fit
    size_t slab_bytes = (size_t)MCTS_WARPS_PER_CTA * (cfg.iterations + 1) * (sizeof(typename G::State) + 24 + 4 * NODE_PLAYS);
    const bool in_smem = slab_bytes <= 160 * 1024;
    if (!in_smem) slab_bytes = 0;
    if (in_smem) {
        if ((e = cudaFuncSetAttribute(mcts_search_kernel<G, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024)) != cudaSuccess) return e;
        if ((e = cudaFuncSetAttribute(mcts_search_kernel<G, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024)) != cudaSuccess) return e;
    }
    if constexpr (std::is_same<G, BgGame>::value) {
        // Rollouts that test the rolled-out state feed the tree, so they cannot run after it.  Lock-step instead: per
        // iteration one tree launch for all games (back-propagate the previous result, select, expand) and one lane-engine
        // launch for their rollouts (lane_run_kernel<LANE_ROLLOUT_CC>, one lane per game).  The games are cut into groups
        // on side streams so that one group's tree step runs beside another group's rollouts; what bounds a group is its
        // longest rollout, iteration after iteration, so throughput comes from the batch size.
        // DIEE_CC_FUSED=1 keeps the round-1 form (the whole search in one launch, warp-per-game rollouts) for comparison.
        // Which form: a lock-step search takes ~0.6 ms per iteration whatever the batch (the longest rollout of the
        // iteration, ~2 us per ply in a thinned-out warp), the fused form grows with it -- measured on B200, 100
        // iterations: 1,024 games 18.8 ms fused / 60 ms lock-step; 16,384 games 110 / 70.6 ms; 65,536 games: 125 ms lock-step.
        const char *fenv = getenv("DIEE_CC_FUSED");
        const bool fused = fenv ? atoi(fenv) != 0 : n < 12288;
        // DIEE_CC_PERSISTENT=1: the persistent form (cc_search_kernel: every game at its own pace inside one launch), for
        // batches whose games all fit the machine's resident slots.  EXPERIMENT, off by default: bit-identical, but measured on
        // B200 (100 iterations) 1,024 games 39.5 ms (fused 18.4), 16,384 games 88.0 ms (lock-step 71.2), 65,536 games 132.9 ms
        // (lock-step 126.1) -- a game no longer waits for the slowest rollout of its group, but its own plies cost ~3 us each
        // through the queues (one thread's dependent chain per ply, as in the lane kernel), and a tree step on the pool in
        // L2 holds a warp for tens of microseconds while the SM's issue slots are busy with plies.
        if (!split && cfg.simulate_round_limit > 0) {
            const char *penv = getenv("DIEE_CC_PERSISTENT");
            int sms = 148;
            { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); if (sms <= 0) sms = 148; }
            long long grid = ((long long)n + 31) / 32;
            if (grid > 2ll * sms) grid = 2ll * sms;
            const long long per_cta = (((long long)n + grid - 1) / grid + 31) / 32 * 32;
            const bool persistent = penv ? atoi(penv) != 0 : false;
            if (persistent && per_cta <= PK_S && cfg.iterations < 65536u) {
                if ((e = cudaFuncSetAttribute(cc_search_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(CcSmem))) != cudaSuccess) return e;
                *pipe.timed = false;
                cc_search_kernel<<<(unsigned)grid, PK_T, sizeof(CcSmem), st>>>(
                    reinterpret_cast<const diee_bg_state *>(r), n, players, cfg, seed, first_game_id, epoch, pool, ln_table, best_out,
                    status_out, stats_out, dump, (int)per_cta, pipe.queue_heads);
                *launches += 1;
                return cudaGetLastError();
            }
        }
        if (!split && !fused && cfg.simulate_round_limit > 0) {
            int groups = n >= 8192 ? 4 : (n >= 2048 ? 2 : 1);
            if (const char *ev = getenv("DIEE_CC_GROUPS")) groups = atoi(ev);
            if (groups < 1) groups = 1;
            if (groups > SEARCH_SLICES) groups = SEARCH_SLICES;
            *pipe.timed = false;
            if (groups > 1) {
                if ((e = cudaEventRecord(pipe.tree_done[0], st)) != cudaSuccess) return e;
                for (int gr = 0; gr < groups; ++gr)
                    if ((e = cudaStreamWaitEvent(pipe.side[gr], pipe.tree_done[0], 0)) != cudaSuccess) return e;
            }
            for (uint32_t it = 0; it <= cfg.iterations; ++it) {  // the last pass only back-propagates and picks the moves
                for (int gr = 0; gr < groups; ++gr) {
                    const int lo = (int)((long long)n * gr / groups), hi = (int)((long long)n * (gr + 1) / groups);
                    if (hi <= lo) continue;
                    cudaStream_t sg = groups > 1 ? pipe.side[gr] : st;
                    const uint32_t it_end = it < cfg.iterations ? it + 1 : it;
                    mcts_search_kernel<G, false, true><<<mcts_grid(hi - lo), MCTS_WARPS_PER_CTA * 32, 0, sg>>>(
                        r, lo, hi, players, cfg, it, it_end, seed, first_game_id, epoch, pool, ln_table, best_out, status_out, stats_out, false, dump);
                    *launches += 1;
                    if ((e = cudaGetLastError()) != cudaSuccess) return e;
                    if (it < cfg.iterations &&
                        (e = launch_bg_rollouts_cc(sg, lo, hi - lo, cfg, it, seed, first_game_id, epoch, pp, players, pp.roll_result, stats_out,
                                                   pipe.queue_heads + 2 * gr, launches)) != cudaSuccess)
                        return e;
                }
            }
            if (groups > 1)
                for (int gr = 0; gr < groups; ++gr) {
                    if ((e = cudaEventRecord(pipe.roll_done[gr], pipe.side[gr])) != cudaSuccess) return e;
                    if ((e = cudaStreamWaitEvent(st, pipe.roll_done[gr], 0)) != cudaSuccess) return e;
                }
            return cudaSuccess;
        }
    }
    if (!split) {
        mcts_search_kernel<G, false><<<mcts_grid(n), MCTS_WARPS_PER_CTA * 32, slab_bytes, st>>>(
            r, 0, n, players, cfg, 0u, cfg.iterations, seed, first_game_id, epoch, pool, ln_table, best_out, status_out, stats_out, in_smem, dump);
        *launches = 1;
        return cudaGetLastError();
    }
    if constexpr (std::is_same<G, BgGame>::value) {
        // backgammon: the rollouts are one lane each (lane_kernels.cu) and never feed back into the tree
        // (quirk Q5), so the search runs as slices of iterations: tree kernel of slice s on the caller's
        // stream, rollouts of slice s on a side stream beside the tree kernel of slice s + 1.
        // Measured on B200 (1,024 games x 100 iterations): 1 slice 3.39 ms, 2 slices 3.36 ms, 4 slices 3.96 ms per
        // search -- the latency-bound tree kernel loses as much to sharing the SMs as the overlap wins, so the
        // default is one slice; DIEE_SEARCH_SLICES keeps the experiment reproducible.
        // With an SM PARTITION (green contexts; SearchPipe::part_tree) the picture changes: the tree slices run on their own
        // SMs at full speed and the rollouts of slices 0 .. S-2 on the others, so a one-wave batch (the BASELINE 1,024
        // games) no longer pays tree + rollouts but about tree + the longest rollout: 4 slices by default there.
        const bool partitioned = pipe.part_tree != nullptr && n <= 2048 && cfg.iterations >= 8;
        uint32_t slices = partitioned ? 4u : 1u;
        if (const char *ev = getenv("DIEE_SEARCH_SLICES")) slices = (uint32_t)atoi(ev);
        if (slices < 1u) slices = 1u;
        if (slices > (uint32_t)SEARCH_SLICES) slices = (uint32_t)SEARCH_SLICES;
        if (slices > cfg.iterations) slices = cfg.iterations;
        *pipe.timed = false;
        if (slices == 1) {  // everything on the caller's stream, with timing marks around the two kernels
            if ((e = cudaEventRecord(pipe.t_begin, st)) != cudaSuccess) return e;
            if (n >= 4096 && in_smem) {
                if ((e = cudaFuncSetAttribute(mcts_search_kernel<G, true, false, DIEE_TREE_MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024)) != cudaSuccess) return e;
                mcts_search_kernel<G, true, false, DIEE_TREE_MINB><<<mcts_grid(n), MCTS_WARPS_PER_CTA * 32, slab_bytes, st>>>(
                    r, 0, n, players, cfg, 0u, cfg.iterations, seed, first_game_id, epoch, pool, ln_table, best_out, status_out, stats_out, in_smem, dump);
            } else
            mcts_search_kernel<G, true><<<mcts_grid(n), MCTS_WARPS_PER_CTA * 32, slab_bytes, st>>>(
                r, 0, n, players, cfg, 0u, cfg.iterations, seed, first_game_id, epoch, pool, ln_table, best_out, status_out, stats_out, in_smem, dump);
            *launches += 1;
            if ((e = cudaGetLastError()) != cudaSuccess) return e;
            if ((e = cudaEventRecord(pipe.t_tree, st)) != cudaSuccess) return e;
            if ((e = launch_bg_rollouts(st, n, cfg, 0u, cfg.iterations, seed, first_game_id, epoch, pp, pipe.queue_heads, launches)) != cudaSuccess) return e;
            if ((e = cudaEventRecord(pipe.t_end, st)) != cudaSuccess) return e;
            *pipe.timed = true;
            return launch_bg_rollout_count(st, n, cfg, pp, stats_out, launches);
        }
#ifdef DIEE_TRACE
        trace_init();
        g_trace_slices = (int)slices;
        TRACE(0, st);
#endif
        cudaStream_t ts = st;  // stream of the tree slices
        if (partitioned) {
            ts = pipe.part_tree;
            if ((e = cudaEventRecord(pipe.part_begin, st)) != cudaSuccess) return e;
            if ((e = cudaStreamWaitEvent(ts, pipe.part_begin, 0)) != cudaSuccess) return e;
        }
        for (uint32_t s = 0; s < slices; ++s) {
            const uint32_t a = (uint32_t)((uint64_t)cfg.iterations * s / slices), b = (uint32_t)((uint64_t)cfg.iterations * (s + 1) / slices);
            TRACE(2 + 4 * s, ts);
            mcts_search_kernel<G, true><<<mcts_grid(n), MCTS_WARPS_PER_CTA * 32, slab_bytes, ts>>>(
                r, 0, n, players, cfg, a, b, seed, first_game_id, epoch, pool, ln_table, best_out, status_out, stats_out, in_smem, dump);
            TRACE(3 + 4 * s, ts);
            *launches += 1;
            if ((e = cudaGetLastError()) != cudaSuccess) return e;
            if ((e = cudaEventRecord(pipe.tree_done[s], ts)) != cudaSuccess) return e;
            // rollouts of the slice: beside the next tree slice on the rollout partition; the last slice's on the whole device
            cudaStream_t rs = (partitioned && s + 1 < slices) ? pipe.part_roll[s] : pipe.side[s];
            if ((e = cudaStreamWaitEvent(rs, pipe.tree_done[s], 0)) != cudaSuccess) return e;
            TRACE(4 + 4 * s, rs);
            if ((e = launch_bg_rollouts(rs, n, cfg, a, b, seed, first_game_id, epoch, pp, pipe.queue_heads + 2 * s, launches)) != cudaSuccess) return e;
            TRACE(5 + 4 * s, rs);
            if ((e = cudaEventRecord(pipe.roll_done[s], rs)) != cudaSuccess) return e;
        }
        if (partitioned && (e = cudaStreamWaitEvent(st, pipe.tree_done[slices - 1], 0)) != cudaSuccess) return e;
        for (uint32_t s = 0; s < slices; ++s)
            if ((e = cudaStreamWaitEvent(st, pipe.roll_done[s], 0)) != cudaSuccess) return e;
        return launch_bg_rollout_count(st, n, cfg, pp, stats_out, launches);
    } else {
        mcts_search_kernel<G, true><<<mcts_grid(n), MCTS_WARPS_PER_CTA * 32, slab_bytes, st>>>(
            r, 0, n, players, cfg, 0u, cfg.iterations, seed, first_game_id, epoch, pool, ln_table, best_out, status_out, stats_out, in_smem, dump);
        const long long pairs = (long long)n * cfg.iterations;
        const long long blocks = (pairs + MCTS_WARPS_PER_CTA - 1) / MCTS_WARPS_PER_CTA;
        rollout_kernel<G><<<(unsigned)blocks, MCTS_WARPS_PER_CTA * 32, 0, st>>>(n, cfg, seed, first_game_id, epoch, pool, players,
                                                                             status_out, stats_out);
        *launches = 2;
        return cudaGetLastError();
    }
}

cudaError_t launch_mcts_search(cudaStream_t st, int game_kind, const void *roots, int n, const int8_t *players,
                               const diee_mcts_cfg &cfg, uint64_t seed, uint32_t first_game_id, uint32_t epoch,
                               const PoolPtrs &pp, const SearchPipe &pipe, const float *ln_table, uint32_t *best_out,
                               int32_t *status_out, diee_search_stats *stats_out, bool dump, int *launches) {
    *launches = 0;
    if (n <= 0) return cudaSuccess;
    Pool pool{pp.states, pp.parent, pp.visits, pp.value, pp.action, pp.nmoves, pp.n_nodes, pp.sim_node, pp.finals, pp.pb, pp.roll_result};
    if (game_kind == DIEE_GAME_BACKGAMMON)
        return launch_typed<BgGame>(st, roots, n, players, cfg, seed, first_game_id, epoch, pool, pp, pipe, ln_table, best_out,
                                    status_out, stats_out, dump, launches);
    return launch_typed<TttGame>(st, roots, n, players, cfg, seed, first_game_id, epoch, pool, pp, pipe, ln_table, best_out,
                                 status_out, stats_out, dump, launches);
}

}  // namespace diee
