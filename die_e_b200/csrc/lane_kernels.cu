// lane_kernels.cu -- one LANE per game: the throughput kernels of the path.
//
//   bg_playout_lane_kernel   C2: whole random-vs-random games
//   bg_rollout_lane_kernel   T4: every deferred Node::simulate of a split search (node.rs:176-196),
//                            one lane per (game, iteration)
//
// A ply is get_valid_moves -> uniform choice -> apply_move | skip_turn.  The board lives in the lane's
// registers as bit planes (bg_lane.cuh); the only memory a ply touches is the lane's own column of a
// shared-memory scratch (new-children masks per root), laid out [word][lane] so that a warp's accesses
// never conflict.  HBM: 32 B in and 32 B (+5 B) out per game / per rollout, nothing in between.
// Randomness: one Philox4x32-10 block per ply, counter = (ply, game id, stream, epoch<<16|iteration)
// -- the contract of include/diee.h, identical to the warp-per-game kernels and to the oracle.
#include "bg_lane.cuh"
#include "launchers.h"

namespace diee {

using namespace lane;

constexpr int LANE_CTA = 128;

struct LaneScratch {
    uint32_t w[L_SCRATCH][LANE_CTA];
};

__device__ __forceinline__ void lane_load_state(LaneBoard &g, const diee_bg_state *s) {
    const uint4 a = __ldg(reinterpret_cast<const uint4 *>(s));
    const uint4 b = __ldg(reinterpret_cast<const uint4 *>(s) + 1);
    const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    l_load(g, w);
}
__device__ __forceinline__ void lane_store_state(const LaneBoard &g, diee_bg_state *s) {
    uint32_t w[8];
    l_store(g, w);
    reinterpret_cast<uint4 *>(s)[0] = make_uint4(w[0], w[1], w[2], w[3]);
    reinterpret_cast<uint4 *>(s)[1] = make_uint4(w[4], w[5], w[6], w[7]);
}

// one ply with the Philox block of (ply k, game gid, stream, c3)
__device__ __forceinline__ void lane_ply(LaneBoard &g, uint32_t *scr, uint64_t seed, uint32_t k, uint32_t gid,
                                         uint32_t stream, uint32_t c3) {
    uint32_t o[4];
    l_philox((uint32_t)seed, (uint32_t)(seed >> 32), k, gid, stream, c3, o);
    LaneGen gen;
    l_movegen(g, gen, scr, LANE_CTA);
    LanePlay pl;
    pl.n = 0; pl.x1 = pl.t1 = pl.x2 = pl.t2 = 0;
    if (gen.U > 0) pl = l_pick(gen, scr, LANE_CTA, (int)l_index(o[2], (uint32_t)gen.U));
    l_step(g, pl, l_die(o[0]), l_die(o[1]));
}

__global__ void __launch_bounds__(LANE_CTA)
bg_playout_lane_kernel(const diee_bg_state *__restrict__ starts, int n, uint64_t seed, uint32_t first_game_id,
                       int round_limit, int8_t *__restrict__ winners_out, int32_t *__restrict__ plies_out,
                       diee_bg_state *__restrict__ finals_out) {
    __shared__ LaneScratch scratch;
    const int gidx = blockIdx.x * LANE_CTA + threadIdx.x;
    if (gidx >= n) return;
    uint32_t *scr = &scratch.w[0][threadIdx.x];
    LaneBoard g;
    lane_load_state(g, starts + gidx);
    const uint32_t gid = first_game_id + (uint32_t)gidx;
    int ply = 0;
    int w = l_winner(g);
    while (w == 0 && ply < round_limit) {
        lane_ply(g, scr, seed, (uint32_t)ply, gid, DIEE_STREAM_GAME, 0u);
        ++ply;
        w = l_winner(g);
    }
    winners_out[gidx] = (int8_t)w;
    plies_out[gidx] = ply;
    if (finals_out) lane_store_state(g, finals_out + gidx);
}

// Every deferred rollout of a split (reference-exact) search.  Node::simulate tests the winner of its
// START state (node.rs:181, quirk Q5) and the tree kernel only defers non-terminal starts, so each
// rollout plays exactly `limit` plies; its result is 0 and only the plies and the final state are kept.
__global__ void __launch_bounds__(LANE_CTA)
bg_rollout_lane_kernel(int n_games, uint32_t iterations, uint32_t limit, uint64_t seed, uint32_t first_game_id,
                       uint32_t epoch, const diee_bg_state *__restrict__ node_states, const int32_t *__restrict__ sim_node,
                       diee_bg_state *__restrict__ finals, diee_search_stats *__restrict__ stats_out) {
    __shared__ LaneScratch scratch;
    const long long pair = (long long)blockIdx.x * LANE_CTA + threadIdx.x;
    if (pair >= (long long)n_games * iterations) return;
    const int gm = (int)(pair / iterations);
    const uint32_t it = (uint32_t)(pair - (long long)gm * iterations);
    const int node = sim_node[pair];
    if (node < 0) return;
    uint32_t *scr = &scratch.w[0][threadIdx.x];
    LaneBoard g;
    lane_load_state(g, node_states + (size_t)gm * (iterations + 1) + node);
    const uint32_t gid = first_game_id + (uint32_t)gm;
    const uint32_t c3 = (epoch << 16) | (it & 0xFFFFu);
    for (uint32_t k = 0; k < limit; ++k) lane_ply(g, scr, seed, k, gid, DIEE_STREAM_ROLLOUT, c3);
    lane_store_state(g, finals + pair);
    if (stats_out) atomicAdd(reinterpret_cast<unsigned long long *>(&stats_out[gm].rollout_plies), (unsigned long long)limit);
}

cudaError_t launch_bg_playout(cudaStream_t st, const diee_bg_state *starts, int n, uint64_t seed, uint32_t first_game_id,
                              int round_limit, int8_t *winners_out, int32_t *plies_out, diee_bg_state *finals_out) {
    if (n <= 0) return cudaSuccess;
    bg_playout_lane_kernel<<<(n + LANE_CTA - 1) / LANE_CTA, LANE_CTA, 0, st>>>(starts, n, seed, first_game_id, round_limit,
                                                                               winners_out, plies_out, finals_out);
    return cudaGetLastError();
}

cudaError_t launch_bg_rollouts(cudaStream_t st, int n_games, const diee_mcts_cfg &cfg, uint64_t seed, uint32_t first_game_id,
                               uint32_t epoch, const PoolPtrs &pp, diee_search_stats *stats_out) {
    const long long pairs = (long long)n_games * cfg.iterations;
    if (pairs <= 0 || cfg.simulate_round_limit == 0) return cudaSuccess;
    const long long blocks = (pairs + LANE_CTA - 1) / LANE_CTA;
    bg_rollout_lane_kernel<<<(unsigned)blocks, LANE_CTA, 0, st>>>(
        n_games, cfg.iterations, cfg.simulate_round_limit, seed, first_game_id, epoch,
        static_cast<const diee_bg_state *>(pp.states), pp.sim_node, static_cast<diee_bg_state *>(pp.finals), stats_out);
    return cudaGetLastError();
}

}  // namespace diee
