// arena.cu -- the device-resident arena: `versus::play` (src/versus.rs:160-268) and `get_actions_for_player`
// (:270-318) for Agent::Mcts / Agent::Random with the games living in HBM from the first round to the last.
//
// The reference keeps 400 games in a HashMap and, every round, partitions them by the side to move, asks each side's
// agent for one action per game, applies the actions, tests winner / round limit and retires finished games.  Here a
// round is a fixed handful of launches on the context's stream and ONE 20-byte read-back:
//   arena_prepare_kernel   side s's batch = the live games whose mover is s (everybody else: a finished dummy whose
//                          search is no work), dense and indexed by game so that every stream stays keyed by the game
//   per side               Agent::Mcts:   diee_mcts_search_dev (tree kernel + lane-engine rollouts, epoch = round)
//                          Agent::Random: bg_valid_moves_kernel + arena_pick_kernel (word 2 of the game's GAME block)
//   arena_apply_kernel     apply_move / skip_turn with the round's dice (words 0, 1 of the same block), the winner and
//                          round-limit tests (not after a skipped turn, versus.rs:222-225), retirement, win accounting
// Draw sites and order are those of die_e_b200/versus.py (its header), which is checked game by game against the oracle
// twin of the reference's loop (tests/orc_arena.py); tests/test_gpu_arena.py holds this form to the same results.
#include <cuda_runtime.h>

#include <vector>

#include "bg_device.cuh"
#include "ctx.h"
#include "launchers.h"

using namespace diee;

struct diee_arena {
    int n = 0, round = 0, round_limit = 0;
    uint64_t seed = 0;
    diee_bg_state *states = nullptr, *dense[2] = {nullptr, nullptr};
    int8_t *players[2] = {nullptr, nullptr}, *winners = nullptr;
    uint8_t *live = nullptr;
    uint32_t *best[2] = {nullptr, nullptr};
    int32_t *status[2] = {nullptr, nullptr}, *rounds = nullptr, *counts = nullptr, *summary = nullptr;
    diee_move *moves = nullptr;
};

namespace {

constexpr int ARENA_WARPS = 8;

__global__ void arena_prepare_kernel(const diee_bg_state *__restrict__ states, const uint8_t *__restrict__ live, int n,
                                     diee_bg_state *__restrict__ dense1, diee_bg_state *__restrict__ dense2,
                                     int8_t *__restrict__ players1, int8_t *__restrict__ players2) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n) return;
    diee_bg_state dummy{};  // a finished game: its root is terminal, so its search is no work and returns EMPTY_MOVE
    dummy.off[0] = 15; dummy.roll[0] = 1; dummy.roll[1] = 2; dummy.player = -1;
    const diee_bg_state s = states[g];
    const bool l = live[g] != 0;
    dense1[g] = (l && s.player == -1) ? s : dummy;
    dense2[g] = (l && s.player != -1) ? s : dummy;
    players1[g] = dense1[g].player;
    players2[g] = dense2[g].player;
}

// Agent::Random (versus.rs:307-316): valid_moves.choose() = word 2 of the game's GAME block of this round
__global__ void arena_pick_kernel(const diee_move *__restrict__ moves, const int32_t *__restrict__ counts, int n, uint64_t seed,
                                  uint32_t round, uint32_t *__restrict__ best, int32_t *__restrict__ status) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n) return;
    const int c = counts[g];
    uint32_t seq = SEQ_EMPTY;
    if (c > 0) {
        uint32_t o[4];
        philox4x32_10((uint32_t)seed, (uint32_t)(seed >> 32), round, (uint32_t)g, DIEE_STREAM_GAME, 0, o);
        seq = reinterpret_cast<const uint32_t *>(moves)[(size_t)g * DIEE_MAX_MOVES + index_of(o[2], (uint32_t)c)];
    }
    best[g] = seq;
    status[g] = c < 0 ? c : DIEE_OK;
}

// summary: [0] games retired so far, [1] wins of player 1 (the -1 side), [2] wins of player 2, [3] draws, [4] first error
__global__ void __launch_bounds__(ARENA_WARPS * 32)
arena_apply_kernel(diee_bg_state *__restrict__ states, uint8_t *__restrict__ live, int n, const uint32_t *__restrict__ best1,
                   const uint32_t *__restrict__ best2, const int32_t *__restrict__ status1, const int32_t *__restrict__ status2,
                   uint64_t seed, uint32_t round, int round_limit, int8_t *__restrict__ winners, int32_t *__restrict__ rounds,
                   int32_t *__restrict__ summary) {
    const int lane = threadIdx.x & 31, g = blockIdx.x * ARENA_WARPS + (threadIdx.x >> 5);
    if (g >= n || !live[g]) return;
    BgWarp s;
    bg_load(s, states + g, lane);
    const bool p1 = s.player == -1;
    const int st = p1 ? status1[g] : status2[g];
    if (st != DIEE_OK) {  // e.g. the reference's panic on a no-move node (node.rs:119-121) without DIEE_MODE_PASS_CHILD
        if (lane == 0) atomicCAS(&summary[4], 0, st);
        return;
    }
    const uint32_t seq = p1 ? best1[g] : best2[g];
    uint32_t o[4];
    philox4x32_10((uint32_t)seed, (uint32_t)(seed >> 32), round, (uint32_t)g, DIEE_STREAM_GAME, 0, o);
    bg_step(s, seq, die_of(o[0]), die_of(o[1]), lane);  // EMPTY_MOVE -> skip_turn
    bg_store(s, states + g, lane);
    if (seq == SEQ_EMPTY || lane != 0) return;           // versus.rs:222-225: no winner / round-limit test after a skipped turn
    const int round_count = (int)round + 1;
    int w = bg_winner(s);
    const bool over = w != 0 || round_count >= round_limit;  // :231-248
    if (!over) return;
    live[g] = 0;
    winners[g] = (int8_t)w;
    rounds[g] = round_count;
    atomicAdd(&summary[0], 1);
    atomicAdd(&summary[w == -1 ? 1 : (w == 1 ? 2 : 3)], 1);
}

}  // namespace

extern "C" {

int32_t diee_arena_destroy(diee_ctx *ctx, diee_arena *a) {
    if (!ctx || !a) return DIEE_ERR_INVALID;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    void *ptrs[] = {a->states, a->dense[0], a->dense[1], a->players[0], a->players[1], a->winners, a->live, a->best[0], a->best[1],
                    a->status[0], a->status[1], a->rounds, a->counts, a->summary, a->moves};
    for (void *p : ptrs)
        if (p) cudaFree(p);
    delete a;
    return DIEE_OK;
}

int32_t diee_arena_create(diee_ctx *ctx, int32_t n_games, uint64_t seed, int32_t round_limit, diee_arena **out) {
    if (!ctx || !out || n_games <= 0 || round_limit < 0) return fail(ctx, DIEE_ERR_INVALID, "arena_create: bad argument");
    *out = nullptr;
    CU(cudaSetDevice(ctx->device));
    diee_arena *a = new diee_arena();
    a->n = n_games; a->seed = seed; a->round_limit = round_limit;
    const size_t n = (size_t)n_games;
    bool ok = cudaMalloc(&a->states, 32 * n) == cudaSuccess && cudaMalloc(&a->dense[0], 32 * n) == cudaSuccess &&
              cudaMalloc(&a->dense[1], 32 * n) == cudaSuccess && cudaMalloc(&a->players[0], n) == cudaSuccess &&
              cudaMalloc(&a->players[1], n) == cudaSuccess && cudaMalloc(&a->winners, n) == cudaSuccess &&
              cudaMalloc(&a->live, n) == cudaSuccess && cudaMalloc(&a->best[0], 4 * n) == cudaSuccess &&
              cudaMalloc(&a->best[1], 4 * n) == cudaSuccess && cudaMalloc(&a->status[0], 4 * n) == cudaSuccess &&
              cudaMalloc(&a->status[1], 4 * n) == cudaSuccess && cudaMalloc(&a->rounds, 4 * n) == cudaSuccess &&
              cudaMalloc(&a->counts, 4 * n) == cudaSuccess && cudaMalloc(&a->summary, 32) == cudaSuccess &&
              cudaMalloc(&a->moves, sizeof(diee_move) * DIEE_MAX_MOVES * n) == cudaSuccess;
    if (!ok) { diee_arena_destroy(ctx, a); return fail(ctx, DIEE_ERR_CUDA, "arena_create: out of device memory"); }
    // versus.rs:170-181: T::new(); games of the second half skip_turn first (the other side begins) and then roll
    std::vector<diee_bg_state> init(n);
    static const int8_t opening[24] = {2, 0, 0, 0, 0, -5, 0, -3, 0, 0, 0, 5, -5, 0, 0, 0, 3, 0, 5, 0, 0, 0, 0, -2};
    for (int g = 0; g < n_games; ++g) {
        diee_bg_state &s = init[g];
        memset(&s, 0, sizeof s);
        memcpy(s.pts, opening, 24);
        uint32_t o[4];
        philox4x32_10((uint32_t)seed, (uint32_t)(seed >> 32), 0, (uint32_t)g, DIEE_STREAM_INIT, 0, o);
        const bool second_half = g >= n_games / 2;
        s.player = second_half ? 1 : -1;
        s.roll[0] = (uint8_t)die_of(second_half ? o[2] : o[0]);
        s.roll[1] = (uint8_t)die_of(second_half ? o[3] : o[1]);
    }
    cudaStream_t st = ctx->stream;
    ok = cudaMemcpyAsync(a->states, init.data(), 32 * n, cudaMemcpyHostToDevice, st) == cudaSuccess &&
         cudaMemsetAsync(a->live, 1, n, st) == cudaSuccess && cudaMemsetAsync(a->winners, 0, n, st) == cudaSuccess &&
         cudaMemsetAsync(a->rounds, 0, 4 * n, st) == cudaSuccess && cudaMemsetAsync(a->summary, 0, 32, st) == cudaSuccess &&
         cudaMemsetAsync(a->status[0], 0, 4 * n, st) == cudaSuccess && cudaMemsetAsync(a->status[1], 0, 4 * n, st) == cudaSuccess &&
         cudaStreamSynchronize(st) == cudaSuccess;
    if (!ok) { diee_arena_destroy(ctx, a); return fail(ctx, DIEE_ERR_CUDA, "arena_create: upload failed"); }
    *out = a;
    return DIEE_OK;
}

int32_t diee_arena_round(diee_ctx *ctx, diee_arena *a, int32_t agent_p1, int32_t agent_p2, const diee_mcts_cfg *cfg, int32_t *summary_out) {
    if (!ctx || !a || !summary_out) return fail(ctx, DIEE_ERR_INVALID, "arena_round: bad argument");
    const int32_t agents[2] = {agent_p1, agent_p2};
    for (int s = 0; s < 2; ++s) {
        if (agents[s] != DIEE_AGENT_RANDOM && agents[s] != DIEE_AGENT_MCTS) return fail(ctx, DIEE_ERR_INVALID, "arena_round: unknown agent %d", agents[s]);
        if (agents[s] == DIEE_AGENT_MCTS && !cfg) return fail(ctx, DIEE_ERR_INVALID, "arena_round: Agent::Mcts needs a config");
    }
    CU(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const int n = a->n;
    arena_prepare_kernel<<<(n + 255) / 256, 256, 0, st>>>(a->states, a->live, n, a->dense[0], a->dense[1], a->players[0], a->players[1]);
    CU(cudaGetLastError());
    ctx->launches += 1;
    for (int s = 0; s < 2; ++s) {
        if (agents[s] == DIEE_AGENT_MCTS) {
            int32_t rc = diee_mcts_search_dev(ctx, DIEE_GAME_BACKGAMMON, a->dense[s], n, a->players[s], cfg, a->seed, 0, (uint32_t)a->round,
                                              a->best[s], a->status[s], nullptr);
            if (rc != DIEE_OK) return rc;
        } else {
            CU(launch_bg_valid_moves(st, a->dense[s], n, a->moves, a->counts, nullptr));
            arena_pick_kernel<<<(n + 255) / 256, 256, 0, st>>>(a->moves, a->counts, n, a->seed, (uint32_t)a->round, a->best[s], a->status[s]);
            CU(cudaGetLastError());
            ctx->launches += 2;
        }
    }
    arena_apply_kernel<<<(n + ARENA_WARPS - 1) / ARENA_WARPS, ARENA_WARPS * 32, 0, st>>>(
        a->states, a->live, n, a->best[0], a->best[1], a->status[0], a->status[1], a->seed, (uint32_t)a->round, a->round_limit, a->winners,
        a->rounds, a->summary);
    CU(cudaGetLastError());
    ctx->launches += 1;
    a->round += 1;
    CU(cudaMemcpyAsync(summary_out, a->summary, 5 * sizeof(int32_t), cudaMemcpyDeviceToHost, st));  // the round's only copy
    CU(cudaStreamSynchronize(st));
    if (summary_out[4] != DIEE_OK)
        return fail(ctx, summary_out[4], summary_out[4] == DIEE_ERR_NO_MOVES_PANIC
                    ? "arena_round: expand() called on node with no expandable moves (node.rs:119-121; use DIEE_MODE_PASS_CHILD for arena play)"
                    : "arena_round: a search failed");
    return DIEE_OK;
}

int32_t diee_arena_read(diee_ctx *ctx, diee_arena *a, diee_bg_state *states_out, int8_t *winners_out, int32_t *rounds_out) {
    if (!ctx || !a) return fail(ctx, DIEE_ERR_INVALID, "arena_read: bad argument");
    CU(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const size_t n = (size_t)a->n;
    if (states_out) CU(cudaMemcpyAsync(states_out, a->states, 32 * n, cudaMemcpyDeviceToHost, st));
    if (winners_out) CU(cudaMemcpyAsync(winners_out, a->winners, n, cudaMemcpyDeviceToHost, st));
    if (rounds_out) CU(cudaMemcpyAsync(rounds_out, a->rounds, 4 * n, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return DIEE_OK;
}

}  // extern "C"
