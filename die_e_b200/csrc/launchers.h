// launchers.h -- internal: kernel launchers shared between the .cu files and api.cu
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/diee.h"

namespace diee {

// the pure bear-off play table (bg_pb_table.h), device copies owned by the context
struct PbTable {
    const uint32_t *index;
    const uint16_t *plays;
};

struct PoolPtrs {
    void *states;
    int32_t *parent;
    float *visits, *value;
    uint32_t *action;
    uint32_t *nmoves;
    int32_t *n_nodes;
    int32_t *sim_node;
    void *finals;
    PbTable pb;
    float *roll_result;  // [game]: result of the game's pending rollout (lock-step search)
};

// side streams of a split search: rollouts of one slice of iterations run beside the tree kernel of the next
constexpr int SEARCH_SLICES = 4;
struct SearchPipe {
    cudaStream_t side[SEARCH_SLICES];
    cudaEvent_t tree_done[SEARCH_SLICES], roll_done[SEARCH_SLICES];
    unsigned long long *queue_heads;  // SEARCH_SLICES job queue heads of the persistent lane kernel
    cudaEvent_t t_begin, t_tree, t_end;  // timing marks of an unsliced search (tree kernel | rollouts)
    bool *timed;                         // set when the marks of the last search are valid
    // SM partition (green contexts, api.cu): the tree slices run on their own SMs, the rollouts of all but the last slice on
    // the rest, so the latency-bound tree kernel never shares an SM with rollout warps.  Null streams = no partition.
    cudaStream_t part_tree;
    cudaStream_t part_roll[SEARCH_SLICES];
    cudaEvent_t part_begin;
};

cudaError_t launch_bg_valid_moves(cudaStream_t st, const diee_bg_state *states, int n, diee_move *moves_out,
                                  int32_t *counts_out, uint16_t *ids_out);
cudaError_t launch_bg_apply(cudaStream_t st, diee_bg_state *states, const diee_move *moves, const uint8_t *next_rolls, int n);
cudaError_t launch_bg_playout(cudaStream_t st, const diee_bg_state *starts, int n, uint64_t seed, uint32_t first_game_id,
                              int round_limit, int8_t *winners_out, int32_t *plies_out, diee_bg_state *finals_out,
                              unsigned long long *queue_head, const PbTable &pb, int *launches);
// every deferred rollout of a split backgammon search, one lane per (game, iteration)  (lane_kernels.cu)
cudaError_t launch_bg_rollouts(cudaStream_t st, int n_games, const diee_mcts_cfg &cfg, uint32_t it_begin, uint32_t it_end,
                               uint64_t seed, uint32_t first_game_id, uint32_t epoch, const PoolPtrs &pp,
                               unsigned long long *queue_head, int *launches);
// lock-step search (rollouts that test the rolled-out state): the rollouts of ONE iteration, one lane per game
cudaError_t launch_bg_rollouts_cc(cudaStream_t st, int g0, int n_games, const diee_mcts_cfg &cfg, uint32_t it, uint64_t seed,
                                  uint32_t first_game_id, uint32_t epoch, const PoolPtrs &pp, const int8_t *players, float *results,
                                  diee_search_stats *stats, unsigned long long *queue_head, int *launches);
cudaError_t launch_bg_rollout_count(cudaStream_t st, int n_games, const diee_mcts_cfg &cfg, const PoolPtrs &pp,
                                    diee_search_stats *stats_out, int *launches);
cudaError_t launch_bg_encode_moves(cudaStream_t st, const diee_bg_state *states, const diee_move *moves, int n, uint16_t *ids_out);
cudaError_t launch_bg_decode_moves(cudaStream_t st, const diee_bg_state *states, const uint16_t *ids, int n, diee_move *moves_out);
cudaError_t launch_bg_encode_states(cudaStream_t st, const diee_bg_state *states, int n, float *out);
cudaError_t launch_mcts_search(cudaStream_t st, int game_kind, const void *roots, int n, const int8_t *players,
                               const diee_mcts_cfg &cfg, uint64_t seed, uint32_t first_game_id, uint32_t epoch,
                               const PoolPtrs &pp, const SearchPipe &pipe, const float *ln_table, uint32_t *best_out,
                               int32_t *status_out, diee_search_stats *stats_out, bool dump, int *launches);

}  // namespace diee
