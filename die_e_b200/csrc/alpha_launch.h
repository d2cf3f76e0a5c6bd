// alpha_launch.h -- internal: the AlphaZero node pool and kernel launchers
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/diee.h"

namespace diee {

struct AlphaPool {
    int max_nodes;        // slab size per game
    void *state;          // diee_bg_state [n * max_nodes]
    int32_t *parent, *first, *nchild;
    float *visits, *value, *prior;
    uint32_t *action;
    int32_t *n_nodes;     // [n]
    int32_t *sel_game, *sel_node, *status;  // [n]
    int32_t *any_selected;                  // [iterations]
    diee_bg_state *batch;                   // [n]   states evaluated this iteration
    float *policy;                          // [n * 1352]
    float *value_out;                       // [n]
    const float *dirichlet;                 // [1352]
};

cudaError_t launch_alpha_root(cudaStream_t st, const AlphaPool &P, const diee_bg_state *states, const uint32_t *game_ids, int n,
                              const diee_mcts_cfg &cfg, uint64_t seed, uint32_t epoch);
cudaError_t launch_alpha_select(cudaStream_t st, const AlphaPool &P, int n, const diee_mcts_cfg &cfg, int iter);
cudaError_t launch_alpha_expand(cudaStream_t st, const AlphaPool &P, const uint32_t *game_ids, int n, const diee_mcts_cfg &cfg,
                                uint64_t seed, uint32_t epoch, int iter);
// non-parity: up to K leaves per game and step with virtual loss vl; `budget` = leaves still allowed this step
cudaError_t launch_alpha_select_vl(cudaStream_t st, const AlphaPool &P, int n, const diee_mcts_cfg &cfg, int K, float vl, int budget);
cudaError_t launch_alpha_expand_vl(cudaStream_t st, const AlphaPool &P, const uint32_t *game_ids, int n, const diee_mcts_cfg &cfg,
                                   uint64_t seed, uint32_t epoch, int K, float vl);
cudaError_t launch_alpha_root_out(cudaStream_t st, const AlphaPool &P, int n, uint16_t *ids_out, uint32_t *moves_out,
                                  float *visits_out, int32_t *counts_out);

}  // namespace diee
