// alpha_kernels.cu -- AlphaZero tree kernels: lock-step batched
// search over N games, one warp per game, the net evaluated once per iteration for the whole batch.
//
// Reference: src/mcts/alpha_mcts.rs (alpha_select_leaf_node :14-20, select_alpha :22-33,
// alpha_mcts_parallel :91-202), src/mcts/node.rs (alpha_ucb :98-112, alpha_expand_tensor :157-174),
// src/mcts/utils.rs (turn_policy_to_probs_tensor(_parallel) :60-84), src/mcts/noise.rs :27-34.
//
// Node pool: structure of arrays in HBM, one slab of `max_nodes` nodes per game; a node's children
// are allocated together, so they are the contiguous run [first_child, first_child + n_children)
// in legal-move order and PUCT select reads coalesced runs of visits/value/prior.
// Arithmetic is IEEE f32 with explicit round-to-nearest intrinsics in the reference's evaluation
// order; arg-max keeps the LAST maximum (Rust max_by); masked-policy sums are sequential in
// legal-move order (the contract fixed in include/diee.h).  The reference's quirks are kept:
// Q9 (a game whose leaf was terminal re-evaluates and re-backpropagates its previous selection,
// initially arena node 0 = game 0's root), Q11 (one Dirichlet vector for the batch, mixed before
// masking), Q12 (a no-move leaf is re-selected forever), Q13 (dice frozen at expansion).
#include "alpha_launch.h"
#include "bg_device.cuh"

namespace diee {

constexpr int AW = 4;  // warps per CTA

__device__ __forceinline__ void load_state(BgWarp &g, const AlphaPool &P, size_t node, int lane) {
    bg_load(g, reinterpret_cast<const diee_bg_state *>(P.state) + node, lane);
}

// alpha_expand_tensor (node.rs:157-174) with priors = row[encode(move)] / sum over the legal moves
// (utils.rs:60-84).  mix != nullptr: row value = (1-eps)*P + eps*dirichlet (noise.rs:27-34).
__device__ __forceinline__ void alpha_expand(const AlphaPool &P, int game, int node, const float *__restrict__ row,
                                             const float *__restrict__ mix, float eps, uint64_t seed, uint32_t gid,
                                             uint32_t epoch, WarpSlab &slab, int lane, int &status) {
    const size_t base = (size_t)game * P.max_nodes;
    if (P.nchild[base + node] != 0) return;  // expandable_moves already drained
    BgWarp parent;
    load_state(parent, P, base + node, lane);
    bool ovf = false;
    const int U = bg_movegen(parent, slab, lane, ovf);
    if (ovf) { status = DIEE_ERR_OVERFLOW; return; }
    if (U == 0) return;  // Q12
    const int first = P.n_nodes[game];
    if (first + U > P.max_nodes) { status = DIEE_ERR_OVERFLOW; return; }
    // selected policy value of every legal move -> slab.kept (as float bits)
    const float a = __fsub_rn(1.0f, eps);
    for (int k = lane; k < U; k += 32) {
        const uint32_t id = bg_encode_move(parent.roll0, parent.roll1, slab.raw[k]);
        float v = row[id];
        if (mix) v = __fadd_rn(__fmul_rn(a, v), __fmul_rn(eps, mix[id]));
        slab.kept[k] = __float_as_uint(v);
    }
    __syncwarp();
    float sum = 0.f;
    if (lane == 0)
        for (int k = 0; k < U; ++k) sum = __fadd_rn(sum, __uint_as_float(slab.kept[k]));  // sequential, legal-move order
    sum = __shfl_sync(FULL, sum, 0);
    for (int k0 = 0; k0 < U; k0 += 32) {
        const int k = k0 + lane;
        uint32_t blk[4] = {0, 0, 0, 0};
        if (k < U) {
            const size_t ch = base + first + k;
            P.parent[ch] = node;
            P.first[ch] = -1;
            P.nchild[ch] = 0;
            P.visits[ch] = 0.f;
            P.value[ch] = 0.f;
            P.prior[ch] = __fdiv_rn(__uint_as_float(slab.kept[k]), sum);
            P.action[ch] = slab.raw[k];
            philox4x32_10((uint32_t)seed, (uint32_t)(seed >> 32), (uint32_t)(first + k), gid, DIEE_STREAM_EXPAND, epoch, blk);
        }
        const int cnt = min(32, U - k0);
        for (int j = 0; j < cnt; ++j) {  // child states: the warp applies one move at a time
            BgWarp c = parent;
            const int d0 = die_of(__shfl_sync(FULL, blk[0], j));
            const int d1 = die_of(__shfl_sync(FULL, blk[1], j));
            bg_step(c, slab.raw[k0 + j], d0, d1, lane);
            bg_store(c, reinterpret_cast<diee_bg_state *>(P.state) + base + first + k0 + j, lane);
        }
    }
    if (lane == 0) {
        P.first[base + node] = first;
        P.nchild[base + node] = U;
        P.n_nodes[game] = first + U;
    }
    __syncwarp();
}

__device__ __forceinline__ void alpha_backprop(const AlphaPool &P, int game, int node, float v, int lane) {
    if (lane == 0) {
        const size_t base = (size_t)game * P.max_nodes;
        for (int i = node; i >= 0; i = P.parent[base + i]) {
            P.visits[base + i] = __fadd_rn(P.visits[base + i], 1.0f);
            P.value[base + i] = __fadd_rn(P.value[base + i], v);
        }
    }
    __syncwarp();
}

// root phase (alpha_mcts.rs:97-127): policy rows of the N roots are already in P.policy
__global__ void __launch_bounds__(AW * 32)
alpha_root_kernel(AlphaPool P, const diee_bg_state *__restrict__ states, const uint32_t *__restrict__ game_ids, int n,
                  diee_mcts_cfg cfg, uint64_t seed, uint32_t epoch) {
    __shared__ WarpSlab slabs[AW];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, g = blockIdx.x * AW + wib;
    if (g >= n) return;
    const size_t base = (size_t)g * P.max_nodes;
    BgWarp root;
    bg_load(root, states + g, lane);
    bg_store(root, reinterpret_cast<diee_bg_state *>(P.state) + base, lane);
    if (lane == 0) {
        P.parent[base] = -1; P.first[base] = -1; P.nchild[base] = 0;
        P.visits[base] = 1.0f;  // root.visits = 1 (:123)
        P.value[base] = 0.f; P.prior[base] = 0.f; P.action[base] = SEQ_EMPTY;
        P.n_nodes[g] = 1;
        P.sel_game[g] = 0;  // selected_nodes_idxs = vec![0; n] (:142): arena node 0 for every game (Q9)
        P.sel_node[g] = 0;
        P.status[g] = DIEE_OK;
    }
    __syncwarp();
    int status = DIEE_OK;
    alpha_expand(P, g, 0, P.policy + (size_t)g * DIEE_ACTION_SPACE, P.dirichlet, cfg.dirichlet_epsilon, seed, game_ids[g], epoch,
                 slabs[wib], lane, status);
    if (lane == 0 && status != DIEE_OK) P.status[g] = status;
}

// selection half of one iteration (:151-170) + the batch row of this slot (:175-183)
__global__ void __launch_bounds__(AW * 32)
alpha_select_kernel(AlphaPool P, int n, diee_mcts_cfg cfg, int iter) {
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, g = blockIdx.x * AW + wib;
    if (g >= n) return;
    const size_t base = (size_t)g * P.max_nodes;
    int cur = 0;
    for (;;) {  // alpha_select_leaf_node :14-20
        const int nc = P.nchild[base + cur];
        if (nc == 0) break;
        const int first = P.first[base + cur];
        const float s = __fsqrt_rn(P.visits[base + cur]);
        float bs = -INFINITY;
        int bi = -1;
        for (int k0 = 0; k0 < nc; k0 += 32) {
            const int k = k0 + lane;
            if (k < nc) {
                const size_t ch = base + first + k;
                const float vis = P.visits[ch], val = P.value[ch], pri = P.prior[ch];
                // Node::alpha_ucb node.rs:98-112: q + (c * (sqrt(parent.visits) / (visits + 1))) * policy
                const float q = vis == 0.0f ? 0.0f : __fdiv_rn(val, vis);
                const float sc = __fadd_rn(q, __fmul_rn(__fmul_rn(cfg.c, __fdiv_rn(s, __fadd_rn(vis, 1.0f))), pri));
                if (!(bs > sc)) { bs = sc; bi = first + k; }
            }
        }
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) {
            const float os = __shfl_xor_sync(FULL, bs, d);
            const int oi = __shfl_xor_sync(FULL, bi, d);
            const bool take = oi >= 0 && (bi < 0 || (oi > bi ? !(bs > os) : (os > bs)));
            if (take) { bs = os; bi = oi; }
        }
        cur = bi;
    }
    const diee_bg_state *st = reinterpret_cast<const diee_bg_state *>(P.state);
    const int off0 = st[base + cur].off[0], off1 = st[base + cur].off[1];
    const int w = off0 == 15 ? -1 : (off1 == 15 ? 1 : 0);
    if (w != 0) {
        const int rp = st[base].player;  // value w.r.t. the ROOT player (:157-163)
        alpha_backprop(P, g, cur, w == rp ? 1.0f : (w == -rp ? -1.0f : 0.0f), lane);
    } else if (lane == 0) {
        P.sel_game[g] = g;
        P.sel_node[g] = cur;
        P.any_selected[iter] = 1;
    }
    __syncwarp();
    // batch row of slot g: the state of its (possibly stale, Q9) selection
    const size_t src = (size_t)P.sel_game[g] * P.max_nodes + P.sel_node[g];
    reinterpret_cast<unsigned char *>(P.batch + g)[lane] = reinterpret_cast<const unsigned char *>(st + src)[lane];
}

// expansion/backprop half (:192-200) after the batch forward
__global__ void __launch_bounds__(AW * 32)
alpha_expand_kernel(AlphaPool P, const uint32_t *__restrict__ game_ids, int n, diee_mcts_cfg cfg, uint64_t seed, uint32_t epoch,
                    int iter) {
    __shared__ WarpSlab slabs[AW];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, g = blockIdx.x * AW + wib;
    if (g >= n) return;
    if (!P.any_selected[iter]) return;  // `if !node_selected { continue }` (:171-173)
    const int tg = P.sel_game[g], node = P.sel_node[g];
    if (tg == g) {
        int status = DIEE_OK;
        alpha_expand(P, g, node, P.policy + (size_t)g * DIEE_ACTION_SPACE, nullptr, 0.f, seed, game_ids[g], epoch, slabs[wib], lane,
                     status);
        if (lane == 0 && status != DIEE_OK) P.status[g] = status;
        alpha_backprop(P, g, node, P.value_out[g], lane);
    }
    if (g == 0) {
        // Q9: slots that still point at arena node 0 evaluate game 0's root and backpropagate into it, in slot
        // order after slot 0's own update (their expansion is a no-op: the root's moves are drained)
        for (int b0 = 0; b0 < n; b0 += 32) {
            const int s = b0 + lane;
            const bool stale = s > 0 && s < n && P.sel_game[s] == 0 && P.sel_node[s] == 0;
            const float v = stale ? P.value_out[s] : 0.f;
            unsigned m = __ballot_sync(FULL, stale);
            while (m) {
                const int b = __ffs(m) - 1;
                m &= m - 1;
                const float vb = __shfl_sync(FULL, v, b);
                if (lane == 0) {
                    P.visits[0] = __fadd_rn(P.visits[0], 1.0f);
                    P.value[0] = __fadd_rn(P.value[0], vb);
                }
            }
        }
    }
}

// ---------------------------------------------------------------- NON-PARITY throughput mode
// K leaves per game and iteration with virtual loss.  The reference selects ONE leaf per game per iteration
// (alpha_mcts.rs:149-201, quirk Q8), so a search is `iterations` strictly sequential forwards of N boards; here a step
// selects up to K leaves per game -- each descent leaves a virtual loss (visits += 1, value -= vl) on its path so that
// the next one goes elsewhere --, one forward evaluates the N x K rows, and the expansion half replaces every virtual
// loss by the real value.  iterations / K forwards of N x K boards: the same number of evaluations, K times fewer
// dependent steps.  Not reference behaviour (the quirks Q9 / Q12 are dropped too: a slot without a leaf evaluates a
// dummy row that nobody reads); with K = 1 and no terminal leaf it IS the reference's search, bit for bit
// (tests/test_gpu_alpha.py).
__global__ void __launch_bounds__(AW * 32)
alpha_select_vl_kernel(AlphaPool P, int n, diee_mcts_cfg cfg, int K, float vl, int budget) {
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, g = blockIdx.x * AW + wib;
    if (g >= n) return;
    const size_t base = (size_t)g * P.max_nodes;
    const diee_bg_state *st = reinterpret_cast<const diee_bg_state *>(P.state);
    for (int k = 0; k < K; ++k) {
        int leaf = -1;
        if (k < budget) {
            int cur = 0;
            for (;;) {
                const int nc = P.nchild[base + cur];
                if (nc == 0) break;
                const int first = P.first[base + cur];
                const float s = __fsqrt_rn(P.visits[base + cur]);
                float bs = -INFINITY;
                int bi = -1;
                for (int k0 = 0; k0 < nc; k0 += 32) {
                    const int c = k0 + lane;
                    if (c < nc) {
                        const size_t ch = base + first + c;
                        const float vis = P.visits[ch], val = P.value[ch], pri = P.prior[ch];
                        const float q = vis == 0.0f ? 0.0f : __fdiv_rn(val, vis);
                        const float sc = __fadd_rn(q, __fmul_rn(__fmul_rn(cfg.c, __fdiv_rn(s, __fadd_rn(vis, 1.0f))), pri));
                        if (!(bs > sc)) { bs = sc; bi = first + c; }
                    }
                }
#pragma unroll
                for (int d = 16; d >= 1; d >>= 1) {
                    const float os = __shfl_xor_sync(FULL, bs, d);
                    const int oi = __shfl_xor_sync(FULL, bi, d);
                    const bool take = oi >= 0 && (bi < 0 || (oi > bi ? !(bs > os) : (os > bs)));
                    if (take) { bs = os; bi = oi; }
                }
                cur = bi;
            }
            const int off0 = st[base + cur].off[0], off1 = st[base + cur].off[1];
            const int w = off0 == 15 ? -1 : (off1 == 15 ? 1 : 0);
            if (w != 0) {
                const int rp = st[base].player;
                alpha_backprop(P, g, cur, w == rp ? 1.0f : (w == -rp ? -1.0f : 0.0f), lane);  // a real result, at once
            } else if (P.first[base + cur] != -2) {  // -2: already waiting for its evaluation in this step
                leaf = cur;
                if (lane == 0) {
                    P.first[base + cur] = -2;
                    for (int i = cur; i >= 0; i = P.parent[base + i]) {  // virtual loss along the path
                        P.visits[base + i] = __fadd_rn(P.visits[base + i], 1.0f);
                        P.value[base + i] = __fsub_rn(P.value[base + i], vl);
                    }
                }
                __syncwarp();
            }
        }
        const size_t slot = (size_t)g * K + k;
        if (lane == 0) P.sel_node[slot] = leaf;
        const size_t src = base + (leaf >= 0 ? leaf : 0);  // no leaf: a dummy row (the root), never read back
        reinterpret_cast<unsigned char *>(P.batch + slot)[lane] = reinterpret_cast<const unsigned char *>(st + src)[lane];
    }
}

__global__ void __launch_bounds__(AW * 32)
alpha_expand_vl_kernel(AlphaPool P, const uint32_t *__restrict__ game_ids, int n, diee_mcts_cfg cfg, uint64_t seed, uint32_t epoch,
                       int K, float vl) {
    __shared__ WarpSlab slabs[AW];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, g = blockIdx.x * AW + wib;
    if (g >= n) return;
    const size_t base = (size_t)g * P.max_nodes;
    for (int k = 0; k < K; ++k) {
        const size_t slot = (size_t)g * K + k;
        const int node = P.sel_node[slot];
        if (node < 0) continue;
        if (lane == 0) P.first[base + node] = -1;
        __syncwarp();
        int status = DIEE_OK;
        alpha_expand(P, g, node, P.policy + slot * DIEE_ACTION_SPACE, nullptr, 0.f, seed, game_ids[g], epoch, slabs[wib], lane, status);
        if (lane == 0) {
            if (status != DIEE_OK) P.status[g] = status;
            const float v = P.value_out[slot];
            for (int i = node; i >= 0; i = P.parent[base + i])  // the visit is already counted: swap the virtual loss for v
                P.value[base + i] = __fadd_rn(__fadd_rn(P.value[base + i], vl), v);
        }
        __syncwarp();
    }
}

// root children -> (action id, move, visits) in child order (input of get_prob_tensor_parallel, utils.rs:42-58)
__global__ void __launch_bounds__(AW * 32)
alpha_root_out_kernel(AlphaPool P, int n, uint16_t *__restrict__ ids_out, uint32_t *__restrict__ moves_out,
                      float *__restrict__ visits_out, int32_t *__restrict__ counts_out) {
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, g = blockIdx.x * AW + wib;
    if (g >= n) return;
    const size_t base = (size_t)g * P.max_nodes;
    const diee_bg_state *st = reinterpret_cast<const diee_bg_state *>(P.state);
    const int nc = P.nchild[base], first = P.first[base];
    const int r0 = st[base].roll[0], r1 = st[base].roll[1];
    for (int k = lane; k < nc && k < DIEE_MAX_MOVES; k += 32) {
        const uint32_t mv = P.action[base + first + k];
        ids_out[(size_t)g * DIEE_MAX_MOVES + k] = (uint16_t)bg_encode_move(r0, r1, mv);
        moves_out[(size_t)g * DIEE_MAX_MOVES + k] = mv;
        visits_out[(size_t)g * DIEE_MAX_MOVES + k] = P.visits[base + first + k];
    }
    if (lane == 0) counts_out[g] = nc;
}

static inline int agrid(int n) { return (n + AW - 1) / AW; }

cudaError_t launch_alpha_root(cudaStream_t st, const AlphaPool &P, const diee_bg_state *states, const uint32_t *game_ids, int n,
                              const diee_mcts_cfg &cfg, uint64_t seed, uint32_t epoch) {
    alpha_root_kernel<<<agrid(n), AW * 32, 0, st>>>(P, states, game_ids, n, cfg, seed, epoch);
    return cudaGetLastError();
}
cudaError_t launch_alpha_select(cudaStream_t st, const AlphaPool &P, int n, const diee_mcts_cfg &cfg, int iter) {
    alpha_select_kernel<<<agrid(n), AW * 32, 0, st>>>(P, n, cfg, iter);
    return cudaGetLastError();
}
cudaError_t launch_alpha_expand(cudaStream_t st, const AlphaPool &P, const uint32_t *game_ids, int n, const diee_mcts_cfg &cfg,
                                uint64_t seed, uint32_t epoch, int iter) {
    alpha_expand_kernel<<<agrid(n), AW * 32, 0, st>>>(P, game_ids, n, cfg, seed, epoch, iter);
    return cudaGetLastError();
}
cudaError_t launch_alpha_select_vl(cudaStream_t st, const AlphaPool &P, int n, const diee_mcts_cfg &cfg, int K, float vl, int budget) {
    alpha_select_vl_kernel<<<agrid(n), AW * 32, 0, st>>>(P, n, cfg, K, vl, budget);
    return cudaGetLastError();
}
cudaError_t launch_alpha_expand_vl(cudaStream_t st, const AlphaPool &P, const uint32_t *game_ids, int n, const diee_mcts_cfg &cfg,
                                   uint64_t seed, uint32_t epoch, int K, float vl) {
    alpha_expand_vl_kernel<<<agrid(n), AW * 32, 0, st>>>(P, game_ids, n, cfg, seed, epoch, K, vl);
    return cudaGetLastError();
}
cudaError_t launch_alpha_root_out(cudaStream_t st, const AlphaPool &P, int n, uint16_t *ids_out, uint32_t *moves_out,
                                  float *visits_out, int32_t *counts_out) {
    alpha_root_out_kernel<<<agrid(n), AW * 32, 0, st>>>(P, n, ids_out, moves_out, visits_out, counts_out);
    return cudaGetLastError();
}

}  // namespace diee
