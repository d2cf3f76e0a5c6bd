// ctx.h -- internal: the context object and error helpers shared by api.cu and net.cu
#pragma once
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <string>
#include <vector>

#include "../../include/diee.h"

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
};

struct diee_ctx {
    int device = 0;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    std::string err;
    int64_t launches = 0;
    // scratch for host-buffer entry points
    DevBuf s_states, s_moves, s_counts, s_ids, s_aux, s_out, s_players, s_best, s_status, s_plies;
    // pure-MCTS node pool (HBM resident, reused between searches)
    DevBuf p_states, p_parent, p_visits, p_value, p_action, p_nmoves, p_nnodes, p_simnode, p_finals, p_result, ln_table;
    uint32_t ln_table_n = 0;
    DevBuf pb_index, pb_plays;  // pure bear-off play table (bg_pb_table.h)
    DevBuf q_head;  // job queue heads of the persistent lane kernel (lane_kernels.cu)
    // side streams / events of the sliced search (mcts_kernels.cu launch_typed)
    cudaStream_t side[4] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t ev_tree[4] = {nullptr, nullptr, nullptr, nullptr}, ev_roll[4] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t ev_time[3] = {nullptr, nullptr, nullptr};  // begin | tree done | rollouts done of the last split search
    bool search_timed = false;
    // SM partition for the sliced search (green contexts; api.cu ensure_partition): 0 = not tried, 1 = ready, -1 = unavailable
    int part_state = 0;
    void *part_gctx[2] = {nullptr, nullptr};  // CUgreenCtx: tree SMs | rollout SMs
    cudaStream_t part_tree = nullptr, part_roll[4] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t part_begin = nullptr;
    int part_tree_sms = 0, part_roll_sms = 0;
    // AlphaZero search arena + per-iteration batch buffers (alpha.cu)
    DevBuf a_state, a_parent, a_first, a_nchild, a_visits, a_value, a_prior, a_action, a_nnodes, a_selg, a_seln, a_status,
        a_any, a_batch, a_policy, a_valueout, a_dir, a_states_in, a_ids_in, a_root_ids, a_root_moves, a_root_visits,
        a_root_counts, a_moves_in, a_rolls_in;
    std::vector<float> dir_host;
    uint64_t net_evals = 0;
    // NCCL communicator of this rank (comm.cu), bound at run time
    void *comm = nullptr;
    int comm_ranks = 0, comm_rank = 0;
    DevBuf c_counts, c_send, c_recv;
};

static inline int32_t fail(diee_ctx *ctx, int32_t code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (ctx) ctx->err = buf;
    return code;
}

#define CU(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) return fail(ctx, DIEE_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e_)); \
    } while (0)

static inline int32_t reserve(diee_ctx *ctx, DevBuf &b, size_t bytes) {
    if (bytes <= b.cap) return DIEE_OK;
    if (b.p) CU(cudaFree(b.p));
    b.p = nullptr;
    b.cap = 0;
    size_t want = bytes + bytes / 4 + 256;
    CU(cudaMalloc(&b.p, want));
    b.cap = want;
    return DIEE_OK;
}
#define RESERVE(buf, bytes)                                  \
    do {                                                     \
        int32_t r_ = reserve(ctx, buf, bytes);               \
        if (r_ != DIEE_OK) return r_;                        \
    } while (0)

