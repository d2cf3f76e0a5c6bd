// api.cu -- the extern "C" boundary of libdiee_cuda.so (include/diee.h).
// Host-buffer entry points stage through ctx-owned device scratch; *_dev entry points are
// stream-ordered on the ctx stream.  There is no CPU fallback anywhere in this library.
#include <cuda.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "bg_device.cuh"
#include "launchers.h"

using namespace diee;

#include "bg_pb_table.h"
#include "ctx.h"

// ---------------- SM partition for the sliced search ----------------
// DIEE_TREE_SMS=n: two green contexts split the device's SMs -- n SMs for the tree kernel (a chain of 100 dependent
// iterations per game, a few warps per SM, bound by the latency of one iteration) and the rest for the rollouts of the
// slices already expanded, so that the tree never shares an SM with rollout warps (round 1 measured that sharing costs the
// tree kernel what the overlap wins).  EXPERIMENT, off by default.  Measured on B200 (1,024 games x 100 iterations, timeline
// from a -DDIEE_TRACE build, tools/trace_sweep.sh): on 64 SMs the four tree slices end at 0.58 ms (0.46 ms unsliced on the
// whole device) -- but every rollout slice of 25,600 rollouts takes 1.15-1.37 ms, as long as the one launch over all
// 102,400 does: the rollout kernel is bound by its longest chains (~300 played plies at ~4 us per ply under load), not by
// the number of rollouts, so the search ends at 1.57-1.79 ms whatever the slicing (1.67 ms unsliced).  Bit-identical
// results (tests/test_gpu_mcts.py).  The driver entry points are looked up at run time (no link-time libcuda).
template <class F>
static bool driver_fn(const char *name, F &fn) {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess || !p) return false;
    fn = reinterpret_cast<F>(p);
    return true;
}

static void ensure_partition(diee_ctx *ctx) {
    if (ctx->part_state != 0) return;
    ctx->part_state = -1;
    int tree_sms = 0;
    if (const char *e = getenv("DIEE_TREE_SMS")) tree_sms = atoi(e);
    if (tree_sms <= 0) return;  // no partition unless asked for
    CUresult (*getDev)(CUdevice *, int) = nullptr;
    CUresult (*getRes)(CUdevice, CUdevResource *, CUdevResourceType) = nullptr;
    CUresult (*split)(CUdevResource *, unsigned int *, const CUdevResource *, CUdevResource *, unsigned int, unsigned int) = nullptr;
    CUresult (*genDesc)(CUdevResourceDesc *, CUdevResource *, unsigned int) = nullptr;
    CUresult (*gcreate)(CUgreenCtx *, CUdevResourceDesc, CUdevice, unsigned int) = nullptr;
    CUresult (*gstream)(CUstream *, CUgreenCtx, unsigned int, int) = nullptr;
    if (!driver_fn("cuDeviceGet", getDev) || !driver_fn("cuDeviceGetDevResource", getRes) || !driver_fn("cuDevSmResourceSplitByCount", split) ||
        !driver_fn("cuDevResourceGenerateDesc", genDesc) || !driver_fn("cuGreenCtxCreate", gcreate) || !driver_fn("cuGreenCtxStreamCreate", gstream))
        return;
    CUdevice dev;
    CUdevResource all, grp, rem;
    unsigned int nb = 1;
    if (getDev(&dev, ctx->device) != CUDA_SUCCESS || getRes(dev, &all, CU_DEV_RESOURCE_TYPE_SM) != CUDA_SUCCESS) return;
    if ((int)all.sm.smCount < tree_sms + 16) return;
    if (split(&grp, &nb, &all, &rem, 0, (unsigned)tree_sms) != CUDA_SUCCESS || nb != 1) return;
    CUdevResourceDesc d_tree, d_roll;
    CUgreenCtx g_tree = nullptr, g_roll = nullptr;
    if (genDesc(&d_tree, &grp, 1) != CUDA_SUCCESS || genDesc(&d_roll, &rem, 1) != CUDA_SUCCESS) return;
    if (gcreate(&g_tree, d_tree, dev, CU_GREEN_CTX_DEFAULT_STREAM) != CUDA_SUCCESS) return;
    if (gcreate(&g_roll, d_roll, dev, CU_GREEN_CTX_DEFAULT_STREAM) != CUDA_SUCCESS) return;
    CUstream s = nullptr;
    if (gstream(&s, g_tree, CU_STREAM_NON_BLOCKING, 0) != CUDA_SUCCESS) return;
    ctx->part_tree = (cudaStream_t)s;
    for (int i = 0; i < 4; ++i) {
        if (gstream(&s, g_roll, CU_STREAM_NON_BLOCKING, 0) != CUDA_SUCCESS) { ctx->part_tree = nullptr; return; }
        ctx->part_roll[i] = (cudaStream_t)s;
    }
    if (cudaEventCreateWithFlags(&ctx->part_begin, cudaEventDisableTiming) != cudaSuccess) { ctx->part_tree = nullptr; return; }
    ctx->part_gctx[0] = g_tree; ctx->part_gctx[1] = g_roll;
    ctx->part_tree_sms = (int)grp.sm.smCount; ctx->part_roll_sms = (int)rem.sm.smCount;
    ctx->part_state = 1;
}

extern "C" {

const char *diee_version(void) { return "die-e-b200 0.1 (sm_100a)"; }

int32_t diee_ctx_create(int32_t device, diee_ctx **out) {
    if (!out) return DIEE_ERR_INVALID;
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count <= 0 || device < 0 || device >= count) return DIEE_ERR_CUDA;  // no CPU fallback
    diee_ctx *ctx = new diee_ctx();
    ctx->device = device;
    if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete ctx;
        return DIEE_ERR_CUDA;
    }
    ctx->stream = ctx->own_stream;
    for (int i = 0; i < 4; ++i) {
        if (cudaStreamCreateWithFlags(&ctx->side[i], cudaStreamNonBlocking) != cudaSuccess ||
            cudaEventCreateWithFlags(&ctx->ev_tree[i], cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&ctx->ev_roll[i], cudaEventDisableTiming) != cudaSuccess) {
            delete ctx;
            return DIEE_ERR_CUDA;
        }
    }
    for (int i = 0; i < 3; ++i)
        if (cudaEventCreate(&ctx->ev_time[i]) != cudaSuccess) { delete ctx; return DIEE_ERR_CUDA; }
    {   // the pure bear-off play table: filled on the host by the lane engine itself, 1 MB on the device
        std::vector<uint32_t> index;
        std::vector<uint16_t> plays;
        pb_build_table(index, plays);
        if (cudaMalloc(&ctx->pb_index.p, index.size() * 4) != cudaSuccess || cudaMalloc(&ctx->pb_plays.p, plays.size() * 2) != cudaSuccess ||
            cudaMemcpy(ctx->pb_index.p, index.data(), index.size() * 4, cudaMemcpyHostToDevice) != cudaSuccess ||
            cudaMemcpy(ctx->pb_plays.p, plays.data(), plays.size() * 2, cudaMemcpyHostToDevice) != cudaSuccess) {
            delete ctx;
            return DIEE_ERR_CUDA;
        }
        ctx->pb_index.cap = index.size() * 4;
        ctx->pb_plays.cap = plays.size() * 2;
        // a cudaMemcpy from pageable memory may return before its DMA has landed, and ctx->stream / the side streams are
        // non-blocking: without this a search issued right after the create could read a half-written table
        if (cudaDeviceSynchronize() != cudaSuccess) { delete ctx; return DIEE_ERR_CUDA; }
    }
    *out = ctx;
    return DIEE_OK;
}

int32_t diee_ctx_destroy(diee_ctx *ctx) {
    if (!ctx) return DIEE_ERR_INVALID;
    diee_comm_destroy(ctx);
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    DevBuf *bufs[] = {&ctx->s_states, &ctx->s_moves, &ctx->s_counts, &ctx->s_ids, &ctx->s_aux, &ctx->s_out,
                      &ctx->s_players, &ctx->s_best, &ctx->s_status, &ctx->s_plies, &ctx->p_states, &ctx->p_parent,
                      &ctx->p_visits, &ctx->p_value, &ctx->p_action, &ctx->p_nmoves, &ctx->p_nnodes, &ctx->p_simnode, &ctx->p_finals, &ctx->p_result,
                      &ctx->ln_table, &ctx->a_state, &ctx->a_parent, &ctx->a_first, &ctx->a_nchild, &ctx->a_visits, &ctx->a_value,
                      &ctx->a_prior, &ctx->a_action, &ctx->a_nnodes, &ctx->a_selg, &ctx->a_seln, &ctx->a_status, &ctx->a_any,
                      &ctx->a_batch, &ctx->a_policy, &ctx->a_valueout, &ctx->a_dir, &ctx->a_states_in, &ctx->a_ids_in,
                      &ctx->a_root_ids, &ctx->a_root_moves, &ctx->a_root_visits, &ctx->a_root_counts, &ctx->a_moves_in,
                      &ctx->a_rolls_in};
    for (DevBuf *b : bufs)
        if (b->p) cudaFree(b->p);
    for (int i = 0; i < 4; ++i) {
        if (ctx->side[i]) { cudaStreamSynchronize(ctx->side[i]); cudaStreamDestroy(ctx->side[i]); }
        if (ctx->ev_tree[i]) cudaEventDestroy(ctx->ev_tree[i]);
        if (ctx->ev_roll[i]) cudaEventDestroy(ctx->ev_roll[i]);
    }
    for (int i = 0; i < 3; ++i)
        if (ctx->ev_time[i]) cudaEventDestroy(ctx->ev_time[i]);
    if (ctx->part_state == 1) {
        CUresult (*sdestroy)(CUstream) = nullptr;
        CUresult (*gdestroy)(CUgreenCtx) = nullptr;
        if (driver_fn("cuStreamDestroy", sdestroy) && driver_fn("cuGreenCtxDestroy", gdestroy)) {
            cudaStreamSynchronize(ctx->part_tree);
            sdestroy((CUstream)ctx->part_tree);
            for (int i = 0; i < 4; ++i) { cudaStreamSynchronize(ctx->part_roll[i]); sdestroy((CUstream)ctx->part_roll[i]); }
            gdestroy((CUgreenCtx)ctx->part_gctx[0]);
            gdestroy((CUgreenCtx)ctx->part_gctx[1]);
        }
        if (ctx->part_begin) cudaEventDestroy(ctx->part_begin);
    }
    if (ctx->q_head.p) cudaFree(ctx->q_head.p);
    for (DevBuf *b : {&ctx->c_counts, &ctx->c_send, &ctx->c_recv})
        if (b->p) cudaFree(b->p);
    if (ctx->pb_index.p) cudaFree(ctx->pb_index.p);
    if (ctx->pb_plays.p) cudaFree(ctx->pb_plays.p);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    delete ctx;
    return DIEE_OK;
}

int32_t diee_ctx_set_stream(diee_ctx *ctx, void *cuda_stream) {
    if (!ctx) return DIEE_ERR_INVALID;
    ctx->stream = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : ctx->own_stream;
    return DIEE_OK;
}

int32_t diee_sync(diee_ctx *ctx) {
    if (!ctx) return DIEE_ERR_INVALID;
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->stream));
    return DIEE_OK;
}

int32_t diee_search_timing(diee_ctx *ctx, float *tree_ms, float *rollout_ms) {
    if (!ctx || !tree_ms || !rollout_ms) return DIEE_ERR_INVALID;
    if (!ctx->search_timed) return fail(ctx, DIEE_ERR_INVALID, "search_timing: the last search was not a split backgammon search");
    CU(cudaSetDevice(ctx->device));
    CU(cudaEventSynchronize(ctx->ev_time[2]));
    CU(cudaEventElapsedTime(tree_ms, ctx->ev_time[0], ctx->ev_time[1]));
    CU(cudaEventElapsedTime(rollout_ms, ctx->ev_time[1], ctx->ev_time[2]));
    return DIEE_OK;
}

// plies the rollouts of the last split backgammon search actually PLAYED (lane_run_kernel counts them; the plies a
// reference-exact rollout spends passing after both sides have collected everything are resolved in closed form)
int32_t diee_search_work(diee_ctx *ctx, uint64_t *rollout_plies_played) {
    if (!ctx || !rollout_plies_played) return DIEE_ERR_INVALID;
    if (!ctx->q_head.p) return fail(ctx, DIEE_ERR_INVALID, "search_work: no search has run on this context");
    CU(cudaSetDevice(ctx->device));
    unsigned long long h[8];
    CU(cudaMemcpyAsync(h, ctx->q_head.p, sizeof h, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    *rollout_plies_played = h[1] + h[3] + h[5] + h[7];
    return DIEE_OK;
}

const char *diee_last_error(const diee_ctx *ctx) { return ctx ? ctx->err.c_str() : "null ctx"; }
int64_t diee_launch_count(const diee_ctx *ctx) { return ctx ? ctx->launches : 0; }

int32_t diee_dev_alloc(diee_ctx *ctx, uint64_t bytes, void **dptr_out) {
    if (!ctx || !dptr_out) return DIEE_ERR_INVALID;
    CU(cudaSetDevice(ctx->device));
    CU(cudaMalloc(dptr_out, bytes ? bytes : 1));
    return DIEE_OK;
}
int32_t diee_dev_free(diee_ctx *ctx, void *dptr) {
    if (!ctx) return DIEE_ERR_INVALID;
    CU(cudaSetDevice(ctx->device));
    CU(cudaFree(dptr));
    return DIEE_OK;
}
int32_t diee_dev_upload(diee_ctx *ctx, void *dptr, const void *host, uint64_t bytes) {
    if (!ctx) return DIEE_ERR_INVALID;
    CU(cudaSetDevice(ctx->device));
    CU(cudaMemcpyAsync(dptr, host, bytes, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return DIEE_OK;
}
int32_t diee_dev_download(diee_ctx *ctx, void *host, const void *dptr, uint64_t bytes) {
    if (!ctx) return DIEE_ERR_INVALID;
    CU(cudaSetDevice(ctx->device));
    CU(cudaMemcpyAsync(host, dptr, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return DIEE_OK;
}

void diee_philox(uint64_t seed, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t out[4]) {
    philox4x32_10((uint32_t)seed, (uint32_t)(seed >> 32), c0, c1, c2, c3, out);
}

// ---------------- env ----------------
int32_t diee_bg_valid_moves_dev(diee_ctx *ctx, const diee_bg_state *states, int32_t n, diee_move *moves_out,
                                int32_t *counts_out, uint16_t *ids_out) {
    if (!ctx || n < 0 || (n && (!states || !moves_out || !counts_out))) return fail(ctx, DIEE_ERR_INVALID, "bg_valid_moves: bad argument");
    CU(cudaSetDevice(ctx->device));
    CU(launch_bg_valid_moves(ctx->stream, states, n, moves_out, counts_out, ids_out));
    ctx->launches += n > 0;
    return DIEE_OK;
}

int32_t diee_bg_valid_moves(diee_ctx *ctx, const diee_bg_state *states, int32_t n, diee_move *moves_out,
                            int32_t *counts_out, uint16_t *ids_out) {
    if (!ctx || n < 0 || (n && (!states || !moves_out || !counts_out))) return fail(ctx, DIEE_ERR_INVALID, "bg_valid_moves: bad argument");
    if (n == 0) return DIEE_OK;
    CU(cudaSetDevice(ctx->device));
    const size_t nm = (size_t)n * DIEE_MAX_MOVES;
    RESERVE(ctx->s_states, sizeof(diee_bg_state) * n);
    RESERVE(ctx->s_moves, sizeof(diee_move) * nm);
    RESERVE(ctx->s_counts, sizeof(int32_t) * n);
    if (ids_out) RESERVE(ctx->s_ids, sizeof(uint16_t) * nm);
    CU(cudaMemcpyAsync(ctx->s_states.p, states, sizeof(diee_bg_state) * n, cudaMemcpyHostToDevice, ctx->stream));
    int32_t rc = diee_bg_valid_moves_dev(ctx, (const diee_bg_state *)ctx->s_states.p, n, (diee_move *)ctx->s_moves.p,
                                         (int32_t *)ctx->s_counts.p, ids_out ? (uint16_t *)ctx->s_ids.p : nullptr);
    if (rc != DIEE_OK) return rc;
    CU(cudaMemcpyAsync(counts_out, ctx->s_counts.p, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaMemcpyAsync(moves_out, ctx->s_moves.p, sizeof(diee_move) * nm, cudaMemcpyDeviceToHost, ctx->stream));
    if (ids_out) CU(cudaMemcpyAsync(ids_out, ctx->s_ids.p, sizeof(uint16_t) * nm, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return DIEE_OK;
}

int32_t diee_bg_apply_moves_dev(diee_ctx *ctx, diee_bg_state *states, const diee_move *moves,
                                const uint8_t *next_rolls, int32_t n) {
    if (!ctx || n < 0 || (n && (!states || !moves || !next_rolls))) return fail(ctx, DIEE_ERR_INVALID, "bg_apply_moves: bad argument");
    CU(cudaSetDevice(ctx->device));
    CU(launch_bg_apply(ctx->stream, states, moves, next_rolls, n));
    ctx->launches += n > 0;
    return DIEE_OK;
}

int32_t diee_bg_apply_moves(diee_ctx *ctx, diee_bg_state *states, const diee_move *moves,
                            const uint8_t *next_rolls, int32_t n) {
    if (!ctx || n < 0 || (n && (!states || !moves || !next_rolls))) return fail(ctx, DIEE_ERR_INVALID, "bg_apply_moves: bad argument");
    if (n == 0) return DIEE_OK;
    CU(cudaSetDevice(ctx->device));
    RESERVE(ctx->s_states, sizeof(diee_bg_state) * n);
    RESERVE(ctx->s_moves, sizeof(diee_move) * n);
    RESERVE(ctx->s_aux, 2 * (size_t)n);
    CU(cudaMemcpyAsync(ctx->s_states.p, states, sizeof(diee_bg_state) * n, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(ctx->s_moves.p, moves, sizeof(diee_move) * n, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(ctx->s_aux.p, next_rolls, 2 * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
    int32_t rc = diee_bg_apply_moves_dev(ctx, (diee_bg_state *)ctx->s_states.p, (const diee_move *)ctx->s_moves.p,
                                         (const uint8_t *)ctx->s_aux.p, n);
    if (rc != DIEE_OK) return rc;
    CU(cudaMemcpyAsync(states, ctx->s_states.p, sizeof(diee_bg_state) * n, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return DIEE_OK;
}

int32_t diee_bg_playout_dev(diee_ctx *ctx, const diee_bg_state *starts, int32_t n, uint64_t seed,
                            uint32_t first_game_id, int32_t round_limit, int8_t *winners_out,
                            int32_t *plies_out, diee_bg_state *finals_out) {
    if (!ctx || n < 0 || round_limit < 0 || (n && (!starts || !winners_out || !plies_out)))
        return fail(ctx, DIEE_ERR_INVALID, "bg_playout: bad argument");
    if (((uintptr_t)starts | (uintptr_t)finals_out) & 15u)  // states move as two 16-byte vectors per lane
        return fail(ctx, DIEE_ERR_INVALID, "bg_playout: device state arrays must be 16-byte aligned");
    CU(cudaSetDevice(ctx->device));
    if (n == 0) return DIEE_OK;
    int nl = 0;
    RESERVE(ctx->q_head, sizeof(unsigned long long) * 8);
    CU(cudaMemsetAsync(ctx->q_head.p, 0, sizeof(unsigned long long) * 8, ctx->stream));
    CU(launch_bg_playout(ctx->stream, starts, n, seed, first_game_id, round_limit, winners_out, plies_out, finals_out,
                         (unsigned long long *)ctx->q_head.p, PbTable{(const uint32_t *)ctx->pb_index.p, (const uint16_t *)ctx->pb_plays.p}, &nl));
    ctx->launches += nl;
    return DIEE_OK;
}

int32_t diee_bg_playout(diee_ctx *ctx, const diee_bg_state *starts, int32_t n, uint64_t seed,
                        uint32_t first_game_id, int32_t round_limit, int8_t *winners_out,
                        int32_t *plies_out, diee_bg_state *finals_out) {
    if (!ctx || n < 0 || round_limit < 0 || (n && (!starts || !winners_out || !plies_out)))
        return fail(ctx, DIEE_ERR_INVALID, "bg_playout: bad argument");
    if (n == 0) return DIEE_OK;
    CU(cudaSetDevice(ctx->device));
    for (int i = 0; i < n; ++i)
        if (starts[i].roll[0] == 0 && starts[i].roll[1] == 0) return fail(ctx, DIEE_ERR_NOT_ROLLED, "bg_playout: state %d has not been rolled", i);
    RESERVE(ctx->s_states, sizeof(diee_bg_state) * n);
    RESERVE(ctx->s_out, sizeof(diee_bg_state) * n);
    RESERVE(ctx->s_counts, sizeof(int32_t) * n);
    RESERVE(ctx->s_aux, (size_t)n);
    CU(cudaMemcpyAsync(ctx->s_states.p, starts, sizeof(diee_bg_state) * n, cudaMemcpyHostToDevice, ctx->stream));
    int32_t rc = diee_bg_playout_dev(ctx, (const diee_bg_state *)ctx->s_states.p, n, seed, first_game_id, round_limit,
                                     (int8_t *)ctx->s_aux.p, (int32_t *)ctx->s_counts.p,
                                     finals_out ? (diee_bg_state *)ctx->s_out.p : nullptr);
    if (rc != DIEE_OK) return rc;
    CU(cudaMemcpyAsync(winners_out, ctx->s_aux.p, (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaMemcpyAsync(plies_out, ctx->s_counts.p, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, ctx->stream));
    if (finals_out) CU(cudaMemcpyAsync(finals_out, ctx->s_out.p, sizeof(diee_bg_state) * n, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return DIEE_OK;
}

int32_t diee_bg_encode_moves(diee_ctx *ctx, const diee_bg_state *states, const diee_move *moves, int32_t n,
                             uint16_t *ids_out) {
    if (!ctx || n < 0 || (n && (!states || !moves || !ids_out))) return fail(ctx, DIEE_ERR_INVALID, "bg_encode_moves: bad argument");
    if (n == 0) return DIEE_OK;
    CU(cudaSetDevice(ctx->device));
    RESERVE(ctx->s_states, sizeof(diee_bg_state) * n);
    RESERVE(ctx->s_moves, sizeof(diee_move) * n);
    RESERVE(ctx->s_ids, sizeof(uint16_t) * n);
    CU(cudaMemcpyAsync(ctx->s_states.p, states, sizeof(diee_bg_state) * n, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(ctx->s_moves.p, moves, sizeof(diee_move) * n, cudaMemcpyHostToDevice, ctx->stream));
    CU(launch_bg_encode_moves(ctx->stream, (const diee_bg_state *)ctx->s_states.p, (const diee_move *)ctx->s_moves.p, n, (uint16_t *)ctx->s_ids.p));
    ctx->launches += 1;
    CU(cudaMemcpyAsync(ids_out, ctx->s_ids.p, sizeof(uint16_t) * n, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return DIEE_OK;
}

int32_t diee_bg_decode_moves(diee_ctx *ctx, const diee_bg_state *states, const uint16_t *ids, int32_t n,
                             diee_move *moves_out) {
    if (!ctx || n < 0 || (n && (!states || !ids || !moves_out))) return fail(ctx, DIEE_ERR_INVALID, "bg_decode_moves: bad argument");
    if (n == 0) return DIEE_OK;
    for (int i = 0; i < n; ++i)
        if (ids[i] >= DIEE_ACTION_SPACE) return fail(ctx, DIEE_ERR_INVALID, "bg_decode_moves: action id %u out of range", (unsigned)ids[i]);
    CU(cudaSetDevice(ctx->device));
    RESERVE(ctx->s_states, sizeof(diee_bg_state) * n);
    RESERVE(ctx->s_moves, sizeof(diee_move) * n);
    RESERVE(ctx->s_ids, sizeof(uint16_t) * n);
    CU(cudaMemcpyAsync(ctx->s_states.p, states, sizeof(diee_bg_state) * n, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(ctx->s_ids.p, ids, sizeof(uint16_t) * n, cudaMemcpyHostToDevice, ctx->stream));
    CU(launch_bg_decode_moves(ctx->stream, (const diee_bg_state *)ctx->s_states.p, (const uint16_t *)ctx->s_ids.p, n, (diee_move *)ctx->s_moves.p));
    ctx->launches += 1;
    CU(cudaMemcpyAsync(moves_out, ctx->s_moves.p, sizeof(diee_move) * n, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return DIEE_OK;
}

int32_t diee_bg_encode_states_dev(diee_ctx *ctx, const diee_bg_state *states, int32_t n, float *out) {
    if (!ctx || n < 0 || (n && (!states || !out))) return fail(ctx, DIEE_ERR_INVALID, "bg_encode_states: bad argument");
    CU(cudaSetDevice(ctx->device));
    CU(launch_bg_encode_states(ctx->stream, states, n, out));
    ctx->launches += n > 0;
    return DIEE_OK;
}

int32_t diee_bg_encode_states(diee_ctx *ctx, const diee_bg_state *states, int32_t n, float *out) {
    if (!ctx || n < 0 || (n && (!states || !out))) return fail(ctx, DIEE_ERR_INVALID, "bg_encode_states: bad argument");
    if (n == 0) return DIEE_OK;
    for (int i = 0; i < n; ++i)
        if (states[i].roll[0] == 0 && states[i].roll[1] == 0) return fail(ctx, DIEE_ERR_NOT_ROLLED, "bg_encode_states: state %d has not been rolled", i);
    CU(cudaSetDevice(ctx->device));
    RESERVE(ctx->s_states, sizeof(diee_bg_state) * n);
    RESERVE(ctx->s_out, sizeof(float) * 144 * (size_t)n);
    CU(cudaMemcpyAsync(ctx->s_states.p, states, sizeof(diee_bg_state) * n, cudaMemcpyHostToDevice, ctx->stream));
    int32_t rc = diee_bg_encode_states_dev(ctx, (const diee_bg_state *)ctx->s_states.p, n, (float *)ctx->s_out.p);
    if (rc != DIEE_OK) return rc;
    CU(cudaMemcpyAsync(out, ctx->s_out.p, sizeof(float) * 144 * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return DIEE_OK;
}

// ---------------- pure MCTS ----------------
static size_t state_size(int game_kind) { return game_kind == DIEE_GAME_BACKGAMMON ? sizeof(diee_bg_state) : sizeof(diee_ttt_state); }

static int32_t ensure_pool(diee_ctx *ctx, int game_kind, int n, const diee_mcts_cfg *cfg) {
    const size_t cap = (size_t)cfg->iterations + 1, total = cap * (size_t)n;
    RESERVE(ctx->p_states, state_size(game_kind) * total);
    RESERVE(ctx->p_parent, sizeof(int32_t) * total);
    RESERVE(ctx->p_visits, sizeof(float) * total);
    RESERVE(ctx->p_value, sizeof(float) * total);
    RESERVE(ctx->p_action, sizeof(uint32_t) * total);
    RESERVE(ctx->p_nmoves, sizeof(uint32_t) * total);
    RESERVE(ctx->p_nnodes, sizeof(int32_t) * (size_t)n);
    RESERVE(ctx->p_simnode, sizeof(int32_t) * (size_t)cfg->iterations * (size_t)n);
    RESERVE(ctx->p_finals, state_size(game_kind) * (size_t)cfg->iterations * (size_t)n);
    RESERVE(ctx->p_result, sizeof(float) * (size_t)n);
    if (ctx->ln_table_n < cfg->iterations + 2) {
        // ln of every possible (integer-valued) visit count, correctly rounded from double:
        // the contract's replacement for f32::ln (node.rs:91)
        const uint32_t m = cfg->iterations + 2;
        std::vector<float> t(m);
        t[0] = -INFINITY;
        for (uint32_t i = 1; i < m; ++i) t[i] = (float)std::log((double)i);
        RESERVE(ctx->ln_table, sizeof(float) * m);
        CU(cudaMemcpyAsync(ctx->ln_table.p, t.data(), sizeof(float) * m, cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        ctx->ln_table_n = m;
    }
    return DIEE_OK;
}

static int32_t check_mcts_args(diee_ctx *ctx, int32_t game_kind, const void *states, int32_t n, const int8_t *players,
                               const diee_mcts_cfg *cfg, void *best, int32_t *status, uint32_t epoch) {
    if (!ctx) return DIEE_ERR_INVALID;
    if (game_kind != DIEE_GAME_BACKGAMMON && game_kind != DIEE_GAME_TICTACTOE) return fail(ctx, DIEE_ERR_INVALID, "mcts_search: unknown game kind %d", game_kind);
    if (n < 0 || !cfg || (n && (!states || !players || !best || !status))) return fail(ctx, DIEE_ERR_INVALID, "mcts_search: bad argument");
    if (cfg->iterations == 0 || cfg->iterations > 65534u) return fail(ctx, DIEE_ERR_INVALID, "mcts_search: iterations must be in 1..65534");
    if (epoch > 0xFFFFu) return fail(ctx, DIEE_ERR_INVALID, "mcts_search: epoch must be < 65536");
    return DIEE_OK;
}

// dump: the caller will read the node pool back, so every node's legal-move count must be materialised
static int32_t mcts_search_dev_impl(diee_ctx *ctx, int32_t game_kind, const void *states, int32_t n,
                                    const int8_t *players, const diee_mcts_cfg *cfg, uint64_t seed,
                                    uint32_t first_game_id, uint32_t epoch, void *best_moves_out,
                                    int32_t *status_out, diee_search_stats *stats_dev, bool dump) {
    int32_t rc = check_mcts_args(ctx, game_kind, states, n, players, cfg, best_moves_out, status_out, epoch);
    if (rc != DIEE_OK) return rc;
    if (n == 0) return DIEE_OK;
    CU(cudaSetDevice(ctx->device));
    rc = ensure_pool(ctx, game_kind, n, cfg);
    if (rc != DIEE_OK) return rc;
    PoolPtrs pp{ctx->p_states.p, (int32_t *)ctx->p_parent.p, (float *)ctx->p_visits.p, (float *)ctx->p_value.p,
                (uint32_t *)ctx->p_action.p, (uint32_t *)ctx->p_nmoves.p, (int32_t *)ctx->p_nnodes.p,
                (int32_t *)ctx->p_simnode.p, ctx->p_finals.p,
                PbTable{(const uint32_t *)ctx->pb_index.p, (const uint16_t *)ctx->pb_plays.p}, (float *)ctx->p_result.p};
    RESERVE(ctx->q_head, sizeof(unsigned long long) * 8);
    CU(cudaMemsetAsync(ctx->q_head.p, 0, sizeof(unsigned long long) * 8, ctx->stream));
    SearchPipe pipe;
    for (int i = 0; i < SEARCH_SLICES; ++i) { pipe.side[i] = ctx->side[i]; pipe.tree_done[i] = ctx->ev_tree[i]; pipe.roll_done[i] = ctx->ev_roll[i]; }
    pipe.queue_heads = (unsigned long long *)ctx->q_head.p;
    pipe.t_begin = ctx->ev_time[0]; pipe.t_tree = ctx->ev_time[1]; pipe.t_end = ctx->ev_time[2];
    pipe.timed = &ctx->search_timed;
    pipe.part_tree = nullptr; pipe.part_begin = nullptr;
    for (int i = 0; i < SEARCH_SLICES; ++i) pipe.part_roll[i] = nullptr;
    if (game_kind == DIEE_GAME_BACKGAMMON && !(cfg->mode_flags & DIEE_MODE_ROLLOUT_CHECK_CURRENT)) {
        ensure_partition(ctx);
        if (ctx->part_state == 1) {
            pipe.part_tree = ctx->part_tree; pipe.part_begin = ctx->part_begin;
            for (int i = 0; i < SEARCH_SLICES; ++i) pipe.part_roll[i] = ctx->part_roll[i];
        }
    }
    const size_t pairs = (size_t)cfg->iterations * (size_t)n;
    CU(cudaMemsetAsync(ctx->p_simnode.p, 0xFF, sizeof(int32_t) * pairs, ctx->stream));
    CU(cudaMemsetAsync(ctx->p_finals.p, 0, state_size(game_kind) * pairs, ctx->stream));
    uint32_t *best32 = (uint32_t *)best_moves_out;
    int nl = 0;
    if (game_kind == DIEE_GAME_TICTACTOE) {  // kernel writes u32 per game; narrow to u8 afterwards
        RESERVE(ctx->s_best, sizeof(uint32_t) * (size_t)n);
        best32 = (uint32_t *)ctx->s_best.p;
    }
    CU(launch_mcts_search(ctx->stream, game_kind, states, n, players, *cfg, seed, first_game_id, epoch, pp, pipe,
                          (const float *)ctx->ln_table.p, best32, status_out, stats_dev, dump, &nl));
    ctx->launches += nl;
    if (game_kind == DIEE_GAME_TICTACTOE) {
        // EMPTY_MOVE = 10 (tictactoe/mod.rs:18); done on the host side of the stream for this tiny case
        std::vector<uint32_t> h((size_t)n);
        CU(cudaMemcpyAsync(h.data(), best32, sizeof(uint32_t) * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        std::vector<uint8_t> b((size_t)n);
        for (int i = 0; i < n; ++i) b[i] = h[i] == SEQ_EMPTY ? 10 : (uint8_t)(h[i] & 0xFFu);
        CU(cudaMemcpyAsync(best_moves_out, b.data(), (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
    }
    return DIEE_OK;
}

int32_t diee_mcts_search_dev(diee_ctx *ctx, int32_t game_kind, const void *states, int32_t n,
                             const int8_t *players, const diee_mcts_cfg *cfg, uint64_t seed,
                             uint32_t first_game_id, uint32_t epoch, void *best_moves_out,
                             int32_t *status_out, diee_search_stats *stats_dev) {
    return mcts_search_dev_impl(ctx, game_kind, states, n, players, cfg, seed, first_game_id, epoch, best_moves_out, status_out,
                                stats_dev, false);
}

int32_t diee_mcts_search(diee_ctx *ctx, int32_t game_kind, const void *states, int32_t n,
                         const int8_t *players, const diee_mcts_cfg *cfg, uint64_t seed,
                         uint32_t first_game_id, uint32_t epoch, void *best_moves_out,
                         int32_t *status_out, diee_node *nodes_out, void *node_states_out,
                         int32_t *n_nodes_out, diee_search_stats *stats_out, void *rollout_finals_out) {
    int32_t rc = check_mcts_args(ctx, game_kind, states, n, players, cfg, best_moves_out, status_out, epoch);
    if (rc != DIEE_OK) return rc;
    if (n == 0) return DIEE_OK;
    CU(cudaSetDevice(ctx->device));
    const size_t ss = state_size(game_kind);
    if (game_kind == DIEE_GAME_BACKGAMMON) {
        const diee_bg_state *s = (const diee_bg_state *)states;
        for (int i = 0; i < n; ++i)
            if (s[i].roll[0] == 0 && s[i].roll[1] == 0) return fail(ctx, DIEE_ERR_NOT_ROLLED, "mcts_search: state %d has not been rolled", i);
    }
    const size_t best_sz = game_kind == DIEE_GAME_BACKGAMMON ? sizeof(diee_move) : 1;
    RESERVE(ctx->s_states, ss * n);
    RESERVE(ctx->s_players, (size_t)n);
    RESERVE(ctx->s_moves, sizeof(uint32_t) * (size_t)n);
    RESERVE(ctx->s_status, sizeof(int32_t) * (size_t)n);
    RESERVE(ctx->s_plies, sizeof(diee_search_stats) * (size_t)n);
    CU(cudaMemcpyAsync(ctx->s_states.p, states, ss * n, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(ctx->s_players.p, players, (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
    rc = mcts_search_dev_impl(ctx, game_kind, ctx->s_states.p, n, (const int8_t *)ctx->s_players.p, cfg, seed, first_game_id,
                              epoch, ctx->s_moves.p, (int32_t *)ctx->s_status.p, (diee_search_stats *)ctx->s_plies.p,
                              nodes_out != nullptr);
    if (rc != DIEE_OK) return rc;
    CU(cudaMemcpyAsync(best_moves_out, ctx->s_moves.p, best_sz * n, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaMemcpyAsync(status_out, ctx->s_status.p, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, ctx->stream));
    if (stats_out) CU(cudaMemcpyAsync(stats_out, ctx->s_plies.p, sizeof(diee_search_stats) * n, cudaMemcpyDeviceToHost, ctx->stream));
    if (n_nodes_out) CU(cudaMemcpyAsync(n_nodes_out, ctx->p_nnodes.p, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, ctx->stream));
    const size_t total = ((size_t)cfg->iterations + 1) * (size_t)n;
    if (node_states_out) CU(cudaMemcpyAsync(node_states_out, ctx->p_states.p, ss * total, cudaMemcpyDeviceToHost, ctx->stream));
    if (rollout_finals_out)
        CU(cudaMemcpyAsync(rollout_finals_out, ctx->p_finals.p, ss * (size_t)cfg->iterations * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    std::vector<int32_t> parent;
    std::vector<float> visits, value;
    std::vector<uint32_t> action, nmoves;
    if (nodes_out) {
        parent.resize(total); visits.resize(total); value.resize(total); action.resize(total); nmoves.resize(total);
        CU(cudaMemcpyAsync(parent.data(), ctx->p_parent.p, 4 * total, cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaMemcpyAsync(visits.data(), ctx->p_visits.p, 4 * total, cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaMemcpyAsync(value.data(), ctx->p_value.p, 4 * total, cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaMemcpyAsync(action.data(), ctx->p_action.p, 4 * total, cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaMemcpyAsync(nmoves.data(), ctx->p_nmoves.p, 4 * total, cudaMemcpyDeviceToHost, ctx->stream));
    }
    CU(cudaStreamSynchronize(ctx->stream));
    if (nodes_out) {
        for (size_t i = 0; i < total; ++i) {
            nodes_out[i].parent = parent[i];
            nodes_out[i].visits = visits[i];
            nodes_out[i].value = value[i];
            memcpy(&nodes_out[i].action, &action[i], 4);
            nodes_out[i].n_moves = (int32_t)(nmoves[i] >> 16);
            nodes_out[i].n_untried = (int32_t)(nmoves[i] & 0xFFFFu);
        }
    }
    return DIEE_OK;
}

}  // extern "C"
