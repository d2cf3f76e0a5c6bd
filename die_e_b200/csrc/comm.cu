// comm.cu -- the one exchange step of the path: finished self-play trajectories are all-gathered
// into every rank's replay buffer, and a promoted model's weights are broadcast, over NCCL (NVLink 5 / NVSwitch on a
// B200 node).  One rank per context / GPU.  Games never interact (the reference runs them as independent rayon
// tasks, versus.rs:303-316), so nothing else on the path communicates.
//
// NCCL is bound at run time (dlopen of libnccl.so.2), not at link time: the library must load on a box that has no
// NCCL at all, and inside a process where torch has already loaded its own copy (that copy is reused; a process
// that will import torch LATER should point DIEE_NCCL_LIB at torch's bundled libnccl.so.2 -- the Python host does
// -- because the dynamic loader hands torch whichever libnccl.so.2 is already mapped).
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <cstdlib>
#include <cstring>
#include <vector>

#include "ctx.h"

namespace {

typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { ncclSuccess = 0 };
enum { ncclInt8 = 0, ncclChar = 0, ncclUint8 = 1, ncclFloat32 = 7, ncclInt64 = 4 };

struct Nccl {
    void *h = nullptr;
    int (*GetUniqueId)(ncclUniqueId *) = nullptr;
    int (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*AllGather)(const void *, void *, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*Broadcast)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
};

Nccl *nccl() {
    static Nccl n;
    static bool tried = false;
    if (!tried) {
        tried = true;
        // the copy the process already has (torch's bundled one, normally) wins; then DIEE_NCCL_LIB; then the
        // system's.  Never RTLD_GLOBAL: symbols of an NCCL loaded here must not leak into libraries loaded later.
        n.h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_LOCAL);
        if (!n.h) { const char *path = getenv("DIEE_NCCL_LIB"); if (path && *path) n.h = dlopen(path, RTLD_NOW | RTLD_LOCAL); }
        if (!n.h) n.h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
        if (!n.h) n.h = dlopen("libnccl.so", RTLD_NOW | RTLD_LOCAL);
        if (n.h) {
            n.GetUniqueId = (int (*)(ncclUniqueId *))dlsym(n.h, "ncclGetUniqueId");
            n.CommInitRank = (int (*)(ncclComm_t *, int, ncclUniqueId, int))dlsym(n.h, "ncclCommInitRank");
            n.CommDestroy = (int (*)(ncclComm_t))dlsym(n.h, "ncclCommDestroy");
            n.AllGather = (int (*)(const void *, void *, size_t, int, ncclComm_t, cudaStream_t))dlsym(n.h, "ncclAllGather");
            n.Broadcast = (int (*)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t))dlsym(n.h, "ncclBroadcast");
            n.GetErrorString = (const char *(*)(int))dlsym(n.h, "ncclGetErrorString");
            if (!n.GetUniqueId || !n.CommInitRank || !n.CommDestroy || !n.AllGather || !n.Broadcast) n.h = nullptr;
        }
    }
    return n.h ? &n : nullptr;
}

#define NC(call)                                                                                                   \
    do {                                                                                                           \
        int r_ = (call);                                                                                           \
        if (r_ != ncclSuccess)                                                                                     \
            return fail(ctx, DIEE_ERR_CUDA, "%s: %s", #call, N->GetErrorString ? N->GetErrorString(r_) : "NCCL error"); \
    } while (0)

}  // namespace

extern "C" {

int32_t diee_comm_unique_id(uint8_t *id_out) {
    Nccl *N = nccl();
    if (!N || !id_out) return DIEE_ERR_CUDA;
    ncclUniqueId id;
    if (N->GetUniqueId(&id) != ncclSuccess) return DIEE_ERR_CUDA;
    memcpy(id_out, &id, DIEE_COMM_ID_BYTES);
    return DIEE_OK;
}

int32_t diee_comm_init(diee_ctx *ctx, int32_t nranks, int32_t rank, const uint8_t *id) {
    if (!ctx || !id || nranks < 1 || rank < 0 || rank >= nranks) return fail(ctx, DIEE_ERR_INVALID, "comm_init: bad argument");
    Nccl *N = nccl();
    if (!N) return fail(ctx, DIEE_ERR_CUDA, "comm_init: libnccl.so.2 could not be loaded");
    if (ctx->comm) return fail(ctx, DIEE_ERR_INVALID, "comm_init: the context already has a communicator");
    CU(cudaSetDevice(ctx->device));
    ncclUniqueId uid;
    memcpy(&uid, id, DIEE_COMM_ID_BYTES);
    ncclComm_t c = nullptr;
    // scratch of the counts / status gathers of diee_traj_allgather: allocated HERE so that nothing rank-local can
    // fail between entering that call and its first collective
    RESERVE(ctx->c_counts, sizeof(long long) * 6 * (size_t)(nranks + 2));
    NC(N->CommInitRank(&c, nranks, uid, rank));
    ctx->comm = c;
    ctx->comm_ranks = nranks;
    ctx->comm_rank = rank;
    return DIEE_OK;
}

int32_t diee_comm_destroy(diee_ctx *ctx) {
    if (!ctx) return DIEE_ERR_INVALID;
    Nccl *N = ctx->comm ? nccl() : nullptr;  // a context that never joined a communicator must not load NCCL on its way out
    if (ctx->comm && N) {
        cudaSetDevice(ctx->device);
        cudaStreamSynchronize(ctx->stream);
        N->CommDestroy((ncclComm_t)ctx->comm);
    }
    ctx->comm = nullptr;
    ctx->comm_ranks = 0;
    return DIEE_OK;
}

// Ragged sizes: the counts first, then slabs padded to the largest contribution; the host keeps each rank's valid prefix,
// rank-major, and rebases pi_offset into the concatenated pi arrays.
//
// Every decision between the collectives is COLLECTIVE: caps, output pointers and local failures are rank-local
// facts, and a rank that bailed out alone would leave the others inside the next ncclAllGather forever.  So each rank
// contributes (n_rec, n_pi, rec_cap, pi_cap, outputs present) to the first gather and all of them derive the same
// verdict from the same table; a rank whose staging (allocation, H2D copy) fails still takes part in a one-word
// status gather and in nothing after it, and every rank returns an error in that case.
int32_t diee_traj_allgather(diee_ctx *ctx, const diee_traj_record *rec, int32_t n_rec, const uint16_t *pi_ids, const float *pi_vals,
                            int32_t n_pi, diee_traj_record *rec_out, int32_t rec_cap, uint16_t *pi_ids_out, float *pi_vals_out,
                            int32_t pi_cap, int32_t *n_rec_out, int32_t *n_pi_out) {
    if (!ctx) return DIEE_ERR_INVALID;
    Nccl *N = nccl();
    if (!N || !ctx->comm) return fail(ctx, DIEE_ERR_INVALID, "traj_allgather: diee_comm_init has not been called on this context");
    // a locally invalid call still enters the first gather (with the `bad` flag up) so that nobody waits for it
    const bool bad_in = n_rec < 0 || n_pi < 0 || !n_rec_out || !n_pi_out || (n_rec > 0 && !rec) || (n_pi > 0 && (!pi_ids || !pi_vals)) ||
                        rec_cap < 0 || pi_cap < 0;
    CU(cudaSetDevice(ctx->device));
    ncclComm_t comm = (ncclComm_t)ctx->comm;
    const int W = ctx->comm_ranks;
    cudaStream_t st = ctx->stream;
    constexpr int F = 6;  // n_rec, n_pi, rec_cap, pi_cap, has rec_out, has pi outputs (all -1 in a row = bad call on that rank)
    // (the buffers below are allocated by diee_comm_init, so nothing here can fail before the first collective)
    if (ctx->c_counts.cap < sizeof(long long) * F * (size_t)(W + 2)) return fail(ctx, DIEE_ERR_INVALID, "traj_allgather: communicator scratch missing");
    long long mine[F] = {n_rec, n_pi, rec_cap, pi_cap, rec_out ? 1 : 0, (pi_ids_out && pi_vals_out) ? 1 : 0};
    if (bad_in) for (int i = 0; i < F; ++i) mine[i] = -1;
    long long *d_mine = (long long *)ctx->c_counts.p, *d_all = d_mine + F;
    std::vector<long long> all((size_t)F * W);
    int local_err = 0;
    if (cudaMemcpyAsync(d_mine, mine, sizeof mine, cudaMemcpyHostToDevice, st) != cudaSuccess) local_err = 1;
    NC(N->AllGather(d_mine, d_all, F, ncclInt64, comm, st));
    CU(cudaMemcpyAsync(all.data(), d_all, sizeof(long long) * F * W, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    long long max_rec = 1, max_pi = 1, tot_rec = 0, tot_pi = 0;
    int bad_rank = -1;
    for (int r = 0; r < W; ++r) {
        const long long *a = &all[(size_t)F * r];
        if (a[0] < 0) { if (bad_rank < 0) bad_rank = r; continue; }
        max_rec = a[0] > max_rec ? a[0] : max_rec;
        max_pi = a[1] > max_pi ? a[1] : max_pi;
        tot_rec += a[0];
        tot_pi += a[1];
    }
    if (bad_rank >= 0) return fail(ctx, DIEE_ERR_INVALID, "traj_allgather: bad argument on rank %d", bad_rank);
    *n_rec_out = (int32_t)tot_rec;
    *n_pi_out = (int32_t)tot_pi;
    // the same table on every rank -> the same verdict on every rank
    for (int r = 0; r < W; ++r) {
        const long long *a = &all[(size_t)F * r];
        if (tot_rec > a[2] || tot_pi > a[3])
            return fail(ctx, DIEE_ERR_OVERFLOW, "traj_allgather: %lld records / %lld pi entries do not fit rank %d's buffers (%lld / %lld)",
                        tot_rec, tot_pi, r, a[2], a[3]);
        if ((tot_rec && !a[4]) || (tot_pi && !a[5])) return fail(ctx, DIEE_ERR_INVALID, "traj_allgather: null output on rank %d", r);
    }
    // slabs: [records | ids | values] of one rank, padded.  Staging may fail locally (allocation): gather one status
    // word before the payload so that all ranks skip it together.
    const size_t b_rec = sizeof(diee_traj_record) * (size_t)max_rec, b_ids = (2 * (size_t)max_pi + 3) & ~(size_t)3, b_val = 4 * (size_t)max_pi;
    const size_t slab = b_rec + b_ids + b_val;
    if (!local_err && (reserve(ctx, ctx->c_send, slab) != DIEE_OK || reserve(ctx, ctx->c_recv, slab * (size_t)W) != DIEE_OK)) local_err = 1;
    unsigned char *d_send = (unsigned char *)ctx->c_send.p, *d_recv = (unsigned char *)ctx->c_recv.p;
    if (!local_err) {
        if (n_rec && cudaMemcpyAsync(d_send, rec, sizeof(diee_traj_record) * (size_t)n_rec, cudaMemcpyHostToDevice, st) != cudaSuccess) local_err = 1;
        if (n_pi && (cudaMemcpyAsync(d_send + b_rec, pi_ids, 2 * (size_t)n_pi, cudaMemcpyHostToDevice, st) != cudaSuccess ||
                     cudaMemcpyAsync(d_send + b_rec + b_ids, pi_vals, 4 * (size_t)n_pi, cudaMemcpyHostToDevice, st) != cudaSuccess)) local_err = 1;
    }
    long long flag = local_err;
    long long *d_flag = d_all + (size_t)F * W, *d_flags = d_mine;  // d_mine..d_all is free again: W <= F*W words
    std::vector<long long> flags((size_t)W);
    CU(cudaMemcpyAsync(d_flag, &flag, sizeof flag, cudaMemcpyHostToDevice, st));
    // (gathers into the first W words of the scratch; the per-rank table was already copied to the host)
    NC(N->AllGather(d_flag, d_flags, 1, ncclInt64, comm, st));
    CU(cudaMemcpyAsync(flags.data(), d_flags, sizeof(long long) * W, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    for (int r = 0; r < W; ++r)
        if (flags[r]) return fail(ctx, DIEE_ERR_CUDA, "traj_allgather: staging failed on rank %d (device memory?)", r);
    NC(N->AllGather(d_send, d_recv, slab, ncclChar, comm, st));
    long long o_rec = 0, o_pi = 0;
    for (int r = 0; r < W; ++r) {
        const long long nr = all[(size_t)F * r], np = all[(size_t)F * r + 1];
        const unsigned char *src = d_recv + slab * (size_t)r;
        if (nr) CU(cudaMemcpyAsync(rec_out + o_rec, src, sizeof(diee_traj_record) * (size_t)nr, cudaMemcpyDeviceToHost, st));
        if (np) {
            CU(cudaMemcpyAsync(pi_ids_out + o_pi, src + b_rec, 2 * (size_t)np, cudaMemcpyDeviceToHost, st));
            CU(cudaMemcpyAsync(pi_vals_out + o_pi, src + b_rec + b_ids, 4 * (size_t)np, cudaMemcpyDeviceToHost, st));
        }
        o_rec += nr;
        o_pi += np;
    }
    CU(cudaStreamSynchronize(st));
    o_rec = 0; o_pi = 0;
    for (int r = 0; r < W; ++r) {  // pi_offset of rank r's records now counts from the start of the concatenated arrays
        for (long long i = 0; i < all[(size_t)F * r]; ++i) rec_out[o_rec + i].pi_offset += (uint32_t)o_pi;
        o_rec += all[(size_t)F * r];
        o_pi += all[(size_t)F * r + 1];
    }
    return DIEE_OK;
}

// every rank passes the same tensor list (host f32 arrays); on return all hold rank `root`'s values
int32_t diee_net_broadcast(diee_ctx *ctx, float *const *tensors, const int64_t *numels, int32_t n_tensors, int32_t root) {
    if (!ctx || !tensors || !numels || n_tensors < 0) return fail(ctx, DIEE_ERR_INVALID, "net_broadcast: bad argument");
    Nccl *N = nccl();
    if (!N || !ctx->comm) return fail(ctx, DIEE_ERR_INVALID, "net_broadcast: diee_comm_init has not been called on this context");
    if (root < 0 || root >= ctx->comm_ranks) return fail(ctx, DIEE_ERR_INVALID, "net_broadcast: bad root");
    CU(cudaSetDevice(ctx->device));
    size_t total = 0;
    for (int i = 0; i < n_tensors; ++i) total += (size_t)numels[i];
    if (total == 0) return DIEE_OK;
    RESERVE(ctx->c_send, total * sizeof(float));
    float *d = (float *)ctx->c_send.p;
    cudaStream_t st = ctx->stream;
    if (ctx->comm_rank == root) {
        size_t off = 0;
        for (int i = 0; i < n_tensors; ++i) {
            CU(cudaMemcpyAsync(d + off, tensors[i], sizeof(float) * (size_t)numels[i], cudaMemcpyHostToDevice, st));
            off += (size_t)numels[i];
        }
    }
    NC(N->Broadcast(d, d, total, ncclFloat32, root, (ncclComm_t)ctx->comm, st));
    if (ctx->comm_rank != root) {
        size_t off = 0;
        for (int i = 0; i < n_tensors; ++i) {
            CU(cudaMemcpyAsync(tensors[i], d + off, sizeof(float) * (size_t)numels[i], cudaMemcpyDeviceToHost, st));
            off += (size_t)numels[i];
        }
    }
    CU(cudaStreamSynchronize(st));
    return DIEE_OK;
}

}  // extern "C"
