// net_kernels.cu -- policy/value ResNet forward for sm_100a.
//
// Reference: src/alphazero/nnet.rs (ResNet::new :57-107, ResBlock :17-45, forward_t :120-133).
// Every 3x3 convolution over the 4x6 board is an implicit GEMM on the 5th-generation tensor
// cores:  M = boards*24 positions, N = output channels, K = 9 taps * C_in.
//   * activations live in HBM as NHWC bf16 ([board][h][w][c]); a 4-D TMA tensor map
//     (c, w, h, board) with a box of [64 ch][6][4][NB boards] (NB = 16, 8 or 4) loads, for tap (kh,kw), the box
//     shifted by (kw-1, kh-1): the out-of-bounds halo is ZERO-FILLED by TMA, so the 3x3 padding
//     costs nothing and no im2col matrix is ever materialised;
//   * weights are pre-packed [c_out][tap][c_in] bf16 with BatchNorm folded in (eval mode), loaded
//     by a 2-D tensor map; both operands land in shared memory in the 128-byte swizzle that
//     tcgen05 smem descriptors expect;
//   * one elected thread issues tcgen05.mma (M=128, N=BN, K=16) into TMEM accumulators
//     (3 M-tiles x BN fp32 columns for a 16-board tile; the tile shape is picked per launch, see ConvCfg);
//     tcgen05.commit releases smem stages and
//     finally signals the epilogue warps, which read TMEM with tcgen05.ld, add the folded bias
//     (+ residual), apply ReLU and store bf16 NHWC (or fp32 for the two head convolutions).
//   * warp roles: warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer, warps 2-5 = epilogue.
// The first layer (6 -> F channels, K = 54) consumes a [positions][64] bf16 matrix that
// encode_im2col_kernel builds directly from the packed 32-byte states (as_tensor,
// backgammon_logic.rs:198-252, fused with the tap gather), through the same kernel with one tap.
#include <cstdlib>
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/diee.h"
#include "net_launch.h"

namespace diee {

// ---------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void tma_load_4d(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
            smem_u32(dst)),
        "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
            smem_u32(dst)),
        "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap *map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t *dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem], bf16 inputs, fp32 accumulate
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// 32 lanes x 32 consecutive fp32 columns of this warp's TMEM quadrant
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}
// ---- CTA pair (cta_group::2): two CTAs of a cluster on the two SMs of a TPC execute ONE tcgen05.mma of M = 256 ----
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// both CTAs issue their own loads; the transaction bytes of both land on the LEADER's barrier (peer bit of the address cleared)
constexpr uint32_t PEER_BIT_MASK = 0xFEFFFFFFu;
__device__ __forceinline__ void tma_load_4d_2sm(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
            smem_u32(dst)),
        "l"(map), "r"(smem_u32(bar) & PEER_BIT_MASK), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
            smem_u32(dst)),
        "l"(map), "r"(smem_u32(bar) & PEER_BIT_MASK), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t *dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_bf16_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrives on the barrier at the same offset in BOTH CTAs of the pair once the MMAs issued so far have retired
__device__ __forceinline__ void umma_commit_2sm(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
                 "h"((uint16_t)3)
                 : "memory");
}

__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major operand tile with 128-byte rows in the 128B swizzle: 8-row groups are 1024 B apart.
// (cute::UMMA::SmemDescriptor: start>>4 | LBO(1)<<16 | SBO(64)<<32 | version 1<<46 | SWIZZLE_128B 2<<61)
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
    return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// cute::UMMA::InstrDescriptor, kind::f16: D=F32 (1<<4), A=B=BF16 (1<<7, 1<<10), K-major both, N>>3 at 17, M>>4 at 24
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ---------------------------------------------------------------- the convolution kernel
constexpr int CONV_THREADS = 192;

// A CTA computes NB boards x BN output channels.  NB = 16 (384 positions = 3 MMA M-tiles) x BN = 128 is the shape for
// big batches.  What a K-block costs a CTA does not depend on how many CTAs run (a layer takes the same 32 us at 16
// boards as at 1,024): about 0.2 us plus its MMAs at the rate their operands leave shared memory.  So a
// small batch is spread over as many CTAs as there are SMs, each with a smaller tile (net.cu picks the shape per
// launch), and the small tiles take KC = 2 chunks of 64 input channels per pipeline stage: half as many K-blocks.
// TWO = a CTA PAIR (cluster of 2, cta_group::2) computes 2 x NB boards x BN channels with M = 256 MMAs: each CTA stages its own
// boards and HALF of the weight tile, and both tensor cores read both halves.  What bounds the single-CTA kernel is the
// The hypothesis it tests: the single-CTA main loop is bound by the shared-memory port -- a 128x128x16 MMA reads 8 KB of
// operands per 68 tensor cycles (121 B/cycle) while TMA writes the next stage (79 B/cycle), 200 B/cycle wanted against 128
// B/cycle, i.e. the 64 % tensor-pipe activity ncu shows for a long K loop (profiles/r02_conv3x3_split3_ncu_full_summary.txt);
// the pair needs 6 KB per MMA and 56 KB per stage, 157 B/cycle.  MEASURED: same bits, 4-5 % SLOWER (net.cu use_cta_pairs),
// so that is not the limit; the form stays selectable (DIEE_CONV_2CTA) and tested, and is off by default.
template <int BN, int NB, int KC = 1, bool TWO = false>
struct ConvCfg {
    static constexpr int ROWS = NB * 24;
    static constexpr int MT = (ROWS + 127) / 128;          // the last M-tile may be partly padding (rows >= ROWS: never stored)
    static constexpr int A_BYTES = ROWS * 128;             // one chunk (64 bf16 channels) of the activation tile
    static constexpr int B_ROWS = TWO ? BN / 2 : BN;      // weight rows this CTA stages
    static constexpr int B_BYTES = B_ROWS * 128;
    static constexpr int STAGE_BYTES = KC * (A_BYTES + B_BYTES);  // [A chunk 0 .. A chunk KC-1 | B chunk 0 .. B chunk KC-1]
    static constexpr int STAGES_RAW = (216 * 1024) / STAGE_BYTES;
#ifndef DIEE_CONV_MAX_STAGES
#define DIEE_CONV_MAX_STAGES 8
#endif
    static constexpr int STAGES = STAGES_RAW > DIEE_CONV_MAX_STAGES ? DIEE_CONV_MAX_STAGES : STAGES_RAW;
    static constexpr int SLACK = (ROWS % 128) ? 16 * 1024 : 0;  // a padded M-tile reads past its tile: keep that inside the allocation
    static constexpr bool CAN_FUSE = 2 * MT * BN <= 512;  // room for the split-precision mode's two accumulators
    static constexpr int ACC_COLS = CAN_FUSE ? 2 * MT * BN : MT * BN;
    static constexpr int TMEM_COLS = (ACC_COLS <= 32) ? 32 : (ACC_COLS <= 64) ? 64 : (ACC_COLS <= 128) ? 128 : (ACC_COLS <= 256) ? 256 : 512;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/ + BN * 4 /*bias*/ + SLACK;
    static_assert(STAGE_BYTES % 1024 == 0 && A_BYTES % 1024 == 0 && B_BYTES % 1024 == 0, "operand tiles must stay 1024-byte aligned (128B swizzle atoms)");
    static_assert(STAGES >= 2, "pipeline depth");
    static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
    static_assert(MT * BN <= 512, "TMEM budget");
};

// out_mode: 0 = bf16 NHWC [rows][c_out_total], 1 = fp32 NHWC, 2 = fp32 NHWC through the split-precision epilogue (ConvEpi)
//
// Split precision (DIEE_NET_SPLIT3, the tensor-core mode inside the fp32 tolerance): an fp32 operand is carried as
// three bf16 planes.  The K loop runs over (pair, tap, chunk): `pairs` packs up to eight plane pairs (i, j) as 2+2 bits
// each, plane i of the activations starts a_plane channels further on, plane j of the weights b_plane K-columns further
// on.  Plain bf16 is the one pair (0, 0).
//
// What makes it accurate is WHICH planes (net.cu, split_planes_kernel): the tensor core's fp32 accumulation truncates
// on every MMA, a systematic shrink of ~0.5 ulp per instruction that 144 accumulations of full-size products turn into
// ~3e-6 per layer (measured round 1: 2.2e-4 through 39 layers).  So plane 0 is a 7-bit signed INTEGER digit: every
// board (and every output channel of the weights) is scaled by a power of two so that |x| <= 64 units, s0 = rint(x),
// and planes 1, 2 are the bf16 roundings of what is left (x = s0 + r1 + r2 to 2^-24 of the board's largest value).
// The products s0_x * s0_w are integers <= 4096 and a sum of 9 * C_in <= 2304 of them stays below 2^24: every partial
// sum is exactly representable, so the accumulation of the big pair (0, 0) is EXACT whatever the hardware rounds.
// It runs as its own launch; the five small pairs (2^-7 and below) run before it into an fp32 scratch (their
// truncation costs 2^-7 of the above), and the epilogue of the (0, 0) launch adds the two, applies the two power-of-two
// scales, bias, the fp32 residual and ReLU, stores fp32 and records every board's maximum for the next layer's planes.
struct ConvEpi {
    const float *addend;        // fp32 [rows][c_out_total]: the small pairs' sum, same units as the accumulator (nullable)
    const float *row_scale;     // [board]: units of this layer's activation planes (nullable = 1)
    const float *col_scale;     // [c_out_total]: units of the weight planes per output channel (nullable = 1)
    const float *residual_f32;  // fp32 [rows][c_out_total] (nullable)
    unsigned int *board_max;    // [board]: max of the (post-ReLU, >= 0) outputs as float bits, atomicMax (nullable)
    int fused;                  // 1: ONE launch runs all six pairs -- the five small ones into a first accumulator, the exact pair
                                // (the last one) into a second -- and the epilogue adds the two in registers instead of reading
                                // `addend`: no second launch, no scratch round trip.  Needs 2 x MT x BN TMEM columns (small tiles).
};
template <int BN, int NB, int KC, bool TWO = false>
__global__ void __launch_bounds__(CONV_THREADS, 1)
conv3x3_tc_kernel(const __grid_constant__ CUtensorMap tmapA, const __grid_constant__ CUtensorMap tmapB, int n_boards,
                  int ntaps, int chunks, const float *__restrict__ bias, const __nv_bfloat16 *__restrict__ residual,
                  void *__restrict__ out, int out_mode, int c_out_total, int relu, int npairs, uint32_t pairs, int a_plane,
                  int b_plane, ConvEpi epi) {
    using Cfg = ConvCfg<BN, NB, KC, TWO>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t *tail = smem + Cfg::STAGES * Cfg::STAGE_BYTES;
    uint64_t *full_bar = reinterpret_cast<uint64_t *>(tail);
    uint64_t *empty_bar = full_bar + Cfg::STAGES;
    uint64_t *tmem_full_bar = empty_bar + Cfg::STAGES;
    uint32_t *tmem_ptr_smem = reinterpret_cast<uint32_t *>(tmem_full_bar + 1);
    float *bias_smem = reinterpret_cast<float *>(tail + 256);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int board0 = blockIdx.x * NB;  // (TWO: the cluster is two consecutive blockIdx.x = two consecutive board tiles)
    const int n0 = blockIdx.y * BN;
    const uint32_t cta_rank = TWO ? cluster_ctarank() : 0u;
    const bool leader = cta_rank == 0;   // issues the MMAs of the pair and owns the `full` barriers
    const int per_pair = ntaps * chunks;
    const int num_kb = per_pair * npairs / KC;  // pipeline steps: KC consecutive chunks of one tap each (chunks % KC == 0)

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmapA);
        prefetch_tmap(&tmapB);
        for (int s = 0; s < Cfg::STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        mbar_init(tmem_full_bar, 1);
        fence_barrier_init();
    }
    if (warp == 1) { if (TWO) tmem_alloc_2sm(tmem_ptr_smem, Cfg::TMEM_COLS); else tmem_alloc(tmem_ptr_smem, Cfg::TMEM_COLS); }
    for (int i = threadIdx.x; i < BN; i += CONV_THREADS) bias_smem[i] = bias ? bias[n0 + i] : 0.f;
    tc_fence_before();
    if (TWO) cluster_sync_all();  // the peer's barriers are initialised before anything arrives on them
    else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    // Programmatic dependent launch: everything above (barrier init, TMEM allocation, descriptor prefetch, the bias tile --
    // constants of the net) may run while the previous layer's grid is still draining; its OUTPUT (this layer's
    // activations, residual, scratch) is only touched after this wait.  Dependents may be scheduled as soon as this CTA is here.
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % Cfg::STAGES;
                const uint32_t ph = (uint32_t)(kb / Cfg::STAGES) & 1u;
                mbar_wait(&empty_bar[s], ph ^ 1u);
                uint8_t *sa = smem + s * Cfg::STAGE_BYTES;
                uint8_t *sb = sa + KC * Cfg::A_BYTES;
                // the pair's two loads of a stage complete on the leader's barrier
                if (leader) mbar_expect_tx(&full_bar[s], (uint32_t)Cfg::STAGE_BYTES * (TWO ? 2u : 1u));
                const int k0 = kb * KC;
                const int pr = k0 / per_pair, kin = k0 - pr * per_pair;
                const int pi = (int)((pairs >> (4 * pr)) & 3u), pj = (int)((pairs >> (4 * pr + 2)) & 3u);
                const int tap = kin / chunks, chunk = kin - tap * chunks;
                const int kh = ntaps == 9 ? tap / 3 : 1, kw = ntaps == 9 ? tap - (tap / 3) * 3 : 1;
#pragma unroll
                for (int h = 0; h < KC; ++h) {
                    if (TWO) {
                        tma_load_4d_2sm(sa + h * Cfg::A_BYTES, &tmapA, &full_bar[s], pi * a_plane + (chunk + h) * 64, kw - 1, kh - 1, board0);
                        tma_load_2d_2sm(sb + h * Cfg::B_BYTES, &tmapB, &full_bar[s], pj * b_plane + (kin + h) * 64, n0 + (int)cta_rank * Cfg::B_ROWS);
                    } else {
                        tma_load_4d(sa + h * Cfg::A_BYTES, &tmapA, &full_bar[s], pi * a_plane + (chunk + h) * 64, kw - 1, kh - 1, board0);
                        tma_load_2d(sb + h * Cfg::B_BYTES, &tmapB, &full_bar[s], pj * b_plane + (kin + h) * 64, n0);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (one elected thread; of the pair: the leader's) =====
        constexpr uint32_t idesc = umma_idesc_bf16(TWO ? 256 : 128, BN);
        for (int kb = 0; leader && kb < num_kb; ++kb) {
            const int s = kb % Cfg::STAGES;
            const uint32_t ph = (uint32_t)(kb / Cfg::STAGES) & 1u;
            mbar_wait(&full_bar[s], ph);
            tc_fence_after();
            if (lane == 0) {
                const uint32_t sa = smem_u32(smem + s * Cfg::STAGE_BYTES);
                const uint32_t sb = sa + KC * Cfg::A_BYTES;
                // fused split precision: the last pair (the exact integer one) accumulates in its own TMEM region
                const int kb_big = (npairs - 1) * per_pair / KC;
                const bool big = epi.fused && kb >= kb_big;
                const uint32_t acc0 = tmem_base + (big ? (uint32_t)(Cfg::MT * BN) : 0u);
                const bool fresh = kb == 0 || (epi.fused && kb == kb_big);  // first K-block of an accumulator
#pragma unroll
                for (int h = 0; h < KC; ++h) {
#pragma unroll
                    for (int mt = 0; mt < Cfg::MT; ++mt) {
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk) {  // 4 x (K = 16 bf16 = 32 bytes) inside the 128-byte swizzled row
                            const uint64_t ad = umma_desc_sw128(sa + h * Cfg::A_BYTES + mt * (128 * 128) + kk * 32);
                            const uint64_t bd = umma_desc_sw128(sb + h * Cfg::B_BYTES + kk * 32);
                            const uint32_t accumulate = (fresh && h == 0 && kk == 0) ? 0u : 1u;
                            if (TWO) umma_bf16_2sm(acc0 + (uint32_t)(mt * BN), ad, bd, idesc, accumulate);
                            else umma_bf16(acc0 + (uint32_t)(mt * BN), ad, bd, idesc, accumulate);
                        }
                    }
                }
                if (TWO) {
                    umma_commit_2sm(&empty_bar[s]);                       // frees the stage in BOTH CTAs
                    if (kb == num_kb - 1) umma_commit_2sm(tmem_full_bar);  // both epilogues may read their accumulators
                } else {
                    umma_commit(&empty_bar[s]);                       // frees the smem stage when these MMAs retire
                    if (kb == num_kb - 1) umma_commit(tmem_full_bar);  // accumulators complete
                }
            }
            __syncwarp();
        }
    } else {
        // ===== epilogue: TMEM -> registers -> bias (+residual) -> ReLU -> HBM =====
        mbar_wait(tmem_full_bar, 0);
        tc_fence_after();
        const int q = warp & 3;  // TMEM lane quadrant this warp may read
        const long long total_rows = (long long)n_boards * 24;
#pragma unroll 1
        for (int mt = 0; mt < Cfg::MT; ++mt) {
            const int trow = mt * 128 + q * 32 + lane;
            const long long grow = (long long)board0 * 24 + trow;
            const bool valid = trow < Cfg::ROWS && grow < total_rows;
#pragma unroll 1
            for (int c0 = 0; c0 < BN; c0 += 32) {
                uint32_t v[32];
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(mt * BN + c0);
                const int ncol = (BN - c0) >= 32 ? 32 : 16;
                if (epi.fused) {  // v = exact pair's accumulator + the small pairs' (fp32 add, as the two-launch form does from its scratch)
                    uint32_t sm[32];
                    if (ncol == 32) { tmem_ld32(taddr, sm); tmem_ld32(taddr + (uint32_t)(Cfg::MT * BN), v); }
                    else { tmem_ld16(taddr, sm); tmem_ld16(taddr + (uint32_t)(Cfg::MT * BN), v); }
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(sm[j]));
                } else {
                    if (ncol == 32) tmem_ld32(taddr, v); else tmem_ld16(taddr, v);
                    tmem_ld_wait();
                }
                if (valid && out_mode == 2) {
                    // split-precision epilogue: (acc + small pairs) * 2^(e_board + e_channel) + bias (+ residual), ReLU, fp32
                    const size_t off = (size_t)grow * c_out_total + n0 + c0;
                    const float rs = epi.row_scale ? epi.row_scale[grow / 24] : 1.f;
                    float f[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) f[j] = j < ncol ? __uint_as_float(v[j]) : 0.f;
                    if (epi.addend) {
                        const float4 *ap = reinterpret_cast<const float4 *>(epi.addend + off);
#pragma unroll
                        for (int g4 = 0; g4 < 8; ++g4)
                            if (g4 * 4 < ncol) {
                                const float4 a = ap[g4];
                                f[g4 * 4] += a.x; f[g4 * 4 + 1] += a.y; f[g4 * 4 + 2] += a.z; f[g4 * 4 + 3] += a.w;
                            }
                    }
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (j < ncol) f[j] = f[j] * (rs * (epi.col_scale ? epi.col_scale[n0 + c0 + j] : 1.f)) + bias_smem[c0 + j];
                    if (epi.residual_f32) {
                        const float4 *rp = reinterpret_cast<const float4 *>(epi.residual_f32 + off);
#pragma unroll
                        for (int g4 = 0; g4 < 8; ++g4)
                            if (g4 * 4 < ncol) {
                                const float4 a = rp[g4];
                                f[g4 * 4] += a.x; f[g4 * 4 + 1] += a.y; f[g4 * 4 + 2] += a.z; f[g4 * 4 + 3] += a.w;
                            }
                    }
                    float mx = 0.f;
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        if (relu) f[j] = fmaxf(f[j], 0.f);
                        mx = fmaxf(mx, fabsf(f[j]));
                    }
                    float4 *op = reinterpret_cast<float4 *>(reinterpret_cast<float *>(out) + off);
#pragma unroll
                    for (int g4 = 0; g4 < 8; ++g4)
                        if (g4 * 4 < ncol) op[g4] = make_float4(f[g4 * 4], f[g4 * 4 + 1], f[g4 * 4 + 2], f[g4 * 4 + 3]);
                    if (epi.board_max) atomicMax(epi.board_max + grow / 24, __float_as_uint(mx));  // mx >= 0: bit order = value order
                } else if (valid) {
                    float f[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) f[j] = j < ncol ? __uint_as_float(v[j]) + bias_smem[c0 + j] : 0.f;
                    const size_t off = (size_t)grow * c_out_total + n0 + c0;
                    if (residual) {  // the skip connection (bf16 NHWC)
                        const uint4 *rp = reinterpret_cast<const uint4 *>(residual + off);
#pragma unroll
                        for (int g4 = 0; g4 < 4; ++g4) {
                            const uint4 r = rp[g4];
                            const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
                            for (int h = 0; h < 4; ++h) {
                                const __nv_bfloat162 b2 = *reinterpret_cast<const __nv_bfloat162 *>(&w[h]);
                                f[g4 * 8 + h * 2] += __bfloat162float(b2.x);
                                f[g4 * 8 + h * 2 + 1] += __bfloat162float(b2.y);
                            }
                        }
                    }
                    if (relu) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.f);
                    }
                    if (out_mode == 0) {
                        uint4 *op = reinterpret_cast<uint4 *>(reinterpret_cast<__nv_bfloat16 *>(out) + off);
#pragma unroll
                        for (int g4 = 0; g4 < 4; ++g4) {
                            uint32_t w[4];
#pragma unroll
                            for (int h = 0; h < 4; ++h) {
                                const __nv_bfloat162 b2 = __floats2bfloat162_rn(f[g4 * 8 + h * 2], f[g4 * 8 + h * 2 + 1]);
                                w[h] = *reinterpret_cast<const uint32_t *>(&b2);
                            }
                            if (g4 * 8 < ncol) op[g4] = make_uint4(w[0], w[1], w[2], w[3]);
                        }
                    } else {
                        float4 *op = reinterpret_cast<float4 *>(reinterpret_cast<float *>(out) + off);
#pragma unroll
                        for (int g4 = 0; g4 < 8; ++g4)
                            if (g4 * 4 < ncol) op[g4] = make_float4(f[g4 * 4], f[g4 * 4 + 1], f[g4 * 4 + 2], f[g4 * 4 + 3]);
                    }
                }
            }
        }
        tc_fence_before();
    }
    if (TWO) cluster_sync_all();  // nobody leaves (or frees TMEM) while the peer's tensor core may still read this CTA's tiles
    else __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        if (TWO) tmem_dealloc_2sm(tmem_base, Cfg::TMEM_COLS); else tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
    }
}

// ---------------------------------------------------------------- first-layer operand: as_tensor + tap gather
// out[(board*24 + pos)][k], k = tap*6 + channel (k < 54), zero-padded to 64; bf16 (all values are small integers)
__global__ void encode_im2col_kernel(const diee_bg_state *__restrict__ states, int n, __nv_bfloat16 *__restrict__ out) {
    const long long total = (long long)n * 24 * 64;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const int k = (int)(idx & 63);
        const long long rp = idx >> 6;
        const int pos = (int)(rp % 24);
        const long long b = rp / 24;
        float val = 0.f;
        if (k < 54) {
            const int tap = k / 6, c = k - tap * 6;
            const int hh = pos / 6 + tap / 3 - 1, ww = pos % 6 + tap % 3 - 1;
            if (hh >= 0 && hh < 4 && ww >= 0 && ww < 6) {
                const diee_bg_state &s = states[b];
                const int pt = hh * 6 + ww, half = pt < 12 ? 0 : 1;
                switch (c) {  // as_tensor channel order, backgammon_logic.rs:240-250
                    case 0: val = (float)s.pts[pt]; break;
                    case 1: val = (float)s.player; break;
                    case 2: val = (float)s.bar[half]; break;
                    case 3: val = (float)s.off[half]; break;
                    case 4: val = (float)s.roll[half]; break;
                    default: val = s.second ? 1.f : 0.f; break;
                }
            }
        }
        out[idx] = __float2bfloat16(val);
    }
}

// ---------------------------------------------------------------- split-precision operand planes
// One CTA per board: y fp32 [24][C] -> planes bf16 [24][3][C] = (s0, r1, r2) in units of 2^(e - sb), 2^e >= the board's
// largest |y| (board_max, as float bits; reset to 0 here for the next layer); scale_out[board] = 2^(e - sb).
// s0 = rint(y / unit) is an integer with |s0| <= 2^sb; r1, r2 = bf16 roundings of the remainder.
__global__ void __launch_bounds__(256)
split_planes_kernel(const float *__restrict__ y, unsigned int *__restrict__ board_max, int C, int sb, __nv_bfloat16 *__restrict__ planes,
                    float *__restrict__ scale_out) {
    const int b = blockIdx.x;
    const float mx = __uint_as_float(board_max[b]);
    int e = 0;
    if (mx > 0.f) frexpf(mx, &e);  // mx = m * 2^e, m in [0.5, 1): 2^e > mx
    const float unit = ldexpf(1.f, e - sb), inv = ldexpf(1.f, sb - e);  // powers of two: exact
    const float *yb = y + (size_t)b * 24 * C;
    __nv_bfloat16 *pb = planes + (size_t)b * 24 * 3 * C;
    for (int i = threadIdx.x; i < 24 * C; i += 256) {
        const int pos = i / C, c = i - pos * C;
        const float u = yb[i] * inv;
        const float s0 = rintf(u);
        const float d1 = u - s0;  // exact: |u| <= 2^sb, the difference of two nearby floats
        const __nv_bfloat16 r1 = __float2bfloat16(d1);
        const __nv_bfloat16 r2 = __float2bfloat16(d1 - __bfloat162float(r1));
        __nv_bfloat16 *o = pb + (size_t)pos * 3 * C + c;
        o[0] = __float2bfloat16(s0);  // |s0| <= 64: exact in bf16
        o[C] = r1;
        o[2 * C] = r2;
    }
    __syncthreads();
    if (threadIdx.x == 0) { scale_out[b] = unit; board_max[b] = 0u; }
}

cudaError_t launch_split_planes(cudaStream_t st, const float *y, unsigned int *board_max, int n, int C, int sb, void *planes, float *scale_out) {
    if (n <= 0) return cudaSuccess;
    split_planes_kernel<<<(unsigned)n, 256, 0, st>>>(y, board_max, C, sb, static_cast<__nv_bfloat16 *>(planes), scale_out);
    return cudaGetLastError();
}

// ---------------------------------------------------------------- heads
// Policy Linear(768 -> 1352) (nnet.rs:79-84) as one more tcgen05 GEMM: A = bf16 policy-conv features
// [board][768] (K index = pos*32 + c, the NHWC order the conv epilogue writes), B = the Linear weight
// re-ordered to that K order and zero-padded to 1408 rows, both K-major in the 128B swizzle.
// One CTA = 128 boards x 128 logits, 12 K-blocks of 64; fp32 logits + bias go straight to policy_out.
constexpr int FC_STAGES = 4;
constexpr int FC_STAGE_BYTES = 2 * 128 * 128;
constexpr int FC_SMEM_BYTES = FC_STAGES * FC_STAGE_BYTES + 1024 + 256 + 128 * 4;
constexpr int FC_N_PAD = 1408;  // 11 tiles of 128 >= 1352

__global__ void __launch_bounds__(CONV_THREADS, 1)
fc_tc_kernel(const __grid_constant__ CUtensorMap tmapA, const __grid_constant__ CUtensorMap tmapB, int n_boards,
             const float *__restrict__ bias, float *__restrict__ out) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t *tail = smem + FC_STAGES * FC_STAGE_BYTES;
    uint64_t *full_bar = reinterpret_cast<uint64_t *>(tail);
    uint64_t *empty_bar = full_bar + FC_STAGES;
    uint64_t *tmem_full_bar = empty_bar + FC_STAGES;
    uint32_t *tmem_ptr_smem = reinterpret_cast<uint32_t *>(tmem_full_bar + 1);
    float *bias_smem = reinterpret_cast<float *>(tail + 256);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int board0 = blockIdx.x * 128, n0 = blockIdx.y * 128;
    constexpr int num_kb = 768 / 64;
    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmapA);
        prefetch_tmap(&tmapB);
        for (int s = 0; s < FC_STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        mbar_init(tmem_full_bar, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_ptr_smem, 128);
    for (int i = threadIdx.x; i < 128; i += CONV_THREADS) bias_smem[i] = (n0 + i) < DIEE_ACTION_SPACE ? bias[n0 + i] : 0.f;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    if (warp == 0) {
        if (lane == 0) {
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % FC_STAGES;
                mbar_wait(&empty_bar[s], ((uint32_t)(kb / FC_STAGES) & 1u) ^ 1u);
                uint8_t *sa = smem + s * FC_STAGE_BYTES;
                mbar_expect_tx(&full_bar[s], (uint32_t)FC_STAGE_BYTES);
                tma_load_2d(sa, &tmapA, &full_bar[s], kb * 64, board0);
                tma_load_2d(sa + 128 * 128, &tmapB, &full_bar[s], kb * 64, n0);
            }
        }
    } else if (warp == 1) {
        constexpr uint32_t idesc = umma_idesc_bf16(128, 128);
        for (int kb = 0; kb < num_kb; ++kb) {
            const int s = kb % FC_STAGES;
            mbar_wait(&full_bar[s], (uint32_t)(kb / FC_STAGES) & 1u);
            tc_fence_after();
            if (lane == 0) {
                const uint32_t sa = smem_u32(smem + s * FC_STAGE_BYTES), sb = sa + 128 * 128;
#pragma unroll
                for (int kk = 0; kk < 4; ++kk)
                    umma_bf16(tmem_base, umma_desc_sw128(sa + kk * 32), umma_desc_sw128(sb + kk * 32), idesc, (kb | kk) != 0 ? 1u : 0u);
                umma_commit(&empty_bar[s]);
                if (kb == num_kb - 1) umma_commit(tmem_full_bar);
            }
            __syncwarp();
        }
    } else {
        mbar_wait(tmem_full_bar, 0);
        tc_fence_after();
        const int q = warp & 3;
        const int row = board0 + q * 32 + lane;
        const bool valid = row < n_boards;
#pragma unroll 1
        for (int c0 = 0; c0 < 128; c0 += 32) {
            uint32_t v[32];
            tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
            tmem_ld_wait();
            if (valid) {
                float4 *op = reinterpret_cast<float4 *>(out + (size_t)row * DIEE_ACTION_SPACE + n0 + c0);
#pragma unroll
                for (int g4 = 0; g4 < 8; ++g4)
                    if (n0 + c0 + g4 * 4 + 3 < DIEE_ACTION_SPACE)
                        op[g4] = make_float4(__uint_as_float(v[g4 * 4]) + bias_smem[c0 + g4 * 4],
                                             __uint_as_float(v[g4 * 4 + 1]) + bias_smem[c0 + g4 * 4 + 1],
                                             __uint_as_float(v[g4 * 4 + 2]) + bias_smem[c0 + g4 * 4 + 2],
                                             __uint_as_float(v[g4 * 4 + 3]) + bias_smem[c0 + g4 * 4 + 3]);
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 128);
    }
}

// softmax(1) over the 1352 logits in place (nnet.rs:127) and the value head Linear(72 -> 1) + tanh
// (nnet.rs:87-98): one warp per board.  vfeat: fp32 [n*24][16] (3 channels used); wv: fp32 [24*16].
__global__ void __launch_bounds__(256)
softmax_value_kernel(const float *__restrict__ vfeat, const float *__restrict__ wv, float bv, int n, float *__restrict__ policy,
                     float *__restrict__ value_out) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.x * 8 + warp;
    if (b >= n) return;
    float *row = policy + (size_t)b * DIEE_ACTION_SPACE;
    float x[43];
    float mx = -INFINITY;
#pragma unroll
    for (int i = 0; i < 43; ++i) {
        const int j = lane + 32 * i;
        x[i] = j < DIEE_ACTION_SPACE ? row[j] : -INFINITY;
        mx = fmaxf(mx, x[i]);
    }
    for (int d = 16; d; d >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xFFFFFFFFu, mx, d));
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < 43; ++i) { x[i] = expf(x[i] - mx); sum += x[i]; }
    for (int d = 16; d; d >>= 1) sum += __shfl_xor_sync(0xFFFFFFFFu, sum, d);
    const float inv = 1.f / sum;
#pragma unroll
    for (int i = 0; i < 43; ++i) {
        const int j = lane + 32 * i;
        if (j < DIEE_ACTION_SPACE) row[j] = x[i] * inv;
    }
    float acc = 0.f;
    const float *vf = vfeat + (size_t)b * 24 * 16;
    for (int f = lane; f < 24 * 16; f += 32) acc = fmaf(wv[f], vf[f], acc);
    for (int d = 16; d; d >>= 1) acc += __shfl_xor_sync(0xFFFFFFFFu, acc, d);
    if (lane == 0) value_out[b] = tanhf(acc + bv);
}

// ---------------------------------------------------------------- fp32 reference-arithmetic mode (DIEE_NET_FP32)
// The reference computes in fp32 (lib.rs:20, tch conv2d / batch_norm / linear).  This mode does the same on the
// CUDA cores: IEEE single FMAs with round-to-nearest, one accumulator per output, so it differs from the
// reference only by summation order.  It is the PARITY mode of the net (speed is not its purpose; the tensor
// cores accumulate with truncation, which costs ~1e-5 per layer -- see DIEE_NET_SPLIT3).
//   x: fp32 NHWC [n][24][c_in] (or, first layer, the packed states: as_tensor is applied on the fly, c_in = 6)
//   w: fp32 [9][c_in][c_out] with eval-mode BatchNorm folded in; out: fp32 [n][24][out_stride]
// One CTA = one board x 64 output channels; thread = (channel, half of the 24 positions).
__global__ void __launch_bounds__(128)
conv3x3_f32_kernel(const float *__restrict__ x, const diee_bg_state *__restrict__ states, int c_in, const float *__restrict__ w,
                   const float *__restrict__ bias, const float *__restrict__ residual, float *__restrict__ out, int c_out,
                   int out_stride, int relu) {
    extern __shared__ float xs[];  // [24][c_in]
    const int b = blockIdx.x;
    if (states) {
        const diee_bg_state &s = states[b];
        for (int i = threadIdx.x; i < 24 * 6; i += 128) {
            const int pt = i / 6, c = i - pt * 6, half = pt < 12 ? 0 : 1;
            float val;
            switch (c) {  // as_tensor channel order, backgammon_logic.rs:240-250
                case 0: val = (float)s.pts[pt]; break;
                case 1: val = (float)s.player; break;
                case 2: val = (float)s.bar[half]; break;
                case 3: val = (float)s.off[half]; break;
                case 4: val = (float)s.roll[half]; break;
                default: val = s.second ? 1.f : 0.f; break;
            }
            xs[i] = val;
        }
    } else {
        for (int i = threadIdx.x; i < 24 * c_in; i += 128) xs[i] = x[(size_t)b * 24 * c_in + i];
    }
    __syncthreads();
    const int co = blockIdx.y * 64 + (threadIdx.x & 63);
    const int p0 = (threadIdx.x >> 6) * 12;
    if (co >= c_out) return;
    float acc[12];
#pragma unroll
    for (int p = 0; p < 12; ++p) acc[p] = 0.f;
    for (int tap = 0; tap < 9; ++tap) {
        const int dh = tap / 3 - 1, dw = tap % 3 - 1;
        int src[12];
#pragma unroll
        for (int p = 0; p < 12; ++p) {
            const int pos = p0 + p, hh = pos / 6 + dh, ww = pos % 6 + dw;
            src[p] = (hh >= 0 && hh < 4 && ww >= 0 && ww < 6) ? (hh * 6 + ww) * c_in : -1;
        }
        const float *wt = w + (size_t)tap * c_in * c_out + co;
        for (int ci = 0; ci < c_in; ++ci) {
            const float wv = wt[(size_t)ci * c_out];
#pragma unroll
            for (int p = 0; p < 12; ++p)
                if (src[p] >= 0) acc[p] = fmaf(xs[src[p] + ci], wv, acc[p]);
        }
    }
    const float bv = bias[co];
#pragma unroll
    for (int p = 0; p < 12; ++p) {
        const size_t row = (size_t)b * 24 + p0 + p;
        float v = acc[p] + bv;
        if (residual) v += residual[row * out_stride + co];
        if (relu) v = fmaxf(v, 0.f);
        out[row * out_stride + co] = v;
    }
}

// Policy Linear(768 -> 1352) in plain fp32 for the split-precision (fp32-parity) mode: feat fp32 [n][768]
// (K index = pos*32 + c), wT fp32 [768][1352].  One thread per logit, sequential fp32 FMA over K.
__global__ void __launch_bounds__(256)
fc_f32_kernel(const float *__restrict__ feat, const float *__restrict__ wT, const float *__restrict__ bias, int n,
              float *__restrict__ out) {
    __shared__ float f[768];
    const int b = blockIdx.x;
    for (int i = threadIdx.x; i < 768; i += 256) f[i] = feat[(size_t)b * 768 + i];
    __syncthreads();
    for (int j = blockIdx.y * 256 + threadIdx.x; j < DIEE_ACTION_SPACE; j += gridDim.y * 256) {
        float acc = 0.f;
#pragma unroll 8
        for (int k = 0; k < 768; ++k) acc = fmaf(f[k], wT[(size_t)k * DIEE_ACTION_SPACE + j], acc);
        out[(size_t)b * DIEE_ACTION_SPACE + j] = acc + bias[j];
    }
}

// ---------------------------------------------------------------- launchers
// DIEE_CONV_PDL=0 launches the convolutions without programmatic dependent launch (experiments)
static bool pdl_enabled() {
    static const bool on = !(getenv("DIEE_CONV_PDL") && atoi(getenv("DIEE_CONV_PDL")) == 0);
    return on;
}

// cudaFuncSetAttribute is per DEVICE (a process may hold contexts on several): once per device and kernel
template <class F>
static cudaError_t smem_opt_in(F *fn, int bytes, unsigned long long &done) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 64 && ((done >> dev) & 1ull)) return cudaSuccess;
    if ((e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes)) != cudaSuccess) return e;
    if (dev < 64) done |= 1ull << dev;
    return cudaSuccess;
}

template <int BN, int NB, int KC>
static cudaError_t launch_conv_bn(cudaStream_t st, const CUtensorMap &ta, const CUtensorMap &tb, int n_boards, int ntaps,
                                  int chunks, const float *bias, const __nv_bfloat16 *residual, void *out, int out_mode,
                                  int c_out_total, int relu, int npairs, uint32_t pairs, int a_plane, int b_plane, const ConvEpi &epi) {
    static unsigned long long opted = 0;
    if (cudaError_t e = smem_opt_in(conv3x3_tc_kernel<BN, NB, KC>, ConvCfg<BN, NB, KC>::SMEM_BYTES, opted); e != cudaSuccess) return e;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)((n_boards + NB - 1) / NB), (unsigned)(c_out_total / BN));
    cfg.blockDim = dim3(CONV_THREADS);
    cfg.dynamicSmemBytes = ConvCfg<BN, NB, KC>::SMEM_BYTES;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;  // see griddepcontrol.wait in the kernel
    attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, conv3x3_tc_kernel<BN, NB, KC>, ta, tb, n_boards, ntaps, chunks, bias, residual, out, out_mode,
                              c_out_total, relu, npairs, pairs, a_plane, b_plane, epi);
}

// the CTA-pair form of the 16-board x 128-channel tile: `tb` encoded with a box of 64 rows (each CTA stages half of the tile)
cudaError_t launch_conv_pair(cudaStream_t st, const CUtensorMap &ta, const CUtensorMap &tb, int n_boards, int ntaps, int chunks,
                             const float *bias, const void *residual_v, void *out, int out_mode, int c_out_total, int relu,
                             int npairs, uint32_t pairs, int a_plane, int b_plane, const SplitEpilogue *sp) {
    using Cfg = ConvCfg<128, 16, 1, true>;
    const __nv_bfloat16 *residual = static_cast<const __nv_bfloat16 *>(residual_v);
    ConvEpi epi{};
    if (sp) epi = ConvEpi{sp->addend, sp->row_scale, sp->col_scale, sp->residual_f32, sp->board_max, sp->fused};
    static unsigned long long opted = 0;
    if (cudaError_t e = smem_opt_in(conv3x3_tc_kernel<128, 16, 1, true>, Cfg::SMEM_BYTES, opted); e != cudaSuccess) return e;
    cudaLaunchConfig_t cfg{};
    const unsigned tiles = (unsigned)((n_boards + 15) / 16);
    cfg.gridDim = dim3((tiles + 1u) & ~1u, (unsigned)(c_out_total / 128));  // a tile past the batch loads zeros and stores nothing
    cfg.blockDim = dim3(CONV_THREADS);
    cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, conv3x3_tc_kernel<128, 16, 1, true>, ta, tb, n_boards, ntaps, chunks, bias, residual, out, out_mode,
                              c_out_total, relu, npairs, pairs, a_plane, b_plane, epi);
}

// `ta` must have been encoded with a box of nb boards, `tb` with a box of bn rows
cudaError_t launch_conv_tile(cudaStream_t st, int bn, int nb, const CUtensorMap &ta, const CUtensorMap &tb, int n_boards, int ntaps,
                             int chunks, const float *bias, const void *residual, void *out, int out_mode, int c_out_total, int relu,
                             int npairs, uint32_t pairs, int a_plane, int b_plane, const SplitEpilogue *sp) {
    const __nv_bfloat16 *res = static_cast<const __nv_bfloat16 *>(residual);
    ConvEpi epi{};
    if (sp) epi = ConvEpi{sp->addend, sp->row_scale, sp->col_scale, sp->residual_f32, sp->board_max, sp->fused};

    // several chunks per pipeline stage where the tile is small enough for >= 3 such stages and the layer's chunk count
    // divides (DIEE_CONV_KC=n caps it; 1 keeps one chunk per stage everywhere)
    static const bool one_chunk = getenv("DIEE_CONV_KC") && atoi(getenv("DIEE_CONV_KC")) == 1;
    const int max_kc = one_chunk ? 1 : (getenv("DIEE_CONV_KC") ? atoi(getenv("DIEE_CONV_KC")) : 4);
#define DIEE_CONV_CASE(BN_, NB_, KC_)                                                                                                        \
    if (bn == BN_ && nb == NB_ && KC_ <= max_kc && chunks % KC_ == 0)                                                                                    \
        return launch_conv_bn<BN_, NB_, KC_>(st, ta, tb, n_boards, ntaps, chunks, bias, res, out, out_mode, c_out_total, relu, npairs, pairs, \
                                             a_plane, b_plane, epi);
    DIEE_CONV_CASE(128, 16, 1)
    DIEE_CONV_CASE(32, 16, 1)
    DIEE_CONV_CASE(16, 16, 1)
    DIEE_CONV_CASE(128, 8, 2)
    DIEE_CONV_CASE(128, 8, 1)
    DIEE_CONV_CASE(64, 8, 2)
    DIEE_CONV_CASE(64, 8, 1)
    DIEE_CONV_CASE(32, 8, 2)
    DIEE_CONV_CASE(32, 8, 1)
    DIEE_CONV_CASE(32, 4, 4)
    DIEE_CONV_CASE(32, 4, 2)
    DIEE_CONV_CASE(32, 4, 1)
#undef DIEE_CONV_CASE
    return cudaErrorInvalidValue;
}

cudaError_t launch_conv(cudaStream_t st, int bn, const CUtensorMap &ta, const CUtensorMap &tb, int n_boards, int ntaps, int chunks,
                        const float *bias, const void *residual, void *out, int out_mode, int c_out_total, int relu,
                        int npairs, uint32_t pairs, int a_plane, int b_plane, const SplitEpilogue *sp) {
    return launch_conv_tile(st, bn, 16, ta, tb, n_boards, ntaps, chunks, bias, residual, out, out_mode, c_out_total, relu, npairs, pairs,
                            a_plane, b_plane, sp);
}

cudaError_t launch_encode_im2col(cudaStream_t st, const diee_bg_state *states, int n, void *out) {
    if (n <= 0) return cudaSuccess;
    long long total = (long long)n * 24 * 64;
    int blocks = (int)((total + 255) / 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    encode_im2col_kernel<<<blocks, 256, 0, st>>>(states, n, static_cast<__nv_bfloat16 *>(out));
    return cudaGetLastError();
}

cudaError_t launch_heads(cudaStream_t st, const CUtensorMap &ta, const CUtensorMap &tb, const float *bp, const float *vfeat,
                         const float *wv, float bv, int n, float *policy_out, float *value_out) {
    if (n <= 0) return cudaSuccess;
    static unsigned long long opted = 0;
    if (cudaError_t e0 = smem_opt_in(fc_tc_kernel, FC_SMEM_BYTES, opted); e0 != cudaSuccess) return e0;
    dim3 grid((n + 127) / 128, FC_N_PAD / 128);
    fc_tc_kernel<<<grid, CONV_THREADS, FC_SMEM_BYTES, st>>>(ta, tb, n, bp, policy_out);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    softmax_value_kernel<<<(n + 7) / 8, 256, 0, st>>>(vfeat, wv, bv, n, policy_out, value_out);
    return cudaGetLastError();
}

cudaError_t launch_conv_f32(cudaStream_t st, const float *x, const diee_bg_state *states, int n, int c_in, const float *w,
                            const float *bias, const float *residual, float *out, int c_out, int out_stride, int relu) {
    if (n <= 0) return cudaSuccess;
    conv3x3_f32_kernel<<<dim3((unsigned)n, (unsigned)((c_out + 63) / 64)), 128, (size_t)24 * c_in * sizeof(float), st>>>(
        x, states, c_in, w, bias, residual, out, c_out, out_stride, relu);
    return cudaGetLastError();
}

cudaError_t launch_heads_f32(cudaStream_t st, const float *pfeat, const float *wT, const float *bp, const float *vfeat,
                             const float *wv, float bv, int n, float *policy_out, float *value_out) {
    if (n <= 0) return cudaSuccess;
    fc_f32_kernel<<<dim3((unsigned)n, 6), 256, 0, st>>>(pfeat, wT, bp, n, policy_out);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    softmax_value_kernel<<<(n + 7) / 8, 256, 0, st>>>(vfeat, wv, bv, n, policy_out, value_out);
    return cudaGetLastError();
}

}  // namespace diee
