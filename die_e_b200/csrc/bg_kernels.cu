// bg_kernels.cu -- backgammon env kernels, one warp per game.
#include "bg_device.cuh"
#include "launchers.h"

namespace diee {

constexpr int WARPS_PER_CTA = 8;

// get_valid_moves for n states: moves_out[i][DIEE_MAX_MOVES], counts_out[i], ids_out (nullable)
__global__ void __launch_bounds__(WARPS_PER_CTA * 32)
bg_valid_moves_kernel(const diee_bg_state *__restrict__ states, int n, diee_move *__restrict__ moves_out,
                      int32_t *__restrict__ counts_out, uint16_t *__restrict__ ids_out) {
    __shared__ WarpSlab slabs[WARPS_PER_CTA];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int gidx = blockIdx.x * WARPS_PER_CTA + wib;
    if (gidx >= n) return;
    WarpSlab &slab = slabs[wib];
    BgWarp g;
    bg_load(g, states + gidx, lane);
    if (g.roll0 == 0 && g.roll1 == 0) {  // assert :404
        if (lane == 0) counts_out[gidx] = DIEE_ERR_NOT_ROLLED;
        return;
    }
    bool overflow = false;
    const int U = bg_movegen(g, slab, lane, overflow);
    uint32_t *mo = reinterpret_cast<uint32_t *>(moves_out) + (size_t)gidx * DIEE_MAX_MOVES;
    for (int i = lane; i < U; i += 32) {
        const uint32_t s = slab.raw[i];
        mo[i] = s;
        if (ids_out) ids_out[(size_t)gidx * DIEE_MAX_MOVES + i] = (uint16_t)bg_encode_move(g.roll0, g.roll1, s);
    }
    if (lane == 0) counts_out[gidx] = overflow ? DIEE_ERR_OVERFLOW : U;
}

// apply_move / skip_turn with injected next rolls
__global__ void __launch_bounds__(WARPS_PER_CTA * 32)
bg_apply_kernel(diee_bg_state *__restrict__ states, const diee_move *__restrict__ moves,
                const uint8_t *__restrict__ next_rolls, int n) {
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int gidx = blockIdx.x * WARPS_PER_CTA + wib;
    if (gidx >= n) return;
    BgWarp g;
    bg_load(g, states + gidx, lane);
    const uint32_t seq = reinterpret_cast<const uint32_t *>(moves)[gidx];
    bg_step(g, seq, next_rolls[2 * gidx], next_rolls[2 * gidx + 1], lane);
    bg_store(g, states + gidx, lane);
}

// C2 (whole random-vs-random games) lives in lane_kernels.cu: one lane per game.

// encode / decode one play per state: one thread per state
__global__ void bg_encode_moves_kernel(const diee_bg_state *__restrict__ states, const diee_move *__restrict__ moves,
                                       int n, uint16_t *__restrict__ ids_out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t seq = reinterpret_cast<const uint32_t *>(moves)[i];
    ids_out[i] = (uint16_t)bg_encode_move(states[i].roll[0], states[i].roll[1], seq);
}

__global__ void bg_decode_moves_kernel(const diee_bg_state *__restrict__ states, const uint16_t *__restrict__ ids,
                                       int n, diee_move *__restrict__ moves_out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    reinterpret_cast<uint32_t *>(moves_out)[i] =
        bg_decode_move(states[i].roll[0], states[i].roll[1], states[i].player, ids[i]);
}

// as_tensor :198-252 -> f32 [n,6,4,6]; 144 threads per state write coalesced rows
__global__ void bg_encode_states_kernel(const diee_bg_state *__restrict__ states, int n, float *__restrict__ out) {
    const int total = n * 144;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        const int i = idx / 144, r = idx - i * 144, c = r / 24, pt = r - c * 24;
        const diee_bg_state &s = states[i];
        const int half = pt < 12 ? 0 : 1;
        float val;
        switch (c) {
            case 0: val = (float)s.pts[pt]; break;
            case 1: val = (float)s.player; break;
            case 2: val = (float)s.bar[half]; break;
            case 3: val = (float)s.off[half]; break;
            case 4: val = (float)s.roll[half]; break;
            default: val = s.second ? 1.0f : 0.0f; break;
        }
        out[idx] = val;
    }
}

static inline int warp_grid(int n) { return (n + WARPS_PER_CTA - 1) / WARPS_PER_CTA; }

cudaError_t launch_bg_valid_moves(cudaStream_t st, const diee_bg_state *states, int n, diee_move *moves_out,
                                  int32_t *counts_out, uint16_t *ids_out) {
    if (n <= 0) return cudaSuccess;
    bg_valid_moves_kernel<<<warp_grid(n), WARPS_PER_CTA * 32, 0, st>>>(states, n, moves_out, counts_out, ids_out);
    return cudaGetLastError();
}
cudaError_t launch_bg_apply(cudaStream_t st, diee_bg_state *states, const diee_move *moves, const uint8_t *next_rolls, int n) {
    if (n <= 0) return cudaSuccess;
    bg_apply_kernel<<<warp_grid(n), WARPS_PER_CTA * 32, 0, st>>>(states, moves, next_rolls, n);
    return cudaGetLastError();
}
cudaError_t launch_bg_encode_moves(cudaStream_t st, const diee_bg_state *states, const diee_move *moves, int n, uint16_t *ids_out) {
    if (n <= 0) return cudaSuccess;
    bg_encode_moves_kernel<<<(n + 255) / 256, 256, 0, st>>>(states, moves, n, ids_out);
    return cudaGetLastError();
}
cudaError_t launch_bg_decode_moves(cudaStream_t st, const diee_bg_state *states, const uint16_t *ids, int n, diee_move *moves_out) {
    if (n <= 0) return cudaSuccess;
    bg_decode_moves_kernel<<<(n + 255) / 256, 256, 0, st>>>(states, ids, n, moves_out);
    return cudaGetLastError();
}
cudaError_t launch_bg_encode_states(cudaStream_t st, const diee_bg_state *states, int n, float *out) {
    if (n <= 0) return cudaSuccess;
    const int total = n * 144;
    int blocks = (total + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    bg_encode_states_kernel<<<blocks, 256, 0, st>>>(states, n, out);
    return cudaGetLastError();
}

}  // namespace diee
