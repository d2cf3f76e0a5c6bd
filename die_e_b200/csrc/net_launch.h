// net_launch.h -- internal: launchers of net_kernels.cu
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/diee.h"

namespace diee {
// inputs of the split-precision epilogue (out_mode 2), see ConvEpi in net_kernels.cu
struct SplitEpilogue {
    const float *addend, *row_scale, *col_scale, *residual_f32;
    unsigned int *board_max;
    int fused;  // all six pairs in one launch, two TMEM accumulators (tiles with 2 x MT x BN <= 512 columns only)
};
// one convolution with an explicit CTA tile: nb boards x bn output channels (ta encoded with a box of nb boards, tb with bn rows)
cudaError_t launch_conv_tile(cudaStream_t st, int bn, int nb, const CUtensorMap &ta, const CUtensorMap &tb, int n_boards, int ntaps,
                             int chunks, const float *bias, const void *residual, void *out, int out_mode, int c_out_total, int relu,
                             int npairs = 1, uint32_t pairs = 0, int a_plane = 0, int b_plane = 0, const SplitEpilogue *sp = nullptr);
cudaError_t launch_conv(cudaStream_t st, int bn, const CUtensorMap &ta, const CUtensorMap &tb, int n_boards, int ntaps, int chunks,
                        const float *bias, const void *residual, void *out, int out_mode, int c_out_total, int relu,
                        int npairs = 1, uint32_t pairs = 0, int a_plane = 0, int b_plane = 0, const SplitEpilogue *sp = nullptr);
// the 16-board x 128-channel tile on a CTA PAIR (cta_group::2): ta with a box of 16 boards, tb with a box of 64 rows
cudaError_t launch_conv_pair(cudaStream_t st, const CUtensorMap &ta, const CUtensorMap &tb, int n_boards, int ntaps, int chunks,
                             const float *bias, const void *residual, void *out, int out_mode, int c_out_total, int relu,
                             int npairs = 1, uint32_t pairs = 0, int a_plane = 0, int b_plane = 0, const SplitEpilogue *sp = nullptr);
// fp32 activations -> the three operand planes of the split-precision mode (one CTA per board)
cudaError_t launch_split_planes(cudaStream_t st, const float *y, unsigned int *board_max, int n, int C, int sb, void *planes, float *scale_out);
cudaError_t launch_conv_f32(cudaStream_t st, const float *x, const diee_bg_state *states, int n, int c_in, const float *w,
                            const float *bias, const float *residual, float *out, int c_out, int out_stride, int relu);
cudaError_t launch_heads_f32(cudaStream_t st, const float *pfeat, const float *wT, const float *bp, const float *vfeat,
                             const float *wv, float bv, int n, float *policy_out, float *value_out);
cudaError_t launch_encode_im2col(cudaStream_t st, const diee_bg_state *states, int n, void *out);
cudaError_t launch_heads(cudaStream_t st, const CUtensorMap &ta, const CUtensorMap &tb, const float *bp, const float *vfeat,
                         const float *wv, float bv, int n, float *policy_out, float *value_out);
}  // namespace diee
