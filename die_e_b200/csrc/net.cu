// net.cu -- host side of the policy/value net: weight folding/packing, TMA tensor maps, the
// layer schedule, and the extern "C" entry points diee_net_* (include/diee.h).
// Reference: src/alphazero/nnet.rs:57-155 (architecture, registration order, forward_t / forward_policy).
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstring>
#include <vector>

#include "ctx.h"
#include "net_launch.h"

using namespace diee;

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// activations: bf16 NHWC [n][4][6][C]; box = [64 ch][6][4][nb boards], 128B swizzle, OOB -> zero
static bool make_act_map(CUtensorMap *m, const void *base, int n_boards, int C, int nb = 16) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) return false;
    cuuint64_t dims[4] = {(cuuint64_t)C, 6, 4, (cuuint64_t)n_boards};
    cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)C * 12, (cuuint64_t)C * 48};
    cuuint32_t box[4] = {64, 6, 4, (cuuint32_t)nb};
    cuuint32_t es[4] = {1, 1, 1, 1};
    return fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void *>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// weights: bf16 [c_out][K]; box = [64 k][bn rows]
static bool make_w_map(CUtensorMap *m, const void *base, int c_out, int K, int bn) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) return false;
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)c_out};
    cuuint64_t strides[1] = {(cuuint64_t)K * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)bn};
    cuuint32_t es[2] = {1, 1};
    return fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// generic K-major bf16 matrix [rows][K]; box = [64 k][box_rows]
static bool make_kmajor_map(CUtensorMap *m, const void *base, long long rows, int K, int box_rows) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) return false;
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)K * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
    cuuint32_t es[2] = {1, 1};
    return fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

struct ConvLayer {
    __nv_bfloat16 *w = nullptr;  // [c_out_pad][K]
    __nv_bfloat16 *w3 = nullptr; // split precision: [c_out_pad][3 planes][K] = (s0, r1, r2) in units of wscale[co] (net_kernels.cu)
    float *wscale = nullptr;     // split precision: [c_out_pad], the power-of-two unit of output channel co's planes
    float *wf32 = nullptr;       // fp32 mode: [9][c_in][c_out] (BatchNorm folded)
    int c_in = 0, c_out = 0;
    float *bias = nullptr;       // [c_out_pad]
    CUtensorMap wmap, wmap3, wmap3_n64, wmap3_n32;
    CUtensorMap wmap_n64, wmap_n32;  // the same weights through 64- and 32-row boxes: small batches run smaller CTA tiles
    int c_out_pad = 0, K = 0, ntaps = 9, chunks = 4, bn = 128;
};

// Split precision: the plane pairs (i, j) of activations x weights, packed 2+2 bits each.  The big pair (0, 0) --
// integer digits, accumulated EXACTLY -- is its own launch; the five small ones run before it, smallest terms first.
static constexpr uint32_t PAIRS_SMALL5 = (2u << 0 | 0u << 2) | (0u << 4 | 2u << 6) | (1u << 8 | 1u << 10) | (1u << 12 | 0u << 14) |
                                         (0u << 16 | 1u << 18);
static constexpr uint32_t PAIRS_BIG = 0u;  // (0, 0)
// the first layer's activations are small integers (one exact plane): plane 0 against the weights' r2, r1
static constexpr uint32_t PAIRS_SMALL2 = (0u << 0 | 2u << 2) | (0u << 4 | 1u << 6);

// Digit width of plane 0: |s0| <= 2^sb.  9 * C_in products of two digits must stay below 2^24 for the sum to be exact.
static int split_digit_bits(int K) {
    int sb = 6;
    while (sb > 1 && (double)K * std::ldexp(1.0, 2 * sb) >= 16777216.0) --sb;
    return sb;
}

// fp32 mode weights from the packed fp32 matrix wf [c_out_pad][K], K index = tap * c_in + ci
static int32_t upload_f32(diee_ctx *ctx, ConvLayer &L, const std::vector<float> &wf, int c_in, int c_out) {
    std::vector<float> w((size_t)9 * c_in * c_out);
    for (int co = 0; co < c_out; ++co)
        for (int tap = 0; tap < 9; ++tap)
            for (int ci = 0; ci < c_in; ++ci) w[((size_t)tap * c_in + ci) * c_out + co] = wf[(size_t)co * L.K + (size_t)tap * c_in + ci];
    L.c_in = c_in; L.c_out = c_out;
    CU(cudaMalloc(&L.wf32, w.size() * sizeof(float)));
    CU(cudaMemcpy(L.wf32, w.data(), w.size() * sizeof(float), cudaMemcpyHostToDevice));
    return DIEE_OK;
}

static int32_t upload_split(diee_ctx *ctx, ConvLayer &L, const std::vector<float> &wf) {
    // wf: [c_out_pad][K] fp32 (BatchNorm already folded).  Per output channel: unit = 2^(e - sb) with 2^e > max |w|,
    // s0 = rint(w / unit) (an integer, |s0| <= 2^sb), r1 and r2 the bf16 roundings of the remainder.
    const int sb = split_digit_bits(L.K);
    std::vector<__nv_bfloat16> w3((size_t)L.c_out_pad * 3 * L.K);
    std::vector<float> ws((size_t)L.c_out_pad, 1.f);
    for (int co = 0; co < L.c_out_pad; ++co) {
        float mx = 0.f;
        for (int k = 0; k < L.K; ++k) mx = std::fmax(mx, std::fabs(wf[(size_t)co * L.K + k]));
        int e = 0;
        if (mx > 0.f) std::frexp(mx, &e);
        const float unit = std::ldexp(1.f, e - sb), inv = std::ldexp(1.f, sb - e);
        ws[co] = unit;
        for (int k = 0; k < L.K; ++k) {
            const float u = wf[(size_t)co * L.K + k] * inv;
            const float s0 = std::nearbyint(u);
            const float d1 = u - s0;
            const __nv_bfloat16 r1 = __float2bfloat16(d1);
            const __nv_bfloat16 r2 = __float2bfloat16(d1 - __bfloat162float(r1));
            w3[((size_t)co * 3 + 0) * L.K + k] = __float2bfloat16(s0);
            w3[((size_t)co * 3 + 1) * L.K + k] = r1;
            w3[((size_t)co * 3 + 2) * L.K + k] = r2;
        }
    }
    CU(cudaMalloc(&L.w3, w3.size() * sizeof(__nv_bfloat16)));
    CU(cudaMemcpy(L.w3, w3.data(), w3.size() * sizeof(__nv_bfloat16), cudaMemcpyHostToDevice));
    CU(cudaMalloc(&L.wscale, ws.size() * sizeof(float)));
    CU(cudaMemcpy(L.wscale, ws.data(), ws.size() * sizeof(float), cudaMemcpyHostToDevice));
    return DIEE_OK;
}

struct diee_net {
    int filters = 0, blocks = 0;
    std::vector<ConvLayer> convs;  // init, then conv1/conv2 per block
    ConvLayer pconv, vconv;
    __nv_bfloat16 *wp = nullptr;  // policy Linear weight [1408][768] bf16, K index = pos*32 + c
    CUtensorMap wp_map;
    float *bp = nullptr, *wv = nullptr;
    float bv = 0.f;
    // activation scratch (grown on demand)
    float *wpT = nullptr;  // split precision: policy Linear weight fp32 [768][1352], same K order
    int precision = DIEE_NET_SPLIT3;  // inside the reference's fp32 tolerance; bf16 is the explicit fast mode
    DevBuf in0, actA, actB, actC, pfeat, vfeat, s_states, s_policy, s_value;
    DevBuf planes, small, board_max, row_scale;  // split precision: operand planes, the small pairs' sum, per-board maxima / units
    int64_t param_count = 0;
};

static float bf16_round(float x) { return __bfloat162float(__float2bfloat16(x)); }

// conv (+ eval-mode BatchNorm, eps 1e-5) folded in fp64 and packed [c_out][tap][c_in] as bf16
static int32_t build_conv(diee_ctx *ctx, ConvLayer &L, const float *w, const float *b, const float *g, const float *beta,
                          const float *mean, const float *var, int c_out, int c_in, int c_out_pad, int k_pad, int bn) {
    const int K = k_pad ? k_pad : 9 * c_in;
    std::vector<__nv_bfloat16> wp((size_t)c_out_pad * K, __float2bfloat16(0.f));
    std::vector<float> wf((size_t)c_out_pad * K, 0.f);
    std::vector<float> bp((size_t)c_out_pad, 0.f);
    for (int co = 0; co < c_out; ++co) {
        const double scale = (double)g[co] / std::sqrt((double)var[co] + 1e-5);
        bp[co] = (float)(((double)b[co] - (double)mean[co]) * scale + (double)beta[co]);
        for (int ci = 0; ci < c_in; ++ci)
            for (int tap = 0; tap < 9; ++tap) {
                const double v = (double)w[((size_t)co * c_in + ci) * 9 + tap] * scale;
                wp[(size_t)co * K + (size_t)tap * c_in + ci] = __float2bfloat16((float)v);
                wf[(size_t)co * K + (size_t)tap * c_in + ci] = (float)v;
            }
    }
    L.c_out_pad = c_out_pad; L.K = K; L.bn = bn;
    L.ntaps = k_pad ? 1 : 9;
    L.chunks = k_pad ? k_pad / 64 : c_in / 64;
    CU(cudaMalloc(&L.w, wp.size() * sizeof(__nv_bfloat16)));
    CU(cudaMalloc(&L.bias, bp.size() * sizeof(float)));
    CU(cudaMemcpy(L.w, wp.data(), wp.size() * sizeof(__nv_bfloat16), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(L.bias, bp.data(), bp.size() * sizeof(float), cudaMemcpyHostToDevice));
    if (!make_w_map(&L.wmap, L.w, c_out_pad, K, bn)) return fail(ctx, DIEE_ERR_CUDA, "cuTensorMapEncodeTiled (weights) failed");
    if (!make_w_map(&L.wmap_n32, L.w, c_out_pad, K, 32) || !make_w_map(&L.wmap_n64, L.w, c_out_pad, K, c_out_pad >= 64 ? 64 : 32))
        return fail(ctx, DIEE_ERR_CUDA, "cuTensorMapEncodeTiled (weights) failed");
    int32_t rc = upload_split(ctx, L, wf);
    if (rc != DIEE_OK) return rc;
    if (!make_w_map(&L.wmap3, L.w3, c_out_pad, 3 * K, bn)) return fail(ctx, DIEE_ERR_CUDA, "cuTensorMapEncodeTiled (split weights) failed");
    if (c_out_pad >= 64 && !make_w_map(&L.wmap3_n64, L.w3, c_out_pad, 3 * K, 64)) return fail(ctx, DIEE_ERR_CUDA, "cuTensorMapEncodeTiled (split weights) failed");
    if (c_out_pad >= 32 && !make_w_map(&L.wmap3_n32, L.w3, c_out_pad, 3 * K, 32)) return fail(ctx, DIEE_ERR_CUDA, "cuTensorMapEncodeTiled (split weights) failed");
    return upload_f32(ctx, L, wf, c_in, c_out);
}

static int32_t net_build(diee_ctx *ctx, diee_net *net, const float *const *t, int blocks, int F);

extern "C" {

int32_t diee_net_create(diee_ctx *ctx, int32_t game_kind, const float *const *t, const int64_t *numels, int32_t n_tensors,
                        diee_net **out) {
    if (!ctx || !t || !numels || !out) return fail(ctx, DIEE_ERR_INVALID, "net_create: bad argument");
    *out = nullptr;
    if (game_kind != DIEE_GAME_BACKGAMMON) return fail(ctx, DIEE_ERR_INVALID, "net_create: only the backgammon geometry (6x4x6 input) is built");
    if (n_tensors < 22 + 12 || (n_tensors - 22) % 12 != 0) return fail(ctx, DIEE_ERR_INVALID, "net_create: expected 22 + 12*blocks tensors, got %d", n_tensors);
    const int blocks = (n_tensors - 22) / 12;
    if (numels[0] % 54 != 0) return fail(ctx, DIEE_ERR_INVALID, "net_create: init conv weight has %lld elements", (long long)numels[0]);
    const int F = (int)(numels[0] / 54);
    if (F % 128 != 0 || F > 1024) return fail(ctx, DIEE_ERR_INVALID, "net_create: filters=%d must be a multiple of 128", F);
    // shape checks along the registration order (nnet.rs:62-98, :37-44)
    auto expect = [&](int idx, int64_t want) { return numels[idx] == want; };
    int idx = 0;
    bool ok = expect(0, (int64_t)F * 54) && expect(1, F);
    for (int k = 2; k < 6; ++k) ok = ok && expect(k, F);
    idx = 6;
    for (int b = 0; b < blocks && ok; ++b) {
        ok = expect(idx, (int64_t)F * F * 9) && expect(idx + 1, F) && expect(idx + 2, (int64_t)F * F * 9) && expect(idx + 3, F);
        for (int k = 4; k < 12; ++k) ok = ok && expect(idx + k, F);
        idx += 12;
    }
    ok = ok && expect(idx, (int64_t)32 * F * 9) && expect(idx + 1, 32);
    for (int k = 2; k < 6; ++k) ok = ok && expect(idx + k, 32);
    ok = ok && expect(idx + 6, (int64_t)DIEE_ACTION_SPACE * 768) && expect(idx + 7, DIEE_ACTION_SPACE);
    ok = ok && expect(idx + 8, (int64_t)3 * F * 9) && expect(idx + 9, 3);
    for (int k = 10; k < 14; ++k) ok = ok && expect(idx + k, 3);
    ok = ok && expect(idx + 14, 72) && expect(idx + 15, 1);
    if (!ok) return fail(ctx, DIEE_ERR_INVALID, "net_create: tensor shapes do not match the ResNet registration order (nnet.rs:57-107)");

    CU(cudaSetDevice(ctx->device));
    if (!get_encode_fn()) return fail(ctx, DIEE_ERR_CUDA, "cuTensorMapEncodeTiled is not available");
    diee_net *net = new diee_net();
    net->filters = F; net->blocks = blocks;
    for (int i = 0; i < n_tensors; ++i) net->param_count += numels[i];
    // Every layer is built IN PLACE inside the net (the vector never reallocates), so each device buffer belongs to the
    // net from the moment it exists and one diee_net_destroy cleans up after a failure anywhere below.
    net->convs.reserve((size_t)1 + 2 * (size_t)blocks);
    const int32_t rc_build = net_build(ctx, net, t, blocks, F);
    if (rc_build != DIEE_OK) {
        const std::string why = ctx->err;  // destroy does not touch it, but keep the first message anyway
        diee_net_destroy(ctx, net);
        ctx->err = why;
        return rc_build;
    }
    // The uploads above are cudaMemcpy calls from pageable host memory: such a copy may return once the data is staged, before
    // the DMA has landed, and the engine's streams are non-blocking (they do not order themselves behind the null stream).
    CU(cudaDeviceSynchronize());
    *out = net;
    return DIEE_OK;
}

}  // extern "C"

static int32_t net_build(diee_ctx *ctx, diee_net *net, const float *const *t, int blocks, int F) {
    int32_t rc;
    int idx;
    // init conv: C_in = 6, K = 54 padded to 64, one "tap" over the im2col operand
    {
        // repack [F][6][3][3] -> build_conv's [co][tap][ci] with c_in = 6, then pad K to 64
        net->convs.emplace_back();
        ConvLayer &L = net->convs.back();
        std::vector<float> w6((size_t)F * 54);
        memcpy(w6.data(), t[0], sizeof(float) * w6.size());
        // build with k_pad = 64: K index = tap*6 + ci (< 54)
        const int K = 64;
        std::vector<__nv_bfloat16> wp((size_t)F * K, __float2bfloat16(0.f));
        std::vector<float> wf0((size_t)F * K, 0.f);
        std::vector<float> bpv((size_t)F, 0.f);
        for (int co = 0; co < F; ++co) {
            const double scale = (double)t[2][co] / std::sqrt((double)t[5][co] + 1e-5);
            bpv[co] = (float)(((double)t[1][co] - (double)t[4][co]) * scale + (double)t[3][co]);
            for (int ci = 0; ci < 6; ++ci)
                for (int tap = 0; tap < 9; ++tap) {
                    const float v = (float)((double)w6[((size_t)co * 6 + ci) * 9 + tap] * scale);
                    wp[(size_t)co * K + tap * 6 + ci] = __float2bfloat16(v);
                    wf0[(size_t)co * K + tap * 6 + ci] = v;
                }
        }
        L.c_out_pad = F; L.K = K; L.bn = 128; L.ntaps = 1; L.chunks = 1;
        CU(cudaMalloc(&L.w, wp.size() * sizeof(__nv_bfloat16)));
        CU(cudaMalloc(&L.bias, bpv.size() * sizeof(float)));
        CU(cudaMemcpy(L.w, wp.data(), wp.size() * sizeof(__nv_bfloat16), cudaMemcpyHostToDevice));
        CU(cudaMemcpy(L.bias, bpv.data(), bpv.size() * sizeof(float), cudaMemcpyHostToDevice));
        if (!make_w_map(&L.wmap, L.w, F, K, 128)) return fail(ctx, DIEE_ERR_CUDA, "cuTensorMapEncodeTiled (weights) failed");
        if (!make_w_map(&L.wmap_n32, L.w, F, K, 32) || !make_w_map(&L.wmap_n64, L.w, F, K, 64))
            return fail(ctx, DIEE_ERR_CUDA, "cuTensorMapEncodeTiled (weights) failed");
        rc = upload_split(ctx, L, wf0);
        if (rc != DIEE_OK) return rc;
        rc = upload_f32(ctx, L, wf0, 6, F);
        if (rc != DIEE_OK) return rc;
        if (!make_w_map(&L.wmap3, L.w3, F, 3 * K, 128)) return fail(ctx, DIEE_ERR_CUDA, "cuTensorMapEncodeTiled (split weights) failed");
    }
    idx = 6;
    for (int b = 0; b < blocks; ++b) {  // ResBlock::new registers conv1, conv2, bn1, bn2 (nnet.rs:37-44)
        net->convs.emplace_back();
        rc = build_conv(ctx, net->convs.back(), t[idx], t[idx + 1], t[idx + 4], t[idx + 5], t[idx + 6], t[idx + 7], F, F, F, 0, 128);
        if (rc != DIEE_OK) return rc;
        net->convs.emplace_back();
        rc = build_conv(ctx, net->convs.back(), t[idx + 2], t[idx + 3], t[idx + 8], t[idx + 9], t[idx + 10], t[idx + 11], F, F, F, 0, 128);
        if (rc != DIEE_OK) return rc;
        idx += 12;
    }
    rc = build_conv(ctx, net->pconv, t[idx], t[idx + 1], t[idx + 2], t[idx + 3], t[idx + 4], t[idx + 5], 32, F, 32, 0, 32);
    if (rc != DIEE_OK) return rc;
    {   // policy Linear(768 -> 1352): torch flattens NCHW (feature = c*24 + pos); ours is NHWC (pos*32 + c)
        std::vector<__nv_bfloat16> wp((size_t)1408 * 768, __float2bfloat16(0.f));
        const float *W = t[idx + 6];
        for (int j = 0; j < DIEE_ACTION_SPACE; ++j)
            for (int c = 0; c < 32; ++c)
                for (int pos = 0; pos < 24; ++pos) wp[(size_t)j * 768 + pos * 32 + c] = __float2bfloat16(W[(size_t)j * 768 + c * 24 + pos]);
        CU(cudaMalloc(&net->wp, wp.size() * sizeof(__nv_bfloat16)));
        CU(cudaMemcpy(net->wp, wp.data(), wp.size() * sizeof(__nv_bfloat16), cudaMemcpyHostToDevice));
        if (!make_kmajor_map(&net->wp_map, net->wp, 1408, 768, 128)) return fail(ctx, DIEE_ERR_CUDA, "cuTensorMapEncodeTiled (policy linear) failed");
        std::vector<float> wT((size_t)768 * DIEE_ACTION_SPACE);
        for (int j = 0; j < DIEE_ACTION_SPACE; ++j)
            for (int c = 0; c < 32; ++c)
                for (int pos = 0; pos < 24; ++pos) wT[(size_t)(pos * 32 + c) * DIEE_ACTION_SPACE + j] = W[(size_t)j * 768 + c * 24 + pos];
        CU(cudaMalloc(&net->wpT, wT.size() * sizeof(float)));
        CU(cudaMemcpy(net->wpT, wT.data(), wT.size() * sizeof(float), cudaMemcpyHostToDevice));
        CU(cudaMalloc(&net->bp, DIEE_ACTION_SPACE * sizeof(float)));
        CU(cudaMemcpy(net->bp, t[idx + 7], DIEE_ACTION_SPACE * sizeof(float), cudaMemcpyHostToDevice));
    }
    rc = build_conv(ctx, net->vconv, t[idx + 8], t[idx + 9], t[idx + 10], t[idx + 11], t[idx + 12], t[idx + 13], 3, F, 16, 0, 16);
    if (rc != DIEE_OK) return rc;
    {   // value Linear(72 -> 1): feature c*24+pos -> pos*16 + c
        std::vector<float> wv((size_t)24 * 16, 0.f);
        for (int c = 0; c < 3; ++c)
            for (int pos = 0; pos < 24; ++pos) wv[pos * 16 + c] = t[idx + 14][c * 24 + pos];
        CU(cudaMalloc(&net->wv, wv.size() * sizeof(float)));
        CU(cudaMemcpy(net->wv, wv.data(), wv.size() * sizeof(float), cudaMemcpyHostToDevice));
        net->bv = t[idx + 15][0];
    }
    (void)bf16_round;
    return DIEE_OK;
}

extern "C" {

int32_t diee_net_destroy(diee_ctx *ctx, diee_net *net) {
    if (!ctx || !net) return DIEE_ERR_INVALID;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (ConvLayer &L : net->convs) { cudaFree(L.w); cudaFree(L.w3); cudaFree(L.wscale); cudaFree(L.wf32); cudaFree(L.bias); }
    cudaFree(net->pconv.w); cudaFree(net->pconv.w3); cudaFree(net->pconv.wscale); cudaFree(net->pconv.wf32); cudaFree(net->pconv.bias);
    cudaFree(net->vconv.w); cudaFree(net->vconv.w3); cudaFree(net->vconv.wscale); cudaFree(net->vconv.wf32); cudaFree(net->vconv.bias);
    cudaFree(net->wp); cudaFree(net->wpT); cudaFree(net->bp); cudaFree(net->wv);
    DevBuf *bufs[] = {&net->in0, &net->actA, &net->actB, &net->actC, &net->pfeat, &net->vfeat, &net->s_states, &net->s_policy, &net->s_value,
                      &net->planes, &net->small, &net->board_max, &net->row_scale};
    for (DevBuf *b : bufs)
        if (b->p) cudaFree(b->p);
    delete net;
    return DIEE_OK;
}

// The CTA tile of the tower: nb boards x bn output channels.  A CTA's time is set by its K loop (a K-block costs ~0.2 us plus
// its MMAs) whatever the batch, so the smallest tile that still fits the whole layer in one wave of CTAs wins: 16 x 128 at
// 1,024 boards, 8 x 64 at 256, 4 x 32 at 64 -- the long tail of a self-play batch.  DIEE_CONV_TILE="nb,bn" forces one.
static void pick_conv_tile(diee_ctx *ctx, int n, int F, int &nb, int &bn) {
    nb = 16; bn = 128;
    static int sms = 0;
    if (sms == 0) { cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device); if (sms <= 0) sms = 148; }
    const int cand[5][2] = {{4, 32}, {8, 32}, {8, 64}, {8, 128}, {16, 128}};  // ascending bytes per K-block
    for (int c = 0; c < 5; ++c) {
        const int cnb = cand[c][0], cbn = cand[c][1];
        if (F % cbn != 0) continue;
        if ((long long)((n + cnb - 1) / cnb) * (F / cbn) <= sms) { nb = cnb; bn = cbn; break; }
    }
    if (const char *force_tile = getenv("DIEE_CONV_TILE")) {  // "nb,bn" (experiments, tests)
        int a = 0, b = 0;
        if (sscanf(force_tile, "%d,%d", &a, &b) == 2 && F % b == 0) { nb = a; bn = b; }
    }
}

// DIEE_CONV_2CTA=n: run the tower's 16-board x 128-channel tile on CTA pairs (cta_group::2, net_kernels.cu) from n boards on.
// OFF by default: bit-identical to the single-CTA form (tests/test_gpu_net.py) but measured SLOWER on B200 -- one
// alpha_mcts_parallel of 1,024 games: bf16 152.8 -> 160.1 ms, split3 827.6 -> 860.7 ms -- so the shared-memory port is not
// what holds the main loop at 64 % tensor-pipe activity (DESIGN.md 3.6 records the experiment).
static bool use_cta_pairs(int n) {
    const char *e = getenv("DIEE_CONV_2CTA");
    const int min_n = e ? atoi(e) : 0;
    return min_n > 0 && n >= min_n;
}

// forward_t (nnet.rs:120-133): policy = softmax(policy_head(tower(x))), value = tanh(value_head(tower(x)))
int32_t diee_net_forward_dev(diee_ctx *ctx, diee_net *net, const diee_bg_state *states, int32_t n, float *policy_out,
                             float *value_out) {
    if (!ctx || !net || n < 0 || (n && (!states || !policy_out || !value_out))) return fail(ctx, DIEE_ERR_INVALID, "net_forward: bad argument");
    if (n == 0) return DIEE_OK;
    CU(cudaSetDevice(ctx->device));
    const int F = net->filters;
    const size_t rows = (size_t)n * 24;
    if (net->precision == DIEE_NET_FP32) {
        // the reference's own arithmetic: fp32 FMAs on the CUDA cores (parity mode)
        RESERVE(net->actA, rows * F * 4);
        RESERVE(net->actB, rows * F * 4);
        RESERVE(net->actC, rows * F * 4);
        RESERVE(net->pfeat, rows * 32 * 4);
        RESERVE(net->vfeat, rows * 16 * 4);
        cudaStream_t st = ctx->stream;
        float *bx = (float *)net->actA.p, *by = (float *)net->actB.p, *bz = (float *)net->actC.p;
        const ConvLayer &L0 = net->convs[0];
        CU(launch_conv_f32(st, nullptr, states, n, 6, L0.wf32, L0.bias, nullptr, bx, F, F, 1));
        ctx->launches += 1;
        for (int b = 0; b < net->blocks; ++b) {
            const ConvLayer &c1 = net->convs[1 + 2 * b], &c2 = net->convs[2 + 2 * b];
            CU(launch_conv_f32(st, bx, nullptr, n, F, c1.wf32, c1.bias, nullptr, by, F, F, 1));
            CU(launch_conv_f32(st, by, nullptr, n, F, c2.wf32, c2.bias, bx, bz, F, F, 1));
            ctx->launches += 2;
            float *tp = bx; bx = bz; bz = tp;
        }
        CU(cudaMemsetAsync(net->vfeat.p, 0, rows * 16 * 4, st));  // channels 3..15 of the value features are padding
        CU(launch_conv_f32(st, bx, nullptr, n, F, net->pconv.wf32, net->pconv.bias, nullptr, (float *)net->pfeat.p, 32, 32, 1));
        CU(launch_conv_f32(st, bx, nullptr, n, F, net->vconv.wf32, net->vconv.bias, nullptr, (float *)net->vfeat.p, 3, 16, 1));
        CU(launch_heads_f32(st, (const float *)net->pfeat.p, net->wpT, net->bp, (const float *)net->vfeat.p, net->wv, net->bv, n, policy_out, value_out));
        ctx->launches += 4;
        return DIEE_OK;
    }
    if (net->precision == DIEE_NET_SPLIT3) {
        // the tensor-core mode inside the fp32 tolerance (net_kernels.cu, ConvEpi): per convolution one launch for the
        // five small plane pairs (fp32 scratch), one for the exact integer pair whose epilogue finishes the layer in fp32,
        // then the planes of the new activations; heads in fp32
        const int sb = split_digit_bits(9 * F);
        RESERVE(net->in0, rows * 64 * 2);
        RESERVE(net->actA, rows * F * 4);
        RESERVE(net->actB, rows * F * 4);
        RESERVE(net->actC, rows * F * 4);
        RESERVE(net->small, rows * F * 4);
        RESERVE(net->planes, rows * F * 6);
        RESERVE(net->pfeat, rows * 32 * 4);
        RESERVE(net->vfeat, rows * 16 * 4);
        if (net->board_max.cap < (size_t)n * 4) {
            RESERVE(net->board_max, (size_t)n * 4);
            CU(cudaMemsetAsync(net->board_max.p, 0, net->board_max.cap, ctx->stream));  // (split_planes_kernel leaves it zeroed)
        }
        RESERVE(net->row_scale, (size_t)n * 4);
        int nb, bn;
        pick_conv_tile(ctx, n, F, nb, bn);  // the tower's tile (the first layer and the heads keep 16-board tiles)
        CUtensorMap m_in, mP, mPt;
        if (!make_act_map(&m_in, net->in0.p, n, 64) || !make_act_map(&mP, net->planes.p, n, 3 * F) ||
            !make_act_map(&mPt, net->planes.p, n, 3 * F, nb))
            return fail(ctx, DIEE_ERR_CUDA, "cuTensorMapEncodeTiled (activations) failed");
        cudaStream_t st = ctx->stream;
        float *small = (float *)net->small.p, *rscale = (float *)net->row_scale.p;
        unsigned int *bmax = (unsigned int *)net->board_max.p;
        // one convolution: x planes (map mx) * w planes -> fp32 `out` [rows][c_out]; first = the im2col layer (one exact plane, unit 1)
        const bool pairs2 = use_cta_pairs(n);
        auto conv = [&](const CUtensorMap &mx, const ConvLayer &L, bool first, const float *residual, float *out, int c_out, bool want_max) -> int32_t {
            SplitEpilogue sp{small, first ? nullptr : rscale, L.wscale, residual, want_max ? bmax : nullptr, 0};
            const bool tower = !first && L.bn == 128;
            if (pairs2 && tower && nb == 16 && bn == 128) {  // tower layers on CTA pairs
                CU(launch_conv_pair(st, mx, L.wmap3_n64, n, L.ntaps, L.chunks, nullptr, nullptr, small, 1, c_out, 0, 5, PAIRS_SMALL5, F, L.K));
                CU(launch_conv_pair(st, mx, L.wmap3_n64, n, L.ntaps, L.chunks, L.bias, nullptr, out, 2, c_out, 1, 1, PAIRS_BIG, F, L.K, &sp));
            } else if (tower && !(nb == 16 && bn == 128)) {
                // small batches: smaller tiles over all SMs, and ONE launch per layer -- the tile's two accumulators fit TMEM
                // (2 x MT x BN <= 512 columns), so the five small pairs and the exact pair run back to back in one K loop
                // and the epilogue adds them in registers: same bits, no scratch round trip, one pipeline ramp instead of two
                const CUtensorMap &wm = bn == 128 ? L.wmap3 : bn == 64 ? L.wmap3_n64 : L.wmap3_n32;
                static const bool no_fuse = getenv("DIEE_SPLIT_FUSE") && atoi(getenv("DIEE_SPLIT_FUSE")) == 0;
                if (!no_fuse) {
                    sp.addend = nullptr; sp.fused = 1;
                    CU(launch_conv_tile(st, bn, nb, mPt, wm, n, L.ntaps, L.chunks, L.bias, nullptr, out, 2, c_out, 1, 6, PAIRS_SMALL5, F, L.K, &sp));
                    ctx->launches -= 1;
                } else {
                    CU(launch_conv_tile(st, bn, nb, mPt, wm, n, L.ntaps, L.chunks, nullptr, nullptr, small, 1, c_out, 0, 5, PAIRS_SMALL5, F, L.K));
                    CU(launch_conv_tile(st, bn, nb, mPt, wm, n, L.ntaps, L.chunks, L.bias, nullptr, out, 2, c_out, 1, 1, PAIRS_BIG, F, L.K, &sp));
                }
            } else if (!first && L.bn < 128 && !(getenv("DIEE_SPLIT_FUSE") && atoi(getenv("DIEE_SPLIT_FUSE")) == 0)) {
                // the two head convolutions (32 / 16 output channels): their accumulators are narrow, one fused launch each
                sp.addend = nullptr; sp.fused = 1;
                CU(launch_conv(st, L.bn, mx, L.wmap3, n, L.ntaps, L.chunks, L.bias, nullptr, out, 2, c_out, 1, 6, PAIRS_SMALL5, F, L.K, &sp));
                ctx->launches -= 1;
            } else {
                CU(launch_conv(st, L.bn, mx, L.wmap3, n, L.ntaps, L.chunks, nullptr, nullptr, small, 1, c_out, 0, first ? 2 : 5,
                               first ? PAIRS_SMALL2 : PAIRS_SMALL5, first ? 0 : F, L.K));
                CU(launch_conv(st, L.bn, mx, L.wmap3, n, L.ntaps, L.chunks, L.bias, nullptr, out, 2, c_out, 1, 1, PAIRS_BIG, first ? 0 : F, L.K, &sp));
            }
            ctx->launches += 2;
            return DIEE_OK;
        };
        auto planes_of = [&](const float *y) -> int32_t {
            CU(launch_split_planes(st, y, bmax, n, F, sb, net->planes.p, rscale));
            ctx->launches += 1;
            return DIEE_OK;
        };
        CU(launch_encode_im2col(st, states, n, net->in0.p));
        ctx->launches += 1;
        float *bx = (float *)net->actA.p, *by = (float *)net->actB.p, *bz = (float *)net->actC.p;
        int32_t rc = conv(m_in, net->convs[0], true, nullptr, bx, F, true);
        if (rc != DIEE_OK) return rc;
        for (int b = 0; b < net->blocks; ++b) {
            const ConvLayer &c1 = net->convs[1 + 2 * b], &c2 = net->convs[2 + 2 * b];
            if ((rc = planes_of(bx)) != DIEE_OK) return rc;
            if ((rc = conv(mP, c1, false, nullptr, by, F, true)) != DIEE_OK) return rc;
            if ((rc = planes_of(by)) != DIEE_OK) return rc;
            if ((rc = conv(mP, c2, false, bx, bz, F, true)) != DIEE_OK) return rc;
            float *tp = bx; bx = bz; bz = tp;
        }
        if ((rc = planes_of(bx)) != DIEE_OK) return rc;
        if ((rc = conv(mP, net->pconv, false, nullptr, (float *)net->pfeat.p, 32, false)) != DIEE_OK) return rc;
        if ((rc = conv(mP, net->vconv, false, nullptr, (float *)net->vfeat.p, 16, false)) != DIEE_OK) return rc;
        CU(launch_heads_f32(st, (const float *)net->pfeat.p, net->wpT, net->bp, (const float *)net->vfeat.p, net->wv, net->bv, n, policy_out, value_out));
        ctx->launches += 2;
        return DIEE_OK;
    }
    RESERVE(net->in0, rows * 64 * 2);
    RESERVE(net->actA, rows * F * 2);
    RESERVE(net->actB, rows * F * 2);
    RESERVE(net->actC, rows * F * 2);
    RESERVE(net->pfeat, rows * 32 * 2);
    RESERVE(net->vfeat, rows * 16 * 4);
    int nb, bn;
    pick_conv_tile(ctx, n, F, nb, bn);
    CUtensorMap m_in, mA, mB, mC, m_head;
    if (!make_act_map(&m_in, net->in0.p, n, 64, nb) || !make_act_map(&mA, net->actA.p, n, F, nb) || !make_act_map(&mB, net->actB.p, n, F, nb) ||
        !make_act_map(&mC, net->actC.p, n, F, nb))
        return fail(ctx, DIEE_ERR_CUDA, "cuTensorMapEncodeTiled (activations) failed");
    cudaStream_t st = ctx->stream;
    CU(launch_encode_im2col(st, states, n, net->in0.p));
    auto wmap_of = [bn](const ConvLayer &L) -> const CUtensorMap & { return bn == 128 ? L.wmap : bn == 64 ? L.wmap_n64 : L.wmap_n32; };
    const ConvLayer &L0 = net->convs[0];
    CU(launch_conv_tile(st, bn, nb, m_in, wmap_of(L0), n, L0.ntaps, L0.chunks, L0.bias, nullptr, net->actA.p, 0, F, 1));
    ctx->launches += 2;
    // x lives in A; y = relu(bn1(conv1(x))) -> B; x' = relu(bn2(conv2(y)) + x) -> C; rotate
    void *bx = net->actA.p, *by = net->actB.p, *bz = net->actC.p;
    CUtensorMap *mx = &mA, *my = &mB, *mz = &mC;
    for (int b = 0; b < net->blocks; ++b) {
        const ConvLayer &c1 = net->convs[1 + 2 * b], &c2 = net->convs[2 + 2 * b];
        if (nb == 16 && bn == 128 && use_cta_pairs(n)) {
            CU(launch_conv_pair(st, *mx, c1.wmap_n64, n, 9, c1.chunks, c1.bias, nullptr, by, 0, F, 1));
            CU(launch_conv_pair(st, *my, c2.wmap_n64, n, 9, c2.chunks, c2.bias, bx, bz, 0, F, 1));
        } else {
            CU(launch_conv_tile(st, bn, nb, *mx, wmap_of(c1), n, 9, c1.chunks, c1.bias, nullptr, by, 0, F, 1));
            CU(launch_conv_tile(st, bn, nb, *my, wmap_of(c2), n, 9, c2.chunks, c2.bias, bx, bz, 0, F, 1));
        }
        ctx->launches += 2;
        void *tp = bx; bx = bz; bz = tp;
        CUtensorMap *tm = mx; mx = mz; mz = tm;
    }
    // the two head convolutions keep 16-board tiles (their output tiles are 32 / 16 channels wide already)
    if (!make_act_map(&m_head, bx, n, F, 16)) return fail(ctx, DIEE_ERR_CUDA, "cuTensorMapEncodeTiled (activations) failed");
    mx = &m_head;
    CU(launch_conv(st, net->pconv.bn, *mx, net->pconv.wmap, n, 9, net->pconv.chunks, net->pconv.bias, nullptr, net->pfeat.p, 0, 32, 1));
    CU(launch_conv(st, net->vconv.bn, *mx, net->vconv.wmap, n, 9, net->vconv.chunks, net->vconv.bias, nullptr, net->vfeat.p, 1, 16, 1));
    CUtensorMap m_pf;
    if (!make_kmajor_map(&m_pf, net->pfeat.p, n, 768, 128)) return fail(ctx, DIEE_ERR_CUDA, "cuTensorMapEncodeTiled (policy features) failed");
    CU(launch_heads(st, m_pf, net->wp_map, net->bp, (const float *)net->vfeat.p, net->wv, net->bv, n, policy_out, value_out));
    ctx->launches += 4;
    return DIEE_OK;
}

int32_t diee_net_forward(diee_ctx *ctx, diee_net *net, const diee_bg_state *states, int32_t n, float *policy_out, float *value_out) {
    if (!ctx || !net || n < 0 || (n && (!states || !policy_out || !value_out))) return fail(ctx, DIEE_ERR_INVALID, "net_forward: bad argument");
    if (n == 0) return DIEE_OK;
    for (int i = 0; i < n; ++i)
        if (states[i].roll[0] == 0 && states[i].roll[1] == 0) return fail(ctx, DIEE_ERR_NOT_ROLLED, "net_forward: state %d has not been rolled", i);
    CU(cudaSetDevice(ctx->device));
    RESERVE(net->s_states, sizeof(diee_bg_state) * (size_t)n);
    RESERVE(net->s_policy, sizeof(float) * DIEE_ACTION_SPACE * (size_t)n);
    RESERVE(net->s_value, sizeof(float) * (size_t)n);
    CU(cudaMemcpyAsync(net->s_states.p, states, sizeof(diee_bg_state) * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
    int32_t rc = diee_net_forward_dev(ctx, net, (const diee_bg_state *)net->s_states.p, n, (float *)net->s_policy.p, (float *)net->s_value.p);
    if (rc != DIEE_OK) return rc;
    CU(cudaMemcpyAsync(policy_out, net->s_policy.p, sizeof(float) * DIEE_ACTION_SPACE * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaMemcpyAsync(value_out, net->s_value.p, sizeof(float) * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return DIEE_OK;
}

int64_t diee_net_param_count(const diee_net *net) { return net ? net->param_count : 0; }

int32_t diee_net_set_precision(diee_ctx *ctx, diee_net *net, int32_t precision) {
    if (!ctx || !net) return DIEE_ERR_INVALID;
    if (precision != DIEE_NET_BF16 && precision != DIEE_NET_SPLIT3 && precision != DIEE_NET_FP32) return fail(ctx, DIEE_ERR_INVALID, "net_set_precision: unknown precision %d", precision);
    net->precision = precision;
    return DIEE_OK;
}

}  // extern "C"
