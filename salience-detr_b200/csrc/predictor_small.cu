// MaskPredictor of ONE small feature level in two launches -- sdetr_mask_predictor_level.
//
// Reference: models/bricks/salience_transformer.py:16-47 (MaskPredictor: LayerNorm -> Linear(C, C) -> GELU; the second half of
// the channels is replaced by its mean over the level's tokens; Linear(C, C/2) -> GELU -> Linear(C/2, C/4) -> GELU ->
// Linear(C/4, 1)) applied to a level whose tokens were modulated by the upsampled score of the next coarser level (:134-143:
// m + m * bilinear(score, align_corners=True) * alpha[level]).
//
// As library / tensor-core calls this is 15 launches per level (modulate, LayerNorm, 4 GEMMs + cuBLAS helpers, GELU kernels, the
// token mean as two launches, copies): ~80 us for the 546 and 2100 token rows of the two coarsest levels of config 2 -- pure
// launch latency, the arithmetic is 0.06 / 0.22 GMAC.  Here:
//   predictor_front_kernel  8 token rows per CTA: modulate + LayerNorm (one warp per row) -> shared memory; Linear-1 by 256
//                           threads (thread j = output column j, the 8 rows in registers, W1^T rows streamed from L2, coalesced)
//                           -> GELU; columns [0, C/2) go to a (rows, C/2) workspace, columns [C/2, C) are summed over the
//                           CTA's rows into a per-CTA partial (fixed order: reproducible).
//   predictor_back_kernel   8 token rows per CTA: token mean from the partials (CTA order) -> its contribution to Linear-2a is a
//                           per-image vector (the mean is the same for every token), so per token only the local half remains:
//                           Linear-2a (K = C/2) -> GELU -> Linear-2b -> GELU -> Linear-2c -> the level's slice of the raw score.
// fp32 FMA throughout (the same class of arithmetic as the cuBLAS SGEMMs it replaces).  C = 256 only.
#include "common.cuh"

namespace sdetr {

constexpr int kPC = 256, kPHalf = 128, kPQuarter = 64, kPRows = 8;

__device__ __forceinline__ float gelu_exact(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752f)); }

struct PredictorLevel {
    const float *mem;      // first token row of the level in image 0, (HW, C) rows, images mem_bs floats apart
    int64_t mem_bs;
    const float *coarse;   // coarser level's raw scores (Hc * Wc per image, images coarse_bs floats apart) or NULL
    int64_t coarse_bs;
    const float *alpha;
    int alpha_index, H, W, Hc, Wc;
};

__device__ __forceinline__ void cp_async16_p(void *smem_dst, const void *gmem_src) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
}

// Both kernels are latency-bound chains of weight reads from L2, so every weight loop keeps 16 independent loads in flight per
// thread and the K range is split over two thread groups (first version: 4 loads in flight, one group -> 37 / 50 us per launch).
constexpr int kFrontThreads = 512;

__global__ void __launch_bounds__(kFrontThreads) predictor_front_kernel(PredictorLevel lv, const float *__restrict__ ln_g,
                                                                        const float *__restrict__ ln_b, float eps,
                                                                        const float *__restrict__ w1_t /* (C, C): [k][j] */,
                                                                        const float *__restrict__ b1, float *__restrict__ zlocal,
                                                                        float *__restrict__ partial) {
    __shared__ __align__(16) float xs[kPRows][kPC];
    __shared__ float red[kPRows][kPC];
    const int HW = lv.H * lv.W, tile = blockIdx.x, b = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp < kPRows) {  // modulate + LayerNorm: warp w <-> row w of the tile
        const int t = tile * kPRows + warp;
        float4 v[2] = {make_float4(0.f, 0.f, 0.f, 0.f), make_float4(0.f, 0.f, 0.f, 0.f)};
        if (t < HW) {
            float up = 0.f, al = 0.f;
            if (lv.coarse) {
                const int y = t / lv.W, x = t - y * lv.W;
                // ATen upsample_bilinear2d, align_corners=True (same expressions as score_modulate_kernel)
                const float sy = lv.H > 1 ? (float)(lv.Hc - 1) / (float)(lv.H - 1) : 0.f;
                const float sx = lv.W > 1 ? (float)(lv.Wc - 1) / (float)(lv.W - 1) : 0.f;
                const float fy = sy * (float)y, fx = sx * (float)x;
                const int y0 = (int)fy, x0 = (int)fx;
                const int y1 = y0 + (y0 < lv.Hc - 1 ? 1 : 0), x1 = x0 + (x0 < lv.Wc - 1 ? 1 : 0);
                const float ly = fy - (float)y0, lx = fx - (float)x0, hy = 1.f - ly, hx = 1.f - lx;
                const float *cs = lv.coarse + (int64_t)b * lv.coarse_bs;
                up = hy * (hx * __ldg(cs + y0 * lv.Wc + x0) + lx * __ldg(cs + y0 * lv.Wc + x1)) +
                     ly * (hx * __ldg(cs + y1 * lv.Wc + x0) + lx * __ldg(cs + y1 * lv.Wc + x1));
                al = __ldg(lv.alpha + lv.alpha_index);
            }
            const float *src = lv.mem + (int64_t)b * lv.mem_bs + (int64_t)t * kPC;
            float s = 0.f;
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const float4 m = ld_stream_f4(src + i * 128 + lane * 4);
                if (lv.coarse) {  // same expression (and rounding order) as score_modulate_kernel
                    v[i] = make_float4(m.x + m.x * up * al, m.y + m.y * up * al, m.z + m.z * up * al, m.w + m.w * up * al);
                } else {
                    v[i] = m;
                }
                s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            const float mean = s / (float)kPC;
            float ss = 0.f;
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const float dx = v[i].x - mean, dy = v[i].y - mean, dz = v[i].z - mean, dw = v[i].w - mean;
                ss += (dx * dx + dy * dy) + (dz * dz + dw * dw);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
            const float rstd = rsqrtf(ss / (float)kPC + eps);
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int c = i * 128 + lane * 4;
                const float4 g = ldg_f4(ln_g + c), bt = ldg_f4(ln_b + c);
                v[i] = make_float4((v[i].x - mean) * rstd * g.x + bt.x, (v[i].y - mean) * rstd * g.y + bt.y,
                                   (v[i].z - mean) * rstd * g.z + bt.z, (v[i].w - mean) * rstd * g.w + bt.w);
            }
        }
#pragma unroll
        for (int i = 0; i < 2; ++i) *reinterpret_cast<float4 *>(&xs[warp][i * 128 + lane * 4]) = v[i];
    }
    // Linear-1: thread (j, kh) = output column j over k in [128 kh, 128 kh + 128) for the 8 rows; the weight loads of the first
    // batch are issued before the barrier (they do not depend on the rows)
    const int j = threadIdx.x & (kPC - 1), kh = threadIdx.x >> 8;
    const float *wp = w1_t + (int64_t)(kh * 128) * kPC + j;
    float acc[kPRows];
#pragma unroll
    for (int r = 0; r < kPRows; ++r) acc[r] = 0.f;
    float wv[16];
#pragma unroll
    for (int u = 0; u < 16; ++u) wv[u] = __ldg(wp + u * kPC);
    __syncthreads();
#pragma unroll 1
    for (int k0 = 0; k0 < 128; k0 += 16) {
        float wn[16];
        if (k0 + 16 < 128) {
#pragma unroll
            for (int u = 0; u < 16; ++u) wn[u] = __ldg(wp + (k0 + 16 + u) * kPC);
        }
#pragma unroll
        for (int u = 0; u < 16; u += 4) {
#pragma unroll
            for (int r = 0; r < kPRows; ++r) {
                const float4 x = *reinterpret_cast<const float4 *>(&xs[r][kh * 128 + k0 + u]);  // broadcast
                acc[r] = fmaf(x.x, wv[u], acc[r]), acc[r] = fmaf(x.y, wv[u + 1], acc[r]);
                acc[r] = fmaf(x.z, wv[u + 2], acc[r]), acc[r] = fmaf(x.w, wv[u + 3], acc[r]);
            }
        }
        if (k0 + 16 < 128) {
#pragma unroll
            for (int u = 0; u < 16; ++u) wv[u] = wn[u];
        }
    }
    if (kh == 1) {
#pragma unroll
        for (int r = 0; r < kPRows; ++r) red[r][j] = acc[r];
    }
    __syncthreads();
    if (kh == 1) return;
    const float bj = __ldg(b1 + j);
    const int rows_here = min(kPRows, HW - tile * kPRows);
    if (j < kPHalf) {
        float *dst = zlocal + ((int64_t)b * HW + (int64_t)tile * kPRows) * kPHalf + j;
#pragma unroll
        for (int r = 0; r < kPRows; ++r)
            if (r < rows_here) dst[r * kPHalf] = gelu_exact(acc[r] + red[r][j] + bj);
    } else {
        float s = 0.f;
#pragma unroll
        for (int r = 0; r < kPRows; ++r)
            if (r < rows_here) s += gelu_exact(acc[r] + red[r][j] + bj);
        partial[((int64_t)b * gridDim.x + tile) * kPHalf + (j - kPHalf)] = s;
    }
}

constexpr int kBackThreads = 256;
constexpr int kBackSmem = (kPHalf * kPHalf + kPHalf * kPQuarter) * 4;  // W2a^T local half (64 KB) + W2b^T (32 KB)

__global__ void __launch_bounds__(kBackThreads) predictor_back_kernel(const float *__restrict__ zlocal, const float *__restrict__ partial,
                                                                      int tiles_front, int HW, const float *__restrict__ w2a_t /* (C, C/2) */,
                                                                      const float *__restrict__ b2a, const float *__restrict__ w2b_t /* (C/2, C/4) */,
                                                                      const float *__restrict__ b2b, const float *__restrict__ w2c /* C/4 */,
                                                                      const float *__restrict__ b2c, float *__restrict__ out, int64_t out_bs) {
    extern __shared__ __align__(16) float wsm[];
    float *w2a_s = wsm, *w2b_s = wsm + kPHalf * kPHalf;
    __shared__ __align__(16) float zs[kPRows][kPHalf];
    __shared__ __align__(16) float h1s[kPRows][kPHalf];
    __shared__ __align__(16) float h2s[kPRows][kPQuarter];
    __shared__ float mean_s[kPHalf], part_s[kPHalf];
    const int tile = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
    const int o = tid & (kPHalf - 1), h = tid >> 7;
    const int rows_here = min(kPRows, HW - tile * kPRows);
    // the per-token weights stream into shared memory while the mean and its contribution are computed from L2
    for (int i = tid; i < kPHalf * kPHalf / 4; i += kBackThreads) cp_async16_p(w2a_s + 4 * i, w2a_t + 4 * i);
    for (int i = tid; i < kPHalf * kPQuarter / 4; i += kBackThreads) cp_async16_p(w2b_s + 4 * i, w2b_t + 4 * i);
    asm volatile("cp.async.commit_group;" ::: "memory");
    {   // token mean of the global half: the front kernel's partials in CTA order, two halves of the tile range
        const int t0 = h ? tiles_front / 2 : 0, t1 = h ? tiles_front : tiles_front / 2;
        const float *pp = partial + (int64_t)b * tiles_front * kPHalf + o;
        float s = 0.f;
#pragma unroll 16
        for (int t = t0; t < t1; ++t) s += __ldg(pp + (int64_t)t * kPHalf);
        if (h) part_s[o] = s;
        __syncthreads();
        if (!h) mean_s[o] = (s + part_s[o]) / (float)HW;
        const float *zsrc = zlocal + ((int64_t)b * HW + (int64_t)tile * kPRows) * kPHalf + o;
#pragma unroll
        for (int r = 0; r < kPRows / 2; ++r) zs[h * 4 + r][o] = h * 4 + r < rows_here ? zsrc[(h * 4 + r) * kPHalf] : 0.f;
    }
    __syncthreads();
    float g = 0.f;  // the mean's contribution (rows [C/2, C) of W2a^T), k range split over the two thread groups
    {
        const float *wp = w2a_t + (int64_t)(kPHalf + h * 64) * kPHalf + o;
#pragma unroll 1
        for (int k0 = 0; k0 < 64; k0 += 16) {
            float wv[16];
#pragma unroll
            for (int u = 0; u < 16; ++u) wv[u] = __ldg(wp + (k0 + u) * kPHalf);
#pragma unroll
            for (int u = 0; u < 16; ++u) g = fmaf(mean_s[h * 64 + k0 + u], wv[u], g);
        }
    }
    __shared__ float g0_s[kPHalf], g1_s[kPHalf];
    if (h) g1_s[o] = g;
    else g0_s[o] = g;
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    // ---- Linear-2a: thread (o, h) = column o for rows 4h .. 4h + 3, weights from shared memory; every accumulator starts from
    // bias + the mean's contribution (summed in the same order by both groups)
    float acc[4];
    {
        const float gt = (__ldg(b2a + o) + g0_s[o]) + g1_s[o];
#pragma unroll
        for (int r = 0; r < 4; ++r) acc[r] = gt;
    }
#pragma unroll 4
    for (int k = 0; k < kPHalf; k += 4) {
        const float w0 = w2a_s[(k + 0) * kPHalf + o], w1 = w2a_s[(k + 1) * kPHalf + o];
        const float w2 = w2a_s[(k + 2) * kPHalf + o], w3 = w2a_s[(k + 3) * kPHalf + o];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const float4 x = *reinterpret_cast<const float4 *>(&zs[h * 4 + r][k]);
            acc[r] = fmaf(x.x, w0, acc[r]), acc[r] = fmaf(x.y, w1, acc[r]);
            acc[r] = fmaf(x.z, w2, acc[r]), acc[r] = fmaf(x.w, w3, acc[r]);
        }
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) h1s[h * 4 + r][o] = gelu_exact(acc[r]);
    __syncthreads();
    {   // Linear-2b: thread = (column p, row pair rg)
        const int p = tid & (kPQuarter - 1), rg = tid >> 6;  // rg 0..3 -> rows 2 rg, 2 rg + 1
        const float bb = __ldg(b2b + p);
        float a0 = bb, a1 = bb;
#pragma unroll 4
        for (int k = 0; k < kPHalf; k += 4) {
            const float w0 = w2b_s[(k + 0) * kPQuarter + p], w1 = w2b_s[(k + 1) * kPQuarter + p];
            const float w2 = w2b_s[(k + 2) * kPQuarter + p], w3 = w2b_s[(k + 3) * kPQuarter + p];
            const float4 x0 = *reinterpret_cast<const float4 *>(&h1s[2 * rg][k]);
            const float4 x1 = *reinterpret_cast<const float4 *>(&h1s[2 * rg + 1][k]);
            a0 = fmaf(x0.x, w0, a0), a0 = fmaf(x0.y, w1, a0), a0 = fmaf(x0.z, w2, a0), a0 = fmaf(x0.w, w3, a0);
            a1 = fmaf(x1.x, w0, a1), a1 = fmaf(x1.y, w1, a1), a1 = fmaf(x1.z, w2, a1), a1 = fmaf(x1.w, w3, a1);
        }
        h2s[2 * rg][p] = gelu_exact(a0), h2s[2 * rg + 1][p] = gelu_exact(a1);
    }
    __syncthreads();
    {   // Linear-2c: warp w <-> row w
        const int warp = tid >> 5, lane = tid & 31;
        float s = fmaf(h2s[warp][lane], __ldg(w2c + lane), h2s[warp][32 + lane] * __ldg(w2c + 32 + lane));
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
        if (lane == 0 && warp < rows_here) out[(int64_t)b * out_bs + (int64_t)tile * kPRows + warp] = s + __ldg(b2c);
    }
}

}  // namespace sdetr

using namespace sdetr;

extern "C" int64_t sdetr_mask_predictor_level_workspace_floats(int batch, int H, int W) {
    if (batch <= 0 || H <= 0 || W <= 0) return 0;
    const int64_t HW = (int64_t)H * W, tiles = (HW + kPRows - 1) / kPRows;
    return batch * HW * kPHalf + batch * tiles * kPHalf;
}

extern "C" int sdetr_mask_predictor_level(const float *mem, int64_t mem_batch_stride, int batch, int H, int W, int channels,
                                          const float *coarse_score, int64_t coarse_batch_stride, int Hc, int Wc,
                                          const float *alpha, int alpha_index, const float *ln_gamma, const float *ln_beta,
                                          float eps, const float *w1_t, const float *b1, const float *w2a_t, const float *b2a,
                                          const float *w2b_t, const float *b2b, const float *w2c, const float *b2c,
                                          float *workspace, int64_t workspace_floats, float *out, int64_t out_batch_stride,
                                          sdetr_stream_t stream) {
    SDETR_REQUIRE(mem && ln_gamma && ln_beta && w1_t && b1 && w2a_t && b2a && w2b_t && b2b && w2c && b2c && workspace && out,
                  SDETR_ERR_INVALID_ARG, "mask_predictor_level: null pointer");
    SDETR_REQUIRE(channels == kPC, SDETR_ERR_UNSUPPORTED, "mask_predictor_level: channels=%d (only %d)", channels, kPC);
    SDETR_REQUIRE(batch > 0 && batch <= 65535 && H > 0 && W > 0, SDETR_ERR_INVALID_ARG, "mask_predictor_level: bad sizes");
    SDETR_REQUIRE(!coarse_score || (alpha && alpha_index >= 0 && Hc > 0 && Wc > 0), SDETR_ERR_INVALID_ARG,
                  "mask_predictor_level: a coarser level needs its size and alpha");
    SDETR_REQUIRE(aligned16(mem) && mem_batch_stride % 4 == 0 && aligned16(ln_gamma) && aligned16(ln_beta) && aligned16(workspace) &&
                      aligned16(w2a_t) && aligned16(w2b_t),
                  SDETR_ERR_INVALID_ARG, "mask_predictor_level: 16-byte alignment required");
    SDETR_REQUIRE(workspace_floats >= sdetr_mask_predictor_level_workspace_floats(batch, H, W), SDETR_ERR_WORKSPACE,
                  "mask_predictor_level: workspace too small");
    const int HW = H * W, tiles = (HW + kPRows - 1) / kPRows;
    float *zlocal = workspace, *partial = workspace + (int64_t)batch * HW * kPHalf;
    PredictorLevel lv{mem, mem_batch_stride, coarse_score, coarse_batch_stride, alpha, alpha_index, H, W, Hc, Wc};
    static PerDeviceOnce once;
    SDETR_OPT_IN_SMEM(once, predictor_back_kernel, kBackSmem, "mask_predictor_level");
    predictor_front_kernel<<<dim3(tiles, batch), kFrontThreads, 0, (cudaStream_t)stream>>>(lv, ln_gamma, ln_beta, eps, w1_t, b1, zlocal, partial);
    const int rc = check_launch("mask_predictor_level/front");
    if (rc != SDETR_OK) return rc;
    predictor_back_kernel<<<dim3(tiles, batch), kBackThreads, kBackSmem, (cudaStream_t)stream>>>(zlocal, partial, tiles, HW, w2a_t, b2a, w2b_t, b2b, w2c,
                                                                                   b2c, out, out_batch_stride);
    return check_launch("mask_predictor_level/back");
}
