// Everything the path derives from the padding masks, on the device -- sdetr_mask_plan, sdetr_sine_pos_tokens.
//
// Reference semantics restated by these kernels:
//   * valid-token counts and focus budgets: salience_transformer.py:116-121 (`valid_token_nums`, `focus_token_nums =
//     (valid * level_filter_ratio).int()`: an fp32 multiply followed by truncation);
//   * valid ratios: base_transformer.py:48-56 (first column / first row of each level's mask);
//   * proposal keep mask of gen_encoder_output_proposals: base_transformer.py:74-110 (token kept iff not padded and the
//     proposal centre and its 0.05 * 2^level size lie in (0.01, 0.99));
//   * PositionEmbeddingSine(normalize=True): position_encoding.py:48-65 (cumulative sums of the valid mask along y / x,
//     normalised by the last row / column, divided by temperature^(2 (j//2) / num_pos_feats), sin on even / cos on odd j,
//     channels = [pos_y | pos_x]).
// All arithmetic is fp32 with the reference's operation order and explicit round-to-nearest intrinsics (no FMA
// contraction), so the position embedding equals torch's op chain bit for bit wherever sinf/cosf agree with torch's.
#include "common.cuh"

namespace sdetr {

struct PlanLevels {
    int H[kMaxLevels], W[kMaxLevels], start[kMaxLevels];
    float ratio[kMaxLevels];  // level_filter_ratio
    int L;
};

// one CTA per (level, image): column scans (threads over x), row scans (threads over y), counts
__global__ void __launch_bounds__(256) mask_plan_kernel(const uint8_t *__restrict__ mask, int nv, PlanLevels lv, float offset,
                                                        float eps, float scale, float *__restrict__ ynorm,
                                                        float *__restrict__ xnorm, int32_t *__restrict__ valid,
                                                        int32_t *__restrict__ focus, float *__restrict__ valid_ratios) {
    const int l = blockIdx.x, b = blockIdx.y, L = lv.L;
    const int H = lv.H[l], W = lv.W[l];
    const uint8_t *m = mask + (int64_t)b * nv + lv.start[l];
    float *yo = ynorm + (int64_t)b * nv + lv.start[l], *xo = xnorm + (int64_t)b * nv + lv.start[l];
    __shared__ int s_count;
    if (threadIdx.x == 0) s_count = 0;
    __syncthreads();
    int local = 0;
    for (int x = threadIdx.x; x < W; x += blockDim.x) {  // cumsum along y (position_encoding.py:51)
        int total = 0;
        for (int y = 0; y < H; ++y) total += m[y * W + x] ? 0 : 1;
        local += total;
        const float denom = __fadd_rn((float)total, eps);
        int run = 0;
        for (int y = 0; y < H; ++y) {
            run += m[y * W + x] ? 0 : 1;
            yo[y * W + x] = __fmul_rn(__fdiv_rn(__fadd_rn((float)run, offset), denom), scale);
        }
    }
    for (int y = threadIdx.x; y < H; y += blockDim.x) {  // cumsum along x (:52)
        int total = 0;
        for (int x = 0; x < W; ++x) total += m[y * W + x] ? 0 : 1;
        const float denom = __fadd_rn((float)total, eps);
        int run = 0;
        for (int x = 0; x < W; ++x) {
            run += m[y * W + x] ? 0 : 1;
            xo[y * W + x] = __fmul_rn(__fdiv_rn(__fadd_rn((float)run, offset), denom), scale);
        }
    }
    atomicAdd(&s_count, local);
    __syncthreads();
    if (threadIdx.x == 0) {
        int vh = 0, vw = 0;  // base_transformer.py:50-51: valid rows of the first column, valid columns of the first row
        for (int y = 0; y < H; ++y) vh += m[y * W] ? 0 : 1;
        for (int x = 0; x < W; ++x) vw += m[x] ? 0 : 1;
        valid[b * L + l] = s_count;
        focus[b * L + l] = (int32_t)__fmul_rn((float)s_count, lv.ratio[l]);  // int64 * fp32 buffer -> fp32, .int() truncates
        valid_ratios[(b * L + l) * 2 + 0] = __fdiv_rn((float)vw, (float)W);
        valid_ratios[(b * L + l) * 2 + 1] = __fdiv_rn((float)vh, (float)H);
    }
}

// keep[b,t] = !mask & proposal inside (0.01, 0.99)  (base_transformer.py:84-108; the grid is normalised by the VALID size)
__global__ void __launch_bounds__(256) keep_kernel(const uint8_t *__restrict__ mask, int nv, int batch, PlanLevels lv,
                                                   const float *__restrict__ valid_ratios, float *__restrict__ keep) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)batch * nv) return;
    const int b = (int)(i / nv), t = (int)(i - (int64_t)b * nv);
    int l = 0;
#pragma unroll
    for (int u = 1; u < kMaxLevels; ++u)
        if (u < lv.L && t >= lv.start[u]) l = u;
    const int W = lv.W[l], H = lv.H[l], r = t - lv.start[l], y = r / W, x = r - y * W;
    // valid_H / valid_W as the reference counts them; the ratios were stored as vw / W and vh / H: recover the counts
    const float vw = rintf(valid_ratios[(b * lv.L + l) * 2 + 0] * (float)W), vh = rintf(valid_ratios[(b * lv.L + l) * 2 + 1] * (float)H);
    const float gy = __fdiv_rn((float)y + 0.5f, vh), gx = __fdiv_rn((float)x + 0.5f, vw);
    const float wh = 0.05f * exp2f((float)l);
    const bool ok = gy > 0.01f && gy < 0.99f && gx > 0.01f && gx < 0.99f && wh > 0.01f && wh < 0.99f;
    keep[i] = (ok && !mask[i]) ? 1.f : 0.f;
}

// pos[b,t,:] = [sin/cos(ynorm / dim_ty) | sin/cos(xnorm / dim_tx)], one warp per token row, 128-bit stores
__global__ void __launch_bounds__(256) sine_pos_tokens_kernel(const float *__restrict__ ynorm, const float *__restrict__ xnorm,
                                                              const float *__restrict__ dim_ty, const float *__restrict__ dim_tx,
                                                              int64_t rows, int npf, float *__restrict__ pos) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const float y = __ldg(ynorm + row), x = __ldg(xnorm + row);
    float *out = pos + row * (int64_t)(2 * npf);
    for (int c = lane * 4; c < 2 * npf; c += 128) {
        const bool is_y = c < npf;  // npf % 4 == 0: a 4-channel group never straddles the y | x boundary
        const float e = is_y ? y : x;
        const float *dt = is_y ? dim_ty + c : dim_tx + (c - npf);
        float4 v;
        v.x = sinf(__fdiv_rn(e, __ldg(dt + 0)));
        v.y = cosf(__fdiv_rn(e, __ldg(dt + 1)));
        v.z = sinf(__fdiv_rn(e, __ldg(dt + 2)));
        v.w = cosf(__fdiv_rn(e, __ldg(dt + 3)));
        st_stream_f4(out + c, v);
    }
}

}  // namespace sdetr

using namespace sdetr;

extern "C" int sdetr_mask_plan(const uint8_t *mask, int batch, int num_value, int num_levels, const int32_t *level_h_host,
                               const int32_t *level_w_host, const float *level_filter_ratio_host, float pos_offset,
                               float pos_eps, float pos_scale, float *ynorm, float *xnorm, int32_t *valid_token_nums,
                               int32_t *focus_token_nums, float *valid_ratios, float *keep, sdetr_stream_t stream) {
    SDETR_REQUIRE(mask && level_h_host && level_w_host && level_filter_ratio_host && ynorm && xnorm && valid_token_nums &&
                      focus_token_nums && valid_ratios && keep,
                  SDETR_ERR_INVALID_ARG, "mask_plan: null pointer");
    SDETR_REQUIRE(batch > 0 && batch <= 65535 && num_levels > 0 && num_levels <= kMaxLevels, SDETR_ERR_INVALID_ARG,
                  "mask_plan: bad batch / level count");
    PlanLevels lv{};
    lv.L = num_levels;
    int nv = 0;
    for (int l = 0; l < num_levels; ++l) {
        SDETR_REQUIRE(level_h_host[l] > 0 && level_w_host[l] > 0, SDETR_ERR_INVALID_ARG, "mask_plan: level %d is empty", l);
        lv.H[l] = level_h_host[l], lv.W[l] = level_w_host[l], lv.start[l] = nv, lv.ratio[l] = level_filter_ratio_host[l];
        nv += level_h_host[l] * level_w_host[l];
    }
    SDETR_REQUIRE(nv == num_value, SDETR_ERR_INVALID_ARG, "mask_plan: levels hold %d tokens, num_value is %d", nv, num_value);
    cudaStream_t s = (cudaStream_t)stream;
    mask_plan_kernel<<<dim3(num_levels, batch), 256, 0, s>>>(mask, nv, lv, pos_offset, pos_eps, pos_scale, ynorm, xnorm,
                                                           valid_token_nums, focus_token_nums, valid_ratios);
    int rc;
    if ((rc = check_launch("mask_plan/scan"))) return rc;
    const int64_t n = (int64_t)batch * nv;
    keep_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(mask, nv, batch, lv, valid_ratios, keep);
    return check_launch("mask_plan/keep");
}

extern "C" int sdetr_sine_pos_tokens(const float *ynorm, const float *xnorm, const float *dim_ty, const float *dim_tx,
                                     int64_t rows, int num_pos_feats, float *pos_tokens, sdetr_stream_t stream) {
    SDETR_REQUIRE(ynorm && xnorm && dim_ty && dim_tx && pos_tokens, SDETR_ERR_INVALID_ARG, "sine_pos_tokens: null pointer");
    SDETR_REQUIRE(num_pos_feats > 0 && num_pos_feats % 4 == 0 && aligned16(pos_tokens), SDETR_ERR_INVALID_ARG,
                  "sine_pos_tokens: num_pos_feats must be a positive multiple of 4, output 16-byte aligned");
    if (rows <= 0) return SDETR_OK;
    sine_pos_tokens_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, (cudaStream_t)stream>>>(ynorm, xnorm, dim_ty, dim_tx, rows,
                                                                                        num_pos_feats, pos_tokens);
    return check_launch("sine_pos_tokens");
}

// ---- salience supervision targets (training side of the filter) -- sdetr_salience_targets -------------------------------------
// Reference semantics: SalienceCriterion.get_pixel_coordinate / get_mask_single_level with noise_scale = 0
// (models/detectors/salience_detr.py:64-114): for every token (pixel centre ((x+.5) s_x, (y+.5) s_y) of its level) and every
// ground-truth box, the four border distances decide (a) inside the box: min > 0, (b) the box belongs to this level:
// lo < max <= hi; the target is max over the boxes the pixel is inside of  1 - sqrt(dx^2 + dy^2) / 2  with
// dx = (l - r) / (l + r), dy = (t - b) / (t + b), and 0 unless some box satisfies (a) and (b).  The reference materialises
// (HW, boxes, 4) tensors per image and level (~15 ATen launches each); here it is one thread per token, one launch.
namespace sdetr {
struct TargetLevels {
    int H[kMaxLevels], W[kMaxLevels], start[kMaxLevels];
    float sy[kMaxLevels], sx[kMaxLevels], lo[kMaxLevels], hi[kMaxLevels];
    int L;
};

__global__ void __launch_bounds__(256) salience_targets_kernel(const float *__restrict__ boxes /* (b, max_boxes, 4) xyxy px */,
                                                               const int32_t *__restrict__ num_boxes, int max_boxes, int batch,
                                                               int nv, TargetLevels lv, float *__restrict__ target) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)batch * nv) return;
    const int b = (int)(i / nv), t = (int)(i - (int64_t)b * nv);
    int l = 0;
#pragma unroll
    for (int u = 1; u < kMaxLevels; ++u)
        if (u < lv.L && t >= lv.start[u]) l = u;
    const int r = t - lv.start[l], y = r / lv.W[l], x = r - y * lv.W[l];
    const float cx = __fmul_rn((float)x + 0.5f, lv.sx[l]), cy = __fmul_rn((float)y + 0.5f, lv.sy[l]);
    const float lo = lv.lo[l], hi = lv.hi[l];
    float best = 0.f;
    bool pos = false;
    const float *bx = boxes + (int64_t)b * max_boxes * 4;
    const int m = __ldg(num_boxes + b);
    for (int j = 0; j < m; ++j) {
        const float4 g = __ldg(reinterpret_cast<const float4 *>(bx) + j);
        const float dl = __fsub_rn(cx, g.x), dt = __fsub_rn(cy, g.y), dr = __fsub_rn(g.z, cx), db = __fsub_rn(g.w, cy);
        const float mn = fminf(fminf(dl, dt), fminf(dr, db)), mx = fmaxf(fmaxf(dl, dt), fmaxf(dr, db));
        const bool inside = mn > 0.f;
        pos |= inside && mx > lo && mx <= hi;
        if (inside) {
            const float ddx = __fdiv_rn(__fsub_rn(dl, dr), __fadd_rn(dl, dr)), ddy = __fdiv_rn(__fsub_rn(dt, db), __fadd_rn(dt, db));
            const float conf = __fsub_rn(1.f, __fdiv_rn(__fsqrt_rn(__fadd_rn(__fmul_rn(ddx, ddx), __fmul_rn(ddy, ddy))), 2.f));
            best = fmaxf(best, conf);
        }
    }
    target[i] = pos ? best : 0.f;
}
}  // namespace sdetr

extern "C" int sdetr_salience_targets(const float *boxes_xyxy, const int32_t *num_boxes, int max_boxes, int batch, int num_value,
                                      int num_levels, const int32_t *level_h_host, const int32_t *level_w_host,
                                      const float *stride_y_host, const float *stride_x_host, const float *limit_lo_host,
                                      const float *limit_hi_host, float *target, sdetr_stream_t stream) {
    SDETR_REQUIRE(num_boxes && level_h_host && level_w_host && stride_y_host && stride_x_host && limit_lo_host && limit_hi_host && target,
                  SDETR_ERR_INVALID_ARG, "salience_targets: null pointer");
    SDETR_REQUIRE(batch > 0 && max_boxes >= 0 && (max_boxes == 0 || (boxes_xyxy && aligned16(boxes_xyxy))) && num_levels > 0 &&
                      num_levels <= kMaxLevels,
                  SDETR_ERR_INVALID_ARG, "salience_targets: bad sizes / alignment");
    TargetLevels lv{};
    lv.L = num_levels;
    int nv = 0;
    for (int l = 0; l < num_levels; ++l) {
        lv.H[l] = level_h_host[l], lv.W[l] = level_w_host[l], lv.start[l] = nv;
        lv.sy[l] = stride_y_host[l], lv.sx[l] = stride_x_host[l], lv.lo[l] = limit_lo_host[l], lv.hi[l] = limit_hi_host[l];
        SDETR_REQUIRE(lv.H[l] > 0 && lv.W[l] > 0, SDETR_ERR_INVALID_ARG, "salience_targets: level %d is empty", l);
        nv += lv.H[l] * lv.W[l];
    }
    SDETR_REQUIRE(nv == num_value, SDETR_ERR_INVALID_ARG, "salience_targets: levels hold %d tokens, num_value is %d", nv, num_value);
    const int64_t n = (int64_t)batch * nv;
    salience_targets_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(boxes_xyxy, num_boxes, max_boxes, batch, nv,
                                                                                         lv, target);
    return check_launch("salience_targets");
}
