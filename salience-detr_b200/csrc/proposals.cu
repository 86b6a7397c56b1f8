// Two-stage proposal selection of the decoder half -- sdetr_nms_topk_index.
//
// Reference semantics: SalienceTransformer.nms_on_topk_index (models/bricks/salience_transformer.py:249-295): the top-k
// encoder tokens (by max class logit) of every image become boxes (x-1, y-1, x+1, y+1) on their level's pixel grid,
// torchvision.ops.batched_nms(boxes, scores, idxs = level + L * image, iou_threshold) keeps a box iff no HIGHER-SCORED KEPT
// box of the same (image, level) overlaps it with IoU > threshold, and the kept token indices come back in score order.
//
// Boxes are 2x2 squares on integer centres, so only the 8 neighbouring pixels of the same level can overlap a box at all
// (|dx|,|dy| <= 1: intersection (2-|dx|)(2-|dy|), union 8 - intersection; at the reference's threshold 0.3 the 4-neighbours
// suppress, 1/3 > 0.3, the diagonal ones do not, 1/7).  That turns NMS into a neighbourhood problem on the token grid:
//   * one CTA per image; shared memory holds rank_of[token] (uint16, 0xffff = not a candidate) and a state byte per
//     candidate (unknown / kept / suppressed);
//   * greedy NMS is resolved by parallel relaxation: a candidate is SUPPRESSED as soon as one higher-ranked overlapping
//     neighbour is kept, KEPT as soon as all its higher-ranked overlapping neighbours are suppressed; rank 0 resolves in
//     round one and every round resolves at least the lowest unresolved rank, so the fixed point is exactly the sequential
//     greedy result (dependency chains on real score maps are a handful of rounds);
//   * an ordered block scan compacts the kept token indices in rank (= descending score) order.
// Input ranks are the caller's top-k order (sdetr_topk_desc: score descending, ties by smaller index).
#include "common.cuh"

namespace sdetr {

struct NmsLevels {
    int start[kMaxLevels], W[kMaxLevels], H[kMaxLevels];
    int L;
};

constexpr int kNmsThreads = 1024;

__global__ void __launch_bounds__(kNmsThreads) nms_topk_index_kernel(const int64_t *__restrict__ topk_index, int k, int nv,
                                                                     NmsLevels lv, float iou_threshold,
                                                                     int64_t *__restrict__ kept_index, int32_t *__restrict__ kept_count,
                                                                     uint8_t *__restrict__ keep_flag) {
    extern __shared__ __align__(16) uint8_t nms_smem[];
    uint16_t *rank_of = reinterpret_cast<uint16_t *>(nms_smem);         // nv entries
    uint8_t *state = nms_smem + (((size_t)nv * 2 + 15) & ~(size_t)15);   // k entries: 0 unknown, 1 kept, 2 suppressed
    __shared__ int s_changed, s_unknown, s_warp_sum[kNmsThreads / 32], s_base;
    const int b = blockIdx.x, tid = threadIdx.x;
    const int64_t *idx = topk_index + (int64_t)b * k;
    for (int t = tid; t < nv; t += kNmsThreads) rank_of[t] = 0xffffu;
    for (int i = tid; i < k; i += kNmsThreads) state[i] = 0;
    __syncthreads();
    for (int i = tid; i < k; i += kNmsThreads) rank_of[(int)idx[i]] = (uint16_t)i;
    __syncthreads();
    // overlap test for the 8 neighbour offsets: IoU = inter / (8 - inter) > threshold (torchvision: strict '>')
    const bool sup_edge = __fdiv_rn(2.f, 6.f) > iou_threshold, sup_diag = __fdiv_rn(1.f, 7.f) > iou_threshold;
    for (;;) {
        if (tid == 0) s_changed = 0, s_unknown = 0;
        __syncthreads();
        int changed = 0, unknown = 0;
        for (int i = tid; i < k; i += kNmsThreads) {
            if (state[i] != 0) continue;
            const int t = (int)idx[i];
            int l = 0;
#pragma unroll
            for (int u = 1; u < kMaxLevels; ++u)
                if (u < lv.L && t >= lv.start[u]) l = u;
            const int W = lv.W[l], H = lv.H[l], r = t - lv.start[l], y = r / W, x = r - y * W;
            bool any_kept = false, any_unknown = false;
#pragma unroll
            for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
                for (int dx = -1; dx <= 1; ++dx) {
                    if ((dx == 0 && dy == 0) || !((dx == 0 || dy == 0) ? sup_edge : sup_diag)) continue;
                    const int xx = x + dx, yy = y + dy;
                    if (xx < 0 || yy < 0 || xx >= W || yy >= H) continue;
                    const unsigned rk = rank_of[lv.start[l] + yy * W + xx];
                    if (rk >= (unsigned)i) continue;  // not a candidate (0xffff) or lower-scored
                    const uint8_t s = state[rk];      // racy read of a monotone state: a stale 0 only delays a round
                    any_kept |= s == 1;
                    any_unknown |= s == 0;
                }
            if (any_kept) state[i] = 2, changed = 1;
            else if (!any_unknown) state[i] = 1, changed = 1;
            else unknown = 1;
        }
        if (changed) s_changed = 1;
        if (unknown) s_unknown = 1;
        __syncthreads();
        const bool done = !s_unknown || !s_changed;  // !changed with unknowns left cannot happen (lowest unknown rank resolves)
        __syncthreads();
        if (done) break;
    }
    // ordered compaction of the kept candidates
    if (tid == 0) s_base = 0;
    __syncthreads();
    for (int i0 = 0; i0 < k; i0 += kNmsThreads) {
        const int i = i0 + tid;
        const int kept = (i < k && state[i] == 1) ? 1 : 0;
        const unsigned ballot = __ballot_sync(0xffffffffu, kept);
        const int lane = tid & 31, warp = tid >> 5;
        if (lane == 0) s_warp_sum[warp] = __popc(ballot);
        __syncthreads();
        int before = s_base;
        for (int w = 0; w < warp; ++w) before += s_warp_sum[w];
        const int pos = before + __popc(ballot & ((1u << lane) - 1u));
        if (kept) kept_index[(int64_t)b * k + pos] = idx[i];
        if (i < k && keep_flag) keep_flag[(int64_t)b * k + i] = (uint8_t)kept;
        __syncthreads();
        if (tid == 0) {
            int tot = 0;
            for (int w = 0; w < kNmsThreads / 32; ++w) tot += s_warp_sum[w];
            s_base += tot;
        }
        __syncthreads();
    }
    if (tid == 0) kept_count[b] = s_base;
}

}  // namespace sdetr

using namespace sdetr;

extern "C" int sdetr_nms_topk_index(const int64_t *topk_index, int batch, int k, int num_value, int num_levels,
                                    const int32_t *level_h_host, const int32_t *level_w_host, float iou_threshold,
                                    int64_t *kept_index, int32_t *kept_count, uint8_t *keep_flag, sdetr_stream_t stream) {
    SDETR_REQUIRE(topk_index && level_h_host && level_w_host && kept_index && kept_count, SDETR_ERR_INVALID_ARG,
                  "nms_topk_index: null pointer");
    SDETR_REQUIRE(batch > 0 && k > 0 && k < 65535 && num_levels > 0 && num_levels <= kMaxLevels, SDETR_ERR_INVALID_ARG,
                  "nms_topk_index: bad sizes (k must be < 65535)");
    NmsLevels lv{};
    lv.L = num_levels;
    int nv = 0;
    for (int l = 0; l < num_levels; ++l) {
        SDETR_REQUIRE(level_h_host[l] > 0 && level_w_host[l] > 0, SDETR_ERR_INVALID_ARG, "nms_topk_index: level %d is empty", l);
        lv.start[l] = nv, lv.W[l] = level_w_host[l], lv.H[l] = level_h_host[l];
        nv += level_h_host[l] * level_w_host[l];
    }
    SDETR_REQUIRE(nv == num_value && k <= nv, SDETR_ERR_INVALID_ARG, "nms_topk_index: levels hold %d tokens, num_value %d, k %d", nv,
                  num_value, k);
    const size_t smem = (((size_t)nv * 2 + 15) & ~(size_t)15) + (size_t)k;
    SDETR_REQUIRE(smem <= 200 * 1024, SDETR_ERR_UNSUPPORTED, "nms_topk_index: %d tokens do not fit the shared-memory rank table", nv);
    static PerDeviceOnce once;
    SDETR_OPT_IN_SMEM(once, nms_topk_index_kernel, 200 * 1024, "nms_topk_index");
    nms_topk_index_kernel<<<batch, kNmsThreads, smem, (cudaStream_t)stream>>>(topk_index, k, nv, lv, iou_threshold, kept_index,
                                                                            kept_count, keep_flag);
    return check_launch("nms_topk_index");
}
