// MSDA core forward for sm_100a -- sdetr_msda_forward / _ex / sdetr_msda_fused_forward.
//
// Reference semantics: models/bricks/ops/cuda/ms_deform_im2col_cuda.cuh:22-73,226-288 (kernel) and
// models/bricks/ms_deform_attn.py:322-344 (softmax + sampling-location arithmetic, fused variant).
//
// Design (not a port of the reference kernel, which maps one thread to one output scalar and issues
// 64 dependent 4-byte gathers per thread):
//   * a "group" of D/4 lanes owns one (image, query, head); every corner fetch is one 128-bit load per
//     lane, so a warp instruction moves four complete 128-byte head rows (D = 32);
//   * the 2*L*P sampling coordinates and L*P attention weights of the (query, head) are loaded ONCE,
//     coalesced, spread over the lanes of the group, and broadcast point by point with warp shuffles;
//   * in the fused variant the lanes run the L*P-way softmax with shuffle reductions and turn raw
//     offsets into sampling locations themselves, so neither tensor round-trips through HBM;
//   * the P points of a level are fetched together (4*P independent 128-bit loads in flight per lane);
//   * level shapes are read once per thread from the reference's int64 device tensors;
//   * queries may be processed in a caller-supplied (spatially tiled) order, either query-major
//     (all heads of 256/(8*M) consecutive queries per CTA) or head-major (one head of a chunk of
//     consecutive queries per CTA) so that co-resident groups gather from the same value neighbourhood
//     and hit in L1 instead of L2.  Output rows stay in the caller's order.
#include "common.cuh"

namespace sdetr {

struct MsdaFwdParams {
    const float *value;
    int64_t v_bstride, v_tstride;  // floats
    const int64_t *shapes, *lsi;
    const float *loc, *attn;  // plain variant
    const float *ref, *proj;  // fused variant
    int64_t proj_stride;
    float *loc_out, *attn_out;
    float *out;
    const int32_t *order;
    int batch, nv, heads, nq, chunk;
};

constexpr int kThreads = 256;

// ---- lane-distributed small arrays -----------------------------------------------------------------
// N floats spread over LANES lanes, PER = ceil(N/LANES) contiguous elements per lane.
template <int N, int LANES>
struct GroupArray {
    static constexpr int PER = (N + LANES - 1) / LANES;
    float r[PER];

    __device__ __forceinline__ void load(const float *base, int lane) {
        if constexpr (N % LANES == 0 && PER % 4 == 0) {
#pragma unroll
            for (int i = 0; i < PER / 4; ++i) {
                float4 v = ld_stream_f4(base + lane * PER + 4 * i);
                r[4 * i] = v.x, r[4 * i + 1] = v.y, r[4 * i + 2] = v.z, r[4 * i + 3] = v.w;
            }
        } else if constexpr (N % LANES == 0 && PER % 2 == 0) {
#pragma unroll
            for (int i = 0; i < PER / 2; ++i) {
                float2 v = __ldg(reinterpret_cast<const float2 *>(base + lane * PER + 2 * i));
                r[2 * i] = v.x, r[2 * i + 1] = v.y;
            }
        } else {
#pragma unroll
            for (int i = 0; i < PER; ++i) r[i] = (lane * PER + i < N) ? __ldg(base + lane * PER + i) : 0.f;
        }
    }
    __device__ __forceinline__ void store(float *base, int lane) const {
        if constexpr (N % LANES == 0 && PER % 4 == 0) {
#pragma unroll
            for (int i = 0; i < PER / 4; ++i)
                st_stream_f4(base + lane * PER + 4 * i, make_float4(r[4 * i], r[4 * i + 1], r[4 * i + 2], r[4 * i + 3]));
        } else if constexpr (N % LANES == 0 && PER % 2 == 0) {
#pragma unroll
            for (int i = 0; i < PER / 2; ++i)
                *reinterpret_cast<float2 *>(base + lane * PER + 2 * i) = make_float2(r[2 * i], r[2 * i + 1]);
        } else {
#pragma unroll
            for (int i = 0; i < PER; ++i)
                if (lane * PER + i < N) base[lane * PER + i] = r[i];
        }
    }
    // broadcast element E (compile time) to every lane of the group
    template <int E>
    __device__ __forceinline__ float get() const {
        return __shfl_sync(0xffffffffu, r[E % PER], E / PER, LANES);
    }
};

template <int LANES>
__device__ __forceinline__ float group_max(float v) {
#pragma unroll
    for (int o = LANES / 2; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o, LANES));
    return v;
}
template <int LANES>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
    for (int o = LANES / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o, LANES);
    return v;
}

struct LevelGeom {
    int H, W;
    float Hf, Wf;
    int64_t start;  // level_start_index
};

// one sampling point of one level: issue the (up to) four corner loads
struct Corner4 {
    float4 v00, v01, v10, v11;
    float w00, w01, w10, w11;
};

__device__ __forceinline__ void fetch_point(Corner4 &c, const float *__restrict__ lvl_base, int64_t tstride,
                                            const LevelGeom &g, float x, float y, float a) {
    const float h_im = fmaf(y, g.Hf, -0.5f), w_im = fmaf(x, g.Wf, -0.5f);  // .cuh:274-275
    const bool in = h_im > -1.f && w_im > -1.f && h_im < g.Hf && w_im < g.Wf;  // .cuh:277
    const float hf = floorf(h_im), wf = floorf(w_im);
    const int y0 = (int)hf, x0 = (int)wf;
    const float ly = h_im - hf, lx = w_im - wf, hy = 1.f - ly, hx = 1.f - lx;
    const bool top = in && y0 >= 0, bot = in && y0 + 1 <= g.H - 1;
    const bool lef = x0 >= 0, rig = x0 + 1 <= g.W - 1;
    const float *p = lvl_base + (int64_t)(y0 * g.W + x0) * tstride;
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    c.v00 = (top && lef) ? ldg_f4(p) : z;
    c.v01 = (top && rig) ? ldg_f4(p + tstride) : z;
    c.v10 = (bot && lef) ? ldg_f4(p + (int64_t)g.W * tstride) : z;
    c.v11 = (bot && rig) ? ldg_f4(p + (int64_t)(g.W + 1) * tstride) : z;
    c.w00 = a * hy * hx, c.w01 = a * hy * lx, c.w10 = a * ly * hx, c.w11 = a * ly * lx;
}

__device__ __forceinline__ void accumulate(float4 &acc, const Corner4 &c) {
    acc.x = fmaf(c.w00, c.v00.x, fmaf(c.w01, c.v01.x, fmaf(c.w10, c.v10.x, fmaf(c.w11, c.v11.x, acc.x))));
    acc.y = fmaf(c.w00, c.v00.y, fmaf(c.w01, c.v01.y, fmaf(c.w10, c.v10.y, fmaf(c.w11, c.v11.y, acc.y))));
    acc.z = fmaf(c.w00, c.v00.z, fmaf(c.w01, c.v01.z, fmaf(c.w10, c.v10.z, fmaf(c.w11, c.v11.z, acc.z))));
    acc.w = fmaf(c.w00, c.v00.w, fmaf(c.w01, c.v01.w, fmaf(c.w10, c.v10.w, fmaf(c.w11, c.v11.w, acc.w))));
}

// compile-time loops over levels / points
template <int L, int P, int LANES, int LVL = 0>
struct LevelLoop {
    template <class LocArr, class AttArr>
    static __device__ __forceinline__ void run(float4 &acc, const LocArr &loc, const AttArr &att,
                                               const LevelGeom (&geo)[L], const float *__restrict__ vbase,
                                               int64_t tstride) {
        Corner4 c[P];
        const float *lvl_base = vbase + geo[LVL].start * tstride;
        fetch_all<0>(c, loc, att, geo[LVL], lvl_base, tstride);
#pragma unroll
        for (int p = 0; p < P; ++p) accumulate(acc, c[p]);
        if constexpr (LVL + 1 < L) LevelLoop<L, P, LANES, LVL + 1>::run(acc, loc, att, geo, vbase, tstride);
    }
    template <int PT, class LocArr, class AttArr>
    static __device__ __forceinline__ void fetch_all(Corner4 (&c)[P], const LocArr &loc, const AttArr &att,
                                                     const LevelGeom &g, const float *__restrict__ lvl_base,
                                                     int64_t tstride) {
        constexpr int E = LVL * P + PT;
        const float x = loc.template get<2 * E>();
        const float y = loc.template get<2 * E + 1>();
        const float a = att.template get<E>();
        fetch_point(c[PT], lvl_base, tstride, g, x, y, a);
        if constexpr (PT + 1 < P) fetch_all<PT + 1>(c, loc, att, g, lvl_base, tstride);
    }
};

// ---- the specialised kernel -------------------------------------------------------------------------
template <int D, int L, int P, bool FUSED, bool HEAD_MAJOR>
__global__ void __launch_bounds__(kThreads, 2) msda_fwd_kernel(const MsdaFwdParams p) {
    constexpr int LANES = D / 4;
    constexpr int GROUPS = kThreads / LANES;
    constexpr int NP = L * P;
    const int lane = threadIdx.x % LANES;
    const int grp = threadIdx.x / LANES;

    LevelGeom geo[L];
#pragma unroll
    for (int l = 0; l < L; ++l) {
        geo[l].H = (int)__ldg(p.shapes + 2 * l);
        geo[l].W = (int)__ldg(p.shapes + 2 * l + 1);
        geo[l].Hf = (float)geo[l].H, geo[l].Wf = (float)geo[l].W;
        geo[l].start = __ldg(p.lsi + l);
    }

    // Work assignment.  Every lane of a warp runs the same number of iterations (the shuffles below
    // use the full mask); out-of-range groups recompute a valid item and skip the store.
    int b, m, q_first, q_last, q_step, iters;
    bool in_range = true;
    if constexpr (HEAD_MAJOR) {
        b = blockIdx.z, m = blockIdx.y;
        q_first = blockIdx.x * p.chunk + grp;
        q_last = min(p.nq, (int)(blockIdx.x + 1) * p.chunk) - 1;  // >= blockIdx.x*chunk by construction
        q_step = GROUPS;
        iters = (p.chunk + GROUPS - 1) / GROUPS;
    } else {
        int64_t item = (int64_t)blockIdx.x * GROUPS + grp;
        in_range = item < (int64_t)p.batch * p.nq * p.heads;
        if (!in_range) item = 0;
        m = (int)(item % p.heads);
        const int64_t bq = item / p.heads;
        q_first = q_last = (int)(bq % p.nq);
        b = (int)(bq / p.nq);
        q_step = 0;
        iters = 1;
    }

    for (int it = 0; it < iters; ++it) {
        int qi = q_first + it * q_step;
        const bool active = in_range && qi <= q_last;
        qi = min(qi, q_last);
        const int q = p.order ? __ldg(p.order + (int64_t)b * p.nq + qi) : qi;
        const int64_t row = (int64_t)b * p.nq + q;  // (image, query) row of every per-query tensor

        GroupArray<2 * NP, LANES> loc;
        GroupArray<NP, LANES> att;
        if constexpr (FUSED) {
            const float *prow = p.proj + row * p.proj_stride;
            loc.load(prow + (int64_t)m * 2 * NP, lane);              // raw offsets of this head
            att.load(prow + (int64_t)p.heads * 2 * NP + (int64_t)m * NP, lane);  // raw logits of this head
            // softmax over the L*P logits (ms_deform_attn.py:326-329)
            float mx = -INFINITY;
#pragma unroll
            for (int i = 0; i < att.PER; ++i)
                if (lane * att.PER + i < NP) mx = fmaxf(mx, att.r[i]);
            mx = group_max<LANES>(mx);
            float sum = 0.f;
#pragma unroll
            for (int i = 0; i < att.PER; ++i) {
                att.r[i] = (lane * att.PER + i < NP) ? expf(att.r[i] - mx) : 0.f;
                sum += att.r[i];
            }
            sum = group_sum<LANES>(sum);
#pragma unroll
            for (int i = 0; i < att.PER; ++i) att.r[i] = att.r[i] / sum;
            // loc = ref + off / (W_l, H_l) (ms_deform_attn.py:339-344)
            const float *rrow = p.ref + row * (2 * L);
#pragma unroll
            for (int i = 0; i < loc.PER; ++i) {
                const int e = lane * loc.PER + i;  // element = (l*P + pt)*2 + xy
                const int l = min(e / (2 * P), L - 1);
                const int xy = e & 1;
                float norm = xy ? geo[0].Hf : geo[0].Wf;
#pragma unroll
                for (int k = 1; k < L; ++k)
                    if (l == k) norm = xy ? geo[k].Hf : geo[k].Wf;
                loc.r[i] = __ldg(rrow + 2 * l + xy) + loc.r[i] / norm;
            }
            if (p.loc_out && active) loc.store(p.loc_out + (row * p.heads + m) * (2 * NP), lane);
            if (p.attn_out && active) att.store(p.attn_out + (row * p.heads + m) * NP, lane);
        } else {
            loc.load(p.loc + (row * p.heads + m) * (2 * NP), lane);
            att.load(p.attn + (row * p.heads + m) * NP, lane);
        }

        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        const float *vbase = p.value + (int64_t)b * p.v_bstride + (int64_t)m * D + lane * 4;
        LevelLoop<L, P, LANES>::run(acc, loc, att, geo, vbase, p.v_tstride);
        if (active) st_stream_f4(p.out + (row * p.heads + m) * D + lane * 4, acc);
    }
}

// ---- generic fallback (any head_dim % 4 == 0, L <= kMaxLevels, any P) ------------------------------------
__global__ void __launch_bounds__(kThreads) msda_fwd_generic_kernel(const MsdaFwdParams p, int D, int L, int P,
                                                                   int fused) {
    const int lanes = D / 4;
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = (int)(gid % lanes);
    const int64_t item = gid / lanes;
    const int m = (int)(item % p.heads);
    const int64_t bq = item / p.heads;
    const int qi = (int)(bq % p.nq);
    const int b = (int)(bq / p.nq);
    if (b >= p.batch) return;
    const int q = p.order ? __ldg(p.order + (int64_t)b * p.nq + qi) : qi;
    const int64_t row = (int64_t)b * p.nq + q;
    const int NP = L * P;
    const float *prow = fused ? p.proj + row * p.proj_stride : nullptr;
    float mx = -INFINITY, sum = 0.f;
    if (fused) {
        const float *lg = prow + (int64_t)p.heads * 2 * NP + (int64_t)m * NP;
        for (int i = 0; i < NP; ++i) mx = fmaxf(mx, __ldg(lg + i));
        for (int i = 0; i < NP; ++i) sum += expf(__ldg(lg + i) - mx);
    }
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    const float *vbase = p.value + (int64_t)b * p.v_bstride + (int64_t)m * D + lane * 4;
    for (int l = 0; l < L; ++l) {
        LevelGeom g;
        g.H = (int)__ldg(p.shapes + 2 * l), g.W = (int)__ldg(p.shapes + 2 * l + 1);
        g.Hf = (float)g.H, g.Wf = (float)g.W, g.start = __ldg(p.lsi + l);
        const float *lvl_base = vbase + g.start * p.v_tstride;
        for (int pt = 0; pt < P; ++pt) {
            const int e = l * P + pt;
            float x, y, a;
            if (fused) {
                const float *off = prow + (int64_t)m * 2 * NP + 2 * e;
                x = __ldg(p.ref + row * (2 * L) + 2 * l) + __ldg(off) / g.Wf;
                y = __ldg(p.ref + row * (2 * L) + 2 * l + 1) + __ldg(off + 1) / g.Hf;
                a = expf(__ldg(prow + (int64_t)p.heads * 2 * NP + (int64_t)m * NP + e) - mx) / sum;
                if (lane == 0) {
                    if (p.loc_out) {
                        p.loc_out[((row * p.heads + m) * NP + e) * 2] = x;
                        p.loc_out[((row * p.heads + m) * NP + e) * 2 + 1] = y;
                    }
                    if (p.attn_out) p.attn_out[(row * p.heads + m) * NP + e] = a;
                }
            } else {
                const float *lp = p.loc + ((row * p.heads + m) * NP + e) * 2;
                x = __ldg(lp), y = __ldg(lp + 1);
                a = __ldg(p.attn + (row * p.heads + m) * NP + e);
            }
            Corner4 c;
            fetch_point(c, lvl_base, p.v_tstride, g, x, y, a);
            accumulate(acc, c);
        }
    }
    *reinterpret_cast<float4 *>(p.out + (row * p.heads + m) * D + lane * 4) = acc;
}

// ---- host side -----------------------------------------------------------------------------------------
template <int D, int L, int P>
static void launch_special(const MsdaFwdParams &p, bool fused, int schedule, cudaStream_t s) {
    constexpr int GROUPS = kThreads / (D / 4);
    if (schedule == 1) {
        dim3 grid((p.nq + p.chunk - 1) / p.chunk, p.heads, p.batch);
        if (fused)
            msda_fwd_kernel<D, L, P, true, true><<<grid, kThreads, 0, s>>>(p);
        else
            msda_fwd_kernel<D, L, P, false, true><<<grid, kThreads, 0, s>>>(p);
    } else {
        const int64_t items = (int64_t)p.batch * p.nq * p.heads;
        dim3 grid((unsigned)((items + GROUPS - 1) / GROUPS));
        if (fused)
            msda_fwd_kernel<D, L, P, true, false><<<grid, kThreads, 0, s>>>(p);
        else
            msda_fwd_kernel<D, L, P, false, false><<<grid, kThreads, 0, s>>>(p);
    }
}

static int msda_forward_dispatch(MsdaFwdParams p, bool fused, int head_dim, int levels, int points, int schedule,
                                 cudaStream_t s) {
    SDETR_REQUIRE(p.value && p.shapes && p.lsi && p.out, SDETR_ERR_INVALID_ARG, "msda_forward: null pointer");
    SDETR_REQUIRE(fused ? (p.ref && p.proj) : (p.loc && p.attn), SDETR_ERR_INVALID_ARG,
                  "msda_forward: null sampling input");
    SDETR_REQUIRE(p.batch > 0 && p.nv > 0 && p.heads > 0 && levels > 0 && points > 0 && p.nq >= 0,
                  SDETR_ERR_INVALID_ARG, "msda_forward: non-positive size");
    SDETR_REQUIRE(head_dim % 4 == 0 && head_dim >= 4 && head_dim <= 1024, SDETR_ERR_UNSUPPORTED,
                  "msda_forward: head_dim %d must be a multiple of 4 (128-bit rows)", head_dim);
    SDETR_REQUIRE(levels <= kMaxLevels, SDETR_ERR_UNSUPPORTED, "msda_forward: more than %d levels", kMaxLevels);
    SDETR_REQUIRE(aligned16(p.value) && aligned16(p.out) && p.v_tstride % 4 == 0 && p.v_bstride % 4 == 0,
                  SDETR_ERR_INVALID_ARG, "msda_forward: value/output must be 16-byte aligned");
    SDETR_REQUIRE(schedule == 0 || schedule == 1, SDETR_ERR_INVALID_ARG, "msda_forward: bad schedule %d", schedule);
    if (p.nq == 0) return SDETR_OK;
    p.chunk = 64;
    bool special = true;
    if (fused) special = (p.proj_stride % 4 == 0) && aligned16(p.proj);
    else special = aligned16(p.loc) && aligned16(p.attn);
    if (special && head_dim == 32 && levels == 4 && points == 4)
        launch_special<32, 4, 4>(p, fused, schedule, s);
    else if (special && head_dim == 32 && levels == 5 && points == 4)
        launch_special<32, 5, 4>(p, fused, schedule, s);
    else if (special && head_dim == 64 && levels == 4 && points == 4)
        launch_special<64, 4, 4>(p, fused, schedule, s);
    else {
        const int64_t threads = (int64_t)p.batch * p.nq * p.heads * (head_dim / 4);
        msda_fwd_generic_kernel<<<(unsigned)((threads + kThreads - 1) / kThreads), kThreads, 0, s>>>(
            p, head_dim, levels, points, fused ? 1 : 0);
    }
    return check_launch("msda_forward");
}

}  // namespace sdetr

using namespace sdetr;

extern "C" int sdetr_msda_forward_ex(const float *value, int64_t value_batch_stride, int64_t value_token_stride,
                                     const int64_t *spatial_shapes, const int64_t *level_start_index,
                                     const float *sampling_loc, const float *attn_weight, float *output, int batch,
                                     int num_value, int num_heads, int head_dim, int num_levels, int num_query,
                                     int num_points, const int32_t *query_order, int schedule,
                                     sdetr_stream_t stream) {
    MsdaFwdParams p{};
    p.value = value, p.v_bstride = value_batch_stride, p.v_tstride = value_token_stride;
    p.shapes = spatial_shapes, p.lsi = level_start_index;
    p.loc = sampling_loc, p.attn = attn_weight, p.out = output, p.order = query_order;
    p.batch = batch, p.nv = num_value, p.heads = num_heads, p.nq = num_query;
    return msda_forward_dispatch(p, false, head_dim, num_levels, num_points, schedule, (cudaStream_t)stream);
}

extern "C" int sdetr_msda_forward(const float *value, const int64_t *spatial_shapes,
                                  const int64_t *level_start_index, const float *sampling_loc,
                                  const float *attn_weight, float *output, int batch, int num_value,
                                  int num_heads, int head_dim, int num_levels, int num_query, int num_points,
                                  sdetr_stream_t stream) {
    return sdetr_msda_forward_ex(value, (int64_t)num_value * num_heads * head_dim, (int64_t)num_heads * head_dim,
                                 spatial_shapes, level_start_index, sampling_loc, attn_weight, output, batch,
                                 num_value, num_heads, head_dim, num_levels, num_query, num_points, nullptr, 0,
                                 stream);
}

extern "C" int sdetr_msda_fused_forward(const float *value, int64_t value_batch_stride, int64_t value_token_stride,
                                        const int64_t *spatial_shapes, const int64_t *level_start_index,
                                        const float *ref_points, const float *proj, int64_t proj_stride,
                                        float *output, float *loc_out, float *attn_out, int batch, int num_value,
                                        int num_heads, int head_dim, int num_levels, int num_query, int num_points,
                                        const int32_t *query_order, int schedule, sdetr_stream_t stream) {
    MsdaFwdParams p{};
    p.value = value, p.v_bstride = value_batch_stride, p.v_tstride = value_token_stride;
    p.shapes = spatial_shapes, p.lsi = level_start_index;
    p.ref = ref_points, p.proj = proj, p.proj_stride = proj_stride, p.loc_out = loc_out, p.attn_out = attn_out;
    p.out = output, p.order = query_order;
    p.batch = batch, p.nv = num_value, p.heads = num_heads, p.nq = num_query;
    SDETR_REQUIRE(proj_stride >= (int64_t)num_heads * num_levels * num_points * 3, SDETR_ERR_INVALID_ARG,
                  "msda_fused_forward: proj_stride too small");
    return msda_forward_dispatch(p, true, head_dim, num_levels, num_points, schedule, (cudaStream_t)stream);
}
