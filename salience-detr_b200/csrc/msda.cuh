// Shared declarations of the MSDA forward kernels (msda_forward.cu, msda_forward_tma.cu).
#pragma once
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
    int ref_dim;  // fused variant: 2 = reference points, 4 = reference boxes (cx, cy, w, h) (ms_deform_attn.py:345-349)
};

// fused variant: raw offset -> sampling location.  2-d: ref + off / (W_l, H_l) (ms_deform_attn.py:339-344);
// 4-d: ref_xy + off / P * ref_wh * 0.5 (:345-349), same operation order as the reference
template <int P>
__device__ __forceinline__ void fused_location(const float *__restrict__ rrow, int l, int ref_dim, float ox, float oy, float Wf,
                                               float Hf, float &x, float &y) {
    if (ref_dim == 2) {
        const float2 r = __ldg(reinterpret_cast<const float2 *>(rrow) + l);
        x = r.x + __fdividef(ox, Wf);
        y = r.y + __fdividef(oy, Hf);
    } else {
        const float4 r = __ldg(reinterpret_cast<const float4 *>(rrow) + l);
        x = r.x + __fmul_rn(__fmul_rn(__fdiv_rn(ox, (float)P), r.z), 0.5f);
        y = r.y + __fmul_rn(__fmul_rn(__fdiv_rn(oy, (float)P), r.w), 0.5f);
    }
}


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

// TMA-staged variant (msda_forward_tma.cu): D = 32, L = 4, P = 4, head-major schedule; returns SDETR_OK or an error code
int launch_msda_tma(const MsdaFwdParams &p, bool fused, cudaStream_t stream);

}  // namespace sdetr
