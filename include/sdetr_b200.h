/*
 * sdetr_b200.h -- C-ABI of the B200-native (sm_100a) Salience-DETR encoder hot path.
 *
 * This is the drop-in boundary: plain pointers and sizes, no torch types.  All pointers are DEVICE
 * pointers unless the parameter name ends in `_host`.  Every call is asynchronous on `stream`
 * (a cudaStream_t passed as void*) of the CURRENT device, allocates nothing, never synchronises, and is
 * re-entrant across streams, threads and devices (per-device kernel attributes are set on first use per device).
 * The only process-global state are the benchmarking knobs sdetr_set_option / sdetr_gemm_set_variant /
 * sdetr_gemm_set_trace (atomics: a call reads them once at launch).  Return value: 0 on success, negative sdetr_status on error (message: sdetr_last_error(),
 * thread-local).  Launch errors are RETURNED (the reference only printf's them:
 * models/bricks/ops/cuda/ms_deform_im2col_cuda.cuh:937-941, 1310-1314).
 *
 * Citations below are relative to the reference root (xiuqhou/Salience-DETR @ 6262e05).
 * "(b,Nq,M,L,P,2)" etc. are row-major contiguous shapes; fp32 unless stated.
 */
#ifndef SDETR_B200_H
#define SDETR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void *sdetr_stream_t; /* cudaStream_t */

enum sdetr_status {
    SDETR_OK = 0,
    SDETR_ERR_INVALID_ARG = -1, /* null pointer, non-positive size, misaligned pointer */
    SDETR_ERR_UNSUPPORTED = -2, /* shape outside the compiled kernel family */
    SDETR_ERR_CUDA = -3,        /* cudaGetLastError() after a launch */
    SDETR_ERR_WORKSPACE = -4    /* workspace too small */
};

/* library version (major*10000 + minor*100 + patch) and last error string of the calling thread */
int sdetr_version(void);
const char *sdetr_last_error(void);
/* number of kernels this library has launched in the calling process (bench.py "gpu_launches") */
unsigned long long sdetr_launch_count(void);
/* The persistent tensor-core kernels (sdetr_gemm_f16x3_pre, sdetr_ffn_fused_layernorm) launch one CTA per SM and leave no room
 * for anything else on it; n > 0 caps them at n CTAs so that concurrently running streams (other lanes of a pipeline) find free
 * SMs for their small kernels.  0 (default) = every SM.  Process-wide. */
int sdetr_set_persistent_ctas(int n);

/* process-wide tuning knobs (benchmark sweeps): "msda_min_blocks" (2..4 resident CTAs/SM the specialised
 * MSDA kernel is compiled for), "msda_chunk" (queries per CTA in the head-major schedule). */
int sdetr_set_option(const char *name, int value);

/* ------------------------------------------------------------------------------------------------
 * MSDA core forward.  Replaces `_C.ms_deform_attn_forward`
 *   (models/bricks/ops/cuda/ms_deform_attn_cuda.cu:12-72, kernel ms_deform_im2col_cuda.cuh:226-288;
 *    numerically the same op as multi_scale_deformable_attn_pytorch, models/bricks/ms_deform_attn.py:159-212)
 *   out[b,q,m,:] = sum_{l,p} attn[b,q,m,l,p] * bilinear_zero_pad(value_l[b,:,m,:], loc[b,q,m,l,p])
 *   pixel coords  x = loc_x*W_l - 0.5, y = loc_y*H_l - 0.5 (align_corners=False).
 * value (b,Nv,M,D); spatial_shapes (L,2) int64 (H,W); level_start_index (L) int64 -- both on the device,
 * exactly the tensors the reference op receives; sampling_loc (b,Nq,M,L,P,2); attn_weight (b,Nq,M,L,P);
 * output (b,Nq,M*D), fully overwritten.  The reference's im2col_step batching (.cu:42-66) has no
 * numerical effect and is not needed (one launch covers the whole batch).
 */
int sdetr_msda_forward(const float *value, const int64_t *spatial_shapes, const int64_t *level_start_index,
                       const float *sampling_loc, const float *attn_weight, float *output, int batch,
                       int num_value, int num_heads, int head_dim, int num_levels, int num_query,
                       int num_points, sdetr_stream_t stream);

/* Extended form used by the encoder mirror:
 *  - value_batch_stride / value_token_stride (in floats) let `value` be a column slice of a wider
 *    projection buffer (the six layers' value_proj outputs share one GEMM; salience_transformer.py:452
 *    feeds the SAME value tokens to every layer);
 *  - query_order (b,Nq) int32, nullable: permutation giving the PROCESSING order of the queries
 *    (spatially tiled for L1 locality; results are written to their original rows);
 *  - schedule: 0 = query-major (4 queries x all heads per CTA), 1 = head-major chunks
 *    (one head of `chunk` consecutive queries of the processing order per CTA). */
int sdetr_msda_forward_ex(const float *value, int64_t value_batch_stride, int64_t value_token_stride,
                          const int64_t *spatial_shapes, const int64_t *level_start_index,
                          const float *sampling_loc, const float *attn_weight, float *output, int batch,
                          int num_value, int num_heads, int head_dim, int num_levels, int num_query,
                          int num_points, const int32_t *query_order, int schedule, sdetr_stream_t stream);

/* Fused attention-weight softmax + sampling-location arithmetic + MSDA core
 *   (models/bricks/ms_deform_attn.py:322-344 fused into the core, 2-d reference points):
 *   proj (b,Nq,proj_stride): per query, M*L*P*2 raw sampling offsets followed by M*L*P raw attention
 *   logits (the concatenated sampling_offsets | attention_weights Linear output);
 *   ref_points (b,Nq,L,2);  loc = ref + off / (W_l,H_l);  attn = softmax over the L*P logits of a head.
 *   loc_out / attn_out: nullable; when given they receive sampling_locations / attention_weights in the
 *   reference layouts (needed by the backward).  Other arguments as sdetr_msda_forward_ex. */
int sdetr_msda_fused_forward(const float *value, int64_t value_batch_stride, int64_t value_token_stride,
                             const int64_t *spatial_shapes, const int64_t *level_start_index,
                             const float *ref_points, const float *proj, int64_t proj_stride, float *output,
                             float *loc_out, float *attn_out, int batch, int num_value, int num_heads,
                             int head_dim, int num_levels, int num_query, int num_points,
                             const int32_t *query_order, int schedule, sdetr_stream_t stream);

/* Same with 4-d reference BOXES (cx, cy, w, h), the decoder's cross-attention (ms_deform_attn.py:345-349):
 * ref_boxes (b,Nq,L,4), 16-byte aligned;  loc = ref_xy + off / P * ref_wh * 0.5  (the reference's operation order). */
int sdetr_msda_fused_forward_boxes(const float *value, int64_t value_batch_stride, int64_t value_token_stride,
                                   const int64_t *spatial_shapes, const int64_t *level_start_index,
                                   const float *ref_boxes, const float *proj, int64_t proj_stride, float *output,
                                   float *loc_out, float *attn_out, int batch, int num_value, int num_heads, int head_dim,
                                   int num_levels, int num_query, int num_points, const int32_t *query_order, int schedule,
                                   sdetr_stream_t stream);

/* Benchmarking variant "msda_tma" (sdetr_set_option("msda_tma", 1); D = 32, L = 4, P = 4, head-major schedule with a query
 * order): the fused / plain forward stages per-level value windows in shared memory with TMA (csrc/msda_forward_tma.cu).
 * TMA tensor maps are encoded on the host, which needs the level shapes as HOST integers: give them here once per
 * geometry (process-global; the default path reads the shapes from the device tensors and needs none of this). */
int sdetr_msda_set_host_shapes(int num_levels, const int32_t *level_h_host, const int32_t *level_w_host);

/* Two-stage proposal selection, NMS on token indices (SalienceTransformer.nms_on_topk_index, salience_transformer.py:
 * 249-295 = torchvision.ops.batched_nms over boxes (x-1,y-1,x+1,y+1) on each (image, level) grid).  topk_index (b,k) int64:
 * candidate tokens in descending score order (the rank decides who suppresses whom).  kept_index (b,k) int64 receives the
 * surviving tokens in the same order (first kept_count[b] entries valid), keep_flag (b,k) uint8 (nullable) the per-
 * candidate verdict.  k < 65535; the rank table (2 bytes per token) lives in shared memory: num_value up to ~98 000. */
int sdetr_nms_topk_index(const int64_t *topk_index, int batch, int k, int num_value, int num_levels,
                         const int32_t *level_h_host, const int32_t *level_w_host, float iou_threshold, int64_t *kept_index,
                         int32_t *kept_count, uint8_t *keep_flag, sdetr_stream_t stream);


/* MSDA core backward.  Replaces `_C.ms_deform_attn_backward`
 *   (ms_deform_attn_cuda.cu:75-145; kernels ms_deform_im2col_cuda.cuh:76-148, 290-392).
 * grad_output (b,Nq,M*D) -> grad_value (b,Nv,M,D), grad_sampling_loc, grad_attn_weight (shapes of the
 * inputs).  All three outputs are fully written (grad_value is zeroed on `stream` first; the reference
 * allocates zeros, .cu:113-115).  grad_value accumulates with fp32 red.global adds like the reference's
 * atomicAdd (run-to-run order of additions is not deterministic). */
int sdetr_msda_backward(const float *value, const int64_t *spatial_shapes, const int64_t *level_start_index,
                        const float *sampling_loc, const float *attn_weight, const float *grad_output,
                        float *grad_value, float *grad_sampling_loc, float *grad_attn_weight, int batch,
                        int num_value, int num_heads, int head_dim, int num_levels, int num_query,
                        int num_points, sdetr_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Hierarchical salience token filter (models/bricks/salience_transformer.py:146-168), one call:
 *   per level l: v = where(mask, min over the whole (b,HW_l) score tensor, score)       (:146)
 *                top-k_l of v per image, indices offset by level_start                  (:150-151)
 *   concatenate levels, sort by score descending, reorder indices                       (:156-158)
 *   foreground_score = where(mask, min over the whole (b,Nv) tensor, score)             (:166-168)
 * Order of equal scores (unspecified in torch.topk/sort): larger score first, then smaller token index.
 * raw_score (b,Nv); mask (b,Nv) uint8, 1 = padding; level_start_host / level_size_host / level_k_host:
 * HOST int32 arrays of length num_levels (k_l = level_token_nums, :120; k_l <= size_l).
 * -> selected_inds (b,K) int64, selected_score (b,K), foreground_score (b,Nv), K = sum k_l.
 * tile_order (b,K) int32, nullable: positions 0..K-1 of each image's selected list ordered by the
 * spatial cell of their token (cell_px x cell_px image pixels; level-l token = level_stride_host[l] px):
 * the processing order handed to sdetr_msda_forward_ex after sdetr_order_prefixes().
 * workspace: sdetr_salience_select_workspace() bytes, 256-byte aligned.
 */
size_t sdetr_salience_select_workspace(int batch, int num_value, int num_levels);
int sdetr_salience_select(const float *raw_score, const uint8_t *mask, const int32_t *level_start_host,
                          const int32_t *level_size_host, const int32_t *level_k_host,
                          const int32_t *level_width_host, const int32_t *level_stride_host, int cell_px,
                          int batch, int num_value, int num_levels, int64_t *selected_inds,
                          float *selected_score, float *foreground_score, int32_t *tile_order,
                          void *workspace, size_t workspace_bytes, sdetr_stream_t stream);

/* Per-layer prefixes of the processing order: for each layer j (prefix length nq_host[j] of the selected
 * list, salience_transformer.py:161-165) emit the positions < nq_j in tile order:
 * out_orders + order_offset_host[j] : (b, nq_j) int32. */
int sdetr_order_prefixes(const int32_t *tile_order, int batch, int K, int num_layers,
                         const int32_t *nq_host, const int64_t *order_offset_host, int32_t *out_orders,
                         sdetr_stream_t stream);

/* Generic segmented top-k by score (descending; ties: smaller position first).  Used for the
 * "top-300 most salient tokens" pre-attention selection (salience_transformer.py:366-367).
 * score (segments, n) -> topk_index (segments, k) int64 positions within the segment.
 * workspace: sdetr_topk_workspace() bytes. */
size_t sdetr_topk_workspace(int segments, int n);
int sdetr_topk_desc(const float *score, int segments, int n, int k, int64_t *topk_index, void *workspace,
                    size_t workspace_bytes, sdetr_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Token movement around the encoder layers (models/bricks/salience_transformer.py:454-495).
 */

/* Fused four-way gather (:454-461).  inds: int64, row stride inds_stride (a prefix view of selected_inds).
 * tokens/pos (b,Nv,C), fg (b,Nv), valid_ratios (b,L,2) ->
 * query (b,Nq,C), query_pos (b,Nq,C), fg_q (b,Nq), ref_q (b,Nq,L,2) where ref_q is the gathered row of
 * get_reference_points (:417-432), recomputed from the token's (level,y,x) with identical fp32 arithmetic.
 * query_sum (b,Nq,C), may be NULL: query + query_pos, the `with_pos_embed` input of the layer's attention (:381). */
int sdetr_token_gather(const float *tokens, const float *pos, const float *fg, const float *valid_ratios,
                       const int64_t *inds, int64_t inds_stride, const int64_t *spatial_shapes,
                       const int64_t *level_start_index, int batch, int num_value, int channels,
                       int num_levels, int num_query, float *query, float *query_pos, float *fg_q,
                       float *ref_q, float *query_sum, sdetr_stream_t stream);

/* In-place scatter-back (:474-485): tokens[b, inds[b,q], :] = query[b,q,:] for q < min(focus[b], Nq).
 * focus_token_nums (b) int32 on the device (no host sync, unlike the reference's per-image slicing). */
int sdetr_token_scatter(float *tokens, const float *query, const int64_t *inds, int64_t inds_stride,
                        const int32_t *focus_token_nums, int batch, int num_value, int channels,
                        int num_query, sdetr_stream_t stream);

/* Background embedding (:488-495, position_encoding.py:81-95), in place:
 * tokens[b,t,:] += [col_embed[x] | row_embed[y]] unless mask[b,t] or t in last_inds[b,:num_last].
 * row_embed / col_embed: (num_embeddings, channels/2) tables (max_num_embedding, :400-407).  A feature map taller or
 * wider than the table is an index error in the reference (nn.Embedding): pass the HOST copy of the level shapes in
 * spatial_shapes_host ((L,2) int64 [H,W], may be NULL) to get SDETR_ERR_INVALID_ARG for it; the kernel itself never
 * reads outside the tables (indices are clamped).  flags: (b,Nv) uint8 scratch. */
int sdetr_background_embed(float *tokens, const uint8_t *mask, const int64_t *last_inds, int64_t inds_stride,
                           int num_last, const float *row_embed, const float *col_embed, int num_embeddings,
                           const int64_t *spatial_shapes, const int64_t *spatial_shapes_host,
                           const int64_t *level_start_index, int batch, int num_value, int channels,
                           int num_levels, uint8_t *flags, sdetr_stream_t stream);

/* Coarse-to-fine score modulation (:134-143): out = mem + mem * up * alpha[alpha_index], up = bilinear resize
 * (align_corners=True) of the coarser level's score map (b,Hc*Wc) to (H,W).  mem/out (b,H*W,C) with row
 * strides; alpha is read on the device. */
int sdetr_score_modulate(const float *mem, int64_t mem_batch_stride, const float *coarse_score,
                         int64_t coarse_batch_stride, const float *alpha, int alpha_index, int batch, int H,
                         int W, int Hc, int Wc, int channels, float *out, sdetr_stream_t stream);

/* value rows of padded tokens -> 0 (ms_deform_attn.py:318-319), only masked rows are touched.
 * rows: (b*Nv) rows of `row_floats` contiguous floats at stride row_stride. */
int sdetr_zero_masked_rows(float *rows, int64_t row_stride, int row_floats, const uint8_t *mask,
                           int64_t num_rows, sdetr_stream_t stream);

/* MaskPredictor middle (salience_transformer.py:40-45), in place on z (b,num_rows,channels):
 * z = GELU(z) (exact erf form); then z[b,:,half:] = mean over the num_rows tokens of image b (padded tokens included,
 * like the reference).  Deterministic two-pass reduction; workspace from sdetr_gelu_colmean_workspace. */
size_t sdetr_gelu_colmean_workspace(int batch, int num_rows, int channels, int half);
int sdetr_gelu_colmean(float *z, int batch, int num_rows, int channels, int half, void *workspace,
                       size_t workspace_bytes, sdetr_stream_t stream);

/* mc_score = max_c(class_logits) * fg (:366).  logits (rows, num_classes) with row pitch row_pitch -> out (rows). */
int sdetr_class_max_times_fg(const float *logits, int64_t row_pitch, const float *fg, int64_t rows, int num_classes,
                             float *out, sdetr_stream_t stream);

/* y = LayerNorm(x + r) * gamma + beta (:390-391, :349-350); rows of `channels` (<= 1024, multiple of 4).
 * May run in place (y == x). */
int sdetr_add_layernorm(const float *x, const float *r, const float *gamma, const float *beta, float eps,
                        int64_t rows, int channels, float *y, sdetr_stream_t stream);

/* Front end: per-level NCHW maps -> token layout in one pass (base_transformer.py:21-32,
 * salience_transformer.py:107-113):  feat_tok[b,t,:] = feat_l[b,:,y,x];  lpos_tok = pos_l + level_embeds[l];
 * x_tok = (feat_tok + lpos_tok) * keep[b,t]  (keep = ~padding & proposal-valid, base_transformer.py:100-108).
 * feats_host / pos_host: HOST arrays of num_levels device pointers to (b,C,H_l*W_l) fp32; level_size_host = H_l*W_l. */
int sdetr_flatten_tokens(const float *const *feats_host, const float *const *pos_host, const float *level_embeds,
                         const float *keep, const int32_t *level_size_host, int batch, int channels, int num_levels,
                         float *feat_tok, float *lpos_tok, float *x_tok, sdetr_stream_t stream);

/* Same, with the position embedding already in token layout (b,Nv,C) (sdetr_sine_pos_tokens): only the feature maps are
 * transposed; lpos_tok = pos_tokens + level_embeds[l]. */
int sdetr_flatten_tokens_pos(const float *const *feats_host, const float *pos_tokens, const float *level_embeds,
                             const float *keep, const int32_t *level_size_host, int batch, int channels, int num_levels,
                             float *feat_tok, float *lpos_tok, float *x_tok, sdetr_stream_t stream);
/* testing knob: 1 (default) = levels with a token count % 4 == 0 use the vectorised register-transpose kernel; 0 = all levels
 * through the 32x32 shared-memory tile kernel.  Results are bit-identical. */
int sdetr_flatten_set_vectorized(int enable);

/* Salience supervision targets (training side of the filter; SalienceCriterion.get_mask_single_level with noise_scale 0,
 * models/detectors/salience_detr.py:64-114): target[b,t] = max over the boxes that contain the token's pixel centre
 * ((x+.5) stride_x, (y+.5) stride_y) of 1 - sqrt(dx^2 + dy^2) / 2, and 0 unless one containing box has its largest border
 * distance in (limit_lo[l], limit_hi[l]].  boxes (b,max_boxes,4) xyxy in pixels, 16-byte aligned; num_boxes (b,) int32. */
int sdetr_salience_targets(const float *boxes_xyxy, const int32_t *num_boxes, int max_boxes, int batch, int num_value,
                           int num_levels, const int32_t *level_h_host, const int32_t *level_w_host, const float *stride_y_host,
                           const float *stride_x_host, const float *limit_lo_host, const float *limit_hi_host, float *target,
                           sdetr_stream_t stream);

/* Layout hand-off to / from a convolutional neck (salience_transformer.py:185-192): tokens (b,Nv,C) <-> per-level NCHW
 * maps; maps_host = HOST array of num_levels device pointers to (b,C,H_l*W_l) fp32.  to_maps != 0: tokens -> maps (the
 * reference's split + transpose + contiguous + reshape); to_maps == 0: maps -> tokens (flatten(2).transpose(1,2) + cat). */
int sdetr_token_map_transpose(float *tokens, float *const *maps_host, const int32_t *level_size_host, int batch, int channels,
                              int num_levels, int to_maps, sdetr_stream_t stream);

/* Everything the path derives from the padding masks, in two launches (replaces ~60 ATen launches and makes a fresh-mask
 * batch cheap).  mask (b,Nv) uint8 (1 = padding), levels given by host (H_l, W_l).
 *   valid_token_nums[b,l] = #valid tokens;  focus_token_nums[b,l] = int(float(valid) * level_filter_ratio[l])
 *     (salience_transformer.py:116-119: fp32 multiply, truncation);
 *   valid_ratios[b,l] = (valid W of the first row / W_l, valid H of the first column / H_l)   (base_transformer.py:48-56);
 *   keep[b,t] = 1.0 iff the token is not padded and its proposal centre ((x+.5)/valid_W, (y+.5)/valid_H) and size
 *     0.05 * 2^l lie in (0.01, 0.99)   (base_transformer.py:84-108);
 *   ynorm / xnorm[b,t] = (cumsum of the valid mask along y / x + pos_offset) / (last + pos_eps) * pos_scale, the
 *     normalised coordinates of PositionEmbeddingSine (models/bricks/position_encoding.py:48-56), fp32, same op order. */
int sdetr_mask_plan(const uint8_t *mask, int batch, int num_value, int num_levels, const int32_t *level_h_host,
                    const int32_t *level_w_host, const float *level_filter_ratio_host, float pos_offset, float pos_eps,
                    float pos_scale, float *ynorm, float *xnorm, int32_t *valid_token_nums, int32_t *focus_token_nums,
                    float *valid_ratios, float *keep, sdetr_stream_t stream);
/* pos_tokens[r, :] = [sin/cos(ynorm[r] / dim_ty[j]) (j < F) | sin/cos(xnorm[r] / dim_tx[j])], sin on even j, cos on odd j
 * (position_encoding.py:58-64); rows = b*Nv, F = num_pos_feats (multiple of 4), dim_t* = temperature^(2 (j//2) / F). */
int sdetr_sine_pos_tokens(const float *ynorm, const float *xnorm, const float *dim_ty, const float *dim_tx, int64_t rows,
                          int num_pos_feats, float *pos_tokens, sdetr_stream_t stream);

/* Dense self-attention core of the 300-token pre-attention (salience_transformer.py:372-376; the attention inside
 * nn.MultiheadAttention): out = softmax(Q K^T / sqrt(d)) V per (image, head).  qk (b,n,2,heads,d): projected queries
 * then keys; v (b,n,heads,d); out (b,n,heads*d).  head_dim must be 32; K^T, V and the score tile live in shared memory:
 * n up to ~460, larger n returns SDETR_ERR_UNSUPPORTED. */
int sdetr_attention_small(const float *qk, const float *v, float *out, int batch, int n, int heads, int head_dim,
                          sdetr_stream_t stream);
/* Same core on a packed projection buffer qkv (b,n,3,heads,d) (what sdetr_mha_in_proj writes). */
int sdetr_attention_qkv(const float *qkv, float *out, int batch, int n, int heads, int head_dim, sdetr_stream_t stream);

/* Pre-attention front (salience_transformer.py:368-372 + the packed in-projection of nn.MultiheadAttention):
 * t_out[b,j,:] = tokens[b,index[b,j],:];  x = t + pos[b,index[b,j],:];
 * qkv[b,j,:] = [x Wq^T + bq | x Wk^T + bk | t Wv^T + bv].
 * tokens/pos (b,num_rows,C); index (b,k) int64; w_in_t = in_proj_weight TRANSPOSED, (C,3C) row-major; b_in (3C);
 * t_out (b,k,C); qkv (b,k,3C).  channels must be 256.  fp32 FMA arithmetic. */
int sdetr_mha_in_proj(const float *tokens, const float *pos, const int64_t *index, int batch, int num_rows, int k,
                      int channels, const float *w_in_t, const float *b_in, float *t_out, float *qkv,
                      sdetr_stream_t stream);

/* Pre-attention back (salience_transformer.py:373-379): y = LayerNorm(t + attn Wo^T + bo) with (gamma, beta, eps);
 * dst[b,index[b,j],:] = y[b,j,:] (indices unique per image).  attn, t (b,k,C); w_out_t = out_proj.weight TRANSPOSED
 * (C,C); dst (b,num_rows,C), updated in place.  channels must be 256.
 * Optional (both or neither): pos (b,num_rows,C) and dst_sum (b,num_rows,C) -- the same rows of dst_sum receive
 * y + pos[b,index[b,j],:], keeping a `query + query_pos` buffer (sdetr_token_gather's query_sum) current. */
int sdetr_mha_out_proj_ln_scatter(const float *attn, const float *t, const float *w_out_t, const float *b_out,
                                  const float *gamma, const float *beta, float eps, const int64_t *index, float *dst,
                                  const float *pos, float *dst_sum, int batch, int num_rows, int k, int channels,
                                  sdetr_stream_t stream);

/* Row gather / scatter by per-image index (the top-k tokens of the pre-attention, salience_transformer.py:368-379):
 * out[b,j,:] = src[b,index[b,j],:]   /   dst[b,index[b,j],:] = src[b,j,:]  (indices unique per image).
 * src/dst (b,num_rows,C), index (b,k) int64. */
int sdetr_rows_gather(const float *src, const int64_t *index, int batch, int num_rows, int k, int channels, float *out,
                      sdetr_stream_t stream);
/* t[b,j,:] = src[b,index[b,j],:],  x[b,j,:] = t[b,j,:] + pos[b,index[b,j],:]  (:368-371 in one pass) */
int sdetr_rows_gather_add(const float *src, const float *pos, const int64_t *index, int batch, int num_rows, int k,
                          int channels, float *t, float *x, sdetr_stream_t stream);
int sdetr_rows_scatter(float *dst, const int64_t *index, int batch, int num_rows, int k, int channels, const float *src,
                       sdetr_stream_t stream);

/* 3xTF32 operand split for the dense projections (tensor cores with fp32-class accuracy):
 * x (rows, K) with row stride x_row_stride -> out (rows, K/chunk, 3, chunk): per K-chunk [hi | hi | lo]
 * (layout_b = 0, activations) or [hi | lo | hi] (layout_b = 1, weights), hi = tf32(x), lo = tf32(x - hi);
 * relu != 0 applies max(x, 0) first (fuses the FFN activation, salience_transformer.py:348).  A TF32 GEMM over
 * a chunk's 3*chunk columns of both operands yields A_hi.B_hi + A_hi.B_lo + A_lo.B_hi for that chunk; chunks are
 * accumulated by the GEMM epilogue in fp32 (chunk == K: a single GEMM). */
int sdetr_split_tf32(const float *x, int64_t x_row_stride, int64_t rows, int K, int chunk, int layout_b, int relu,
                     float *out, sdetr_stream_t stream);

/* Dense projection on the tcgen05 tensor cores with fp32-class accuracy (hand-written sm_100a GEMM):
 *   C[M,N] = act(A)[M,K] . W[N,K]^T + bias,  act = ReLU (relu_a = 1, salience_transformer.py:348), exact GELU (relu_a = 2, :23-29) or identity (0).
 * A (M,K) row pitch lda floats; W_hi/W_lo (N,K) contiguous = sdetr_split_tf32_pair(W); bias (N) nullable;
 * C (M,N) row pitch ldc.  K % 32 == 0; A/W 16-byte aligned, lda % 4 == 0.  The activation is split into TF32 pieces
 * inside the kernel (no extra HBM pass); products A_hi.W_hi + A_hi.W_lo + A_lo.W_hi accumulate in TMEM (fp32).
 * Replaces the cuBLAS path for: value_proj / sampling_offsets|attention_weights / output_proj
 * (models/bricks/ms_deform_attn.py:316,322-328,375), FFN (salience_transformer.py:347-351), class head (:462),
 * MaskPredictor (:16-47), enc_output (base_transformer.py:111). */
/* Same contract with the RAW weight W (N,K): the kernel splits both operands itself (32 KB instead of 48 KB of TMA
 * traffic per k-block), keeps the split activation in tensor memory, and runs two CTAs per SM. */
int sdetr_gemm_3xtf32_raw(const float *A, int64_t lda, const float *W, const float *bias, float *C, int64_t ldc, int M,
                          int N, int K, int act, sdetr_stream_t stream);
/* kernel variants: sdetr_gemm_3xtf32: 0 = "SS" (both operands from shared memory), 1 = "TS" (the split activation is
 * written to tensor memory by the converter warps and the MMAs read A from TMEM); sdetr_gemm_3xtf32_raw: 2 = "TS2"
 * (one tile per CTA, two CTAs per SM), 3 = "P" (default: persistent tile loop, double-buffered TMEM accumulator,
 * dedicated epilogue warps) */
int sdetr_gemm_set_variant(int variant);
/* debugging aid: when set, CTA (0,0) of every sdetr_gemm_3xtf32 launch records clock64() per pipeline event */
int sdetr_gemm_set_trace(long long *device_buffer);
int sdetr_split_tf32_pair(const float *w, int64_t count, float *w_hi, float *w_lo, sdetr_stream_t stream);
int sdetr_gemm_3xtf32(const float *A, int64_t lda, const float *W_hi, const float *W_lo, const float *bias, float *C,
                      int64_t ldc, int M, int N, int K, int relu_a, sdetr_stream_t stream);
/* Same contract as sdetr_gemm_3xtf32 (pre-split weight), on the persistent kernel: one CTA per SM walks the output
 * tiles, the weight halves arrive as two TMA tiles per stage and only the activation is converted in the kernel. */
int sdetr_gemm_3xtf32_pre(const float *A, int64_t lda, const float *W_hi, const float *W_lo, const float *bias, float *C,
                          int64_t ldc, int M, int N, int K, int act, sdetr_stream_t stream);

/* ---- "3xFP16": the same error-compensated product on tcgen05.mma.kind::f16 -----------------------------------------
 * An fp16 significand is as wide as a TF32 one (11 bits), so  A_hi.W_hi + A_hi.W_lo + A_lo.W_hi  with
 * hi = fp16(x), lo = fp16(x - hi) keeps the 22 significand bits per operand of 3xTF32 at twice the MMA rate and half the
 * operand bytes.  Exponent range is restored by exact power-of-two scalings: the kernel multiplies the activation by 16
 * before the split (full accuracy for 2^-7 <= |x| < 4094, absolute error floor 2^-29 below, inf/NaN above), the caller
 * scales the weight by `scale` = 2^s with max|scale * W| in [2^13, 2^14) when splitting, and the epilogue multiplies the
 * fp32 accumulator by 2^-(4+s) before the bias.  Replaces the same reference projections as sdetr_gemm_3xtf32
 * (models/bricks/ms_deform_attn.py:316,322-328,375; salience_transformer.py:347-351,462,16-47; base_transformer.py:111).
 * W (N,K) fp32 contiguous -> W_hi, W_lo (N,K) fp16 (2-byte elements, 16-byte aligned). */
int sdetr_split_f16_pair(const float *W, int64_t count, float scale, void *W_hi, void *W_lo, sdetr_stream_t stream);
/* C[M,N] = act(A)[M,K] . W^T + bias, W given as the pair above and the `w_scale` it was split with; K % 64 == 0;
 * A (M,K) fp32 with row pitch lda (floats, multiple of 4); act: 0 none, 1 ReLU, 2 exact GELU applied to A on load;
 * C row pitch ldc (TMA stores when ldc % 4 == 0). */
int sdetr_gemm_f16x3_pre(const float *A, int64_t lda, const void *W_hi, const void *W_lo, float w_scale, const float *bias,
                         float *C, int64_t ldc, int M, int N, int K, int act, sdetr_stream_t stream);
/* The same product with BOTH scales as device scalars (powers of two; sdetr_pow2_scale): the kernel multiplies A by *a_scale_dev
 * instead of the fixed 16 (activations of unknown magnitude: gradients), the weight pair was split with *w_scale_dev
 * (sdetr_split_f16_pair_dev: weights that change every training step, no host read of max|W|), and the epilogue divides both out.
 * No host synchronisation anywhere.  No fused input activation. */
int sdetr_gemm_f16x3_scaled(const float *A, int64_t lda, const float *a_scale_dev, const void *W_hi, const void *W_lo,
                            const float *w_scale_dev, const float *bias, float *C, int64_t ldc, int M, int N, int K,
                            sdetr_stream_t stream);
int sdetr_split_f16_pair_dev(const float *W, int64_t count, const float *scale_dev, void *W_hi, void *W_lo, sdetr_stream_t stream);
/* *scale = 2^s with max|x| * 2^s in [2^(target_log2 - 1), 2^target_log2) (1 for an all-zero x); one launch, deterministic.
 * state: 8 bytes of device memory, zero before the first call (the kernel leaves them zero). */
int sdetr_pow2_scale(const float *x, int64_t count, int target_log2, void *state, float *scale, sdetr_stream_t stream);
/* benchmarking knob: 0 (default) = always the streaming kernel, 1 = K <= 256 and >= 2 output tiles per work unit use the
 * activation-stationary kernel (the split activation panel stays in tensor memory across the unit's output tiles) */
int sdetr_gemm_f16x3_set_as(int enable);
/* benchmarking knob, streaming kernel epilogue: 0 (default) = two shared 16 KB boxes + TMA bulk stores, 1 = warp-private
 * transpose boxes + coalesced 128-bit global stores (no CTA-level barrier in the epilogue) */
int sdetr_gemm_f16x3_set_epilogue(int variant);
/* benchmarking knob, streaming kernel: 4 (default) or 8 epilogue warps (8: two groups of four, each draining half of a tile's
 * columns through its own store box) */
int sdetr_gemm_f16x3_set_epilogue_warps(int warps);
/* debugging aid: when set, CTA 0 of every streaming sdetr_gemm_f16x3_pre launch records clock64() per pipeline event
 * (device_buffer: 10 x 256 int64) */
int sdetr_gemm_f16x3_set_trace(long long *device_buffer);
/* benchmarking knob: 1 = clusters of two CTAs (adjacent 128-row panels, same output columns) share the weight k-blocks by
 * TMA multicast (measured slower, off by default); 0 (default) = independent CTAs */
int sdetr_gemm_f16x3_set_cluster(int enable);

/* ---- MaskPredictor of one small feature level in two launches ---------------------------------------------------------------
 * Replaces, for a level of few token rows, score modulation + MaskPredictor (models/bricks/salience_transformer.py:16-47, :134-143):
 * x = m + m * bilinear(coarse_score, align_corners=True) * alpha[alpha_index] (coarse_score NULL: x = m), LayerNorm, Linear(C,C),
 * GELU, token mean of the upper half of the channels (over the H*W tokens of each image), Linear(C,C/2), GELU, Linear(C/2,C/4),
 * GELU, Linear(C/4,1).  channels must be 256.  fp32 FMA arithmetic.
 * mem: first token row of the level in image 0, rows of `channels` floats, images mem_batch_stride floats apart;
 * weights TRANSPOSED to (in, out) row-major: w1_t (C,C), w2a_t (C,C/2), w2b_t (C/2,C/4); w2c (C/4);
 * out: the level's raw scores, H*W per image, images out_batch_stride floats apart;
 * workspace: sdetr_mask_predictor_level_workspace_floats(batch, H, W) floats. */
int64_t sdetr_mask_predictor_level_workspace_floats(int batch, int H, int W);
int sdetr_mask_predictor_level(const float *mem, int64_t mem_batch_stride, int batch, int H, int W, int channels,
                               const float *coarse_score, int64_t coarse_batch_stride, int Hc, int Wc, const float *alpha,
                               int alpha_index, const float *ln_gamma, const float *ln_beta, float eps, const float *w1_t,
                               const float *b1, const float *w2a_t, const float *b2a, const float *w2b_t, const float *b2b,
                               const float *w2c, const float *b2c, float *workspace, int64_t workspace_floats, float *out,
                               int64_t out_batch_stride, sdetr_stream_t stream);

/* ---- fused encoder FFN: y = LayerNorm(x + linear2(ReLU(linear1(x)))) ------------------------------------------------------
 * Replaces `forward_ffn` + `norm2` of the encoder layer (models/bricks/salience_transformer.py:347-351, :391) for embed_dim 256:
 * one persistent tcgen05 kernel keeps a 128-row panel on chip from x to the output accumulator (the hidden activations never
 * reach HBM; only the weights stream), then a row kernel adds the partial sums, b2 and the residual and normalises.
 * Same 3xFP16 arithmetic and value domain as sdetr_gemm_f16x3_pre (both weights given as sdetr_split_f16_pair pairs with
 * their scales; W1 (hidden,256), W2 (256,hidden); hidden % 128 == 0).
 * x (M,256) fp32 with row pitch ldx (floats); y (M,256) contiguous, may alias x when ldx == 256;
 * gamma == beta == NULL: y = linear2(ReLU(linear1(x))) without residual / LayerNorm.
 * workspace: sdetr_ffn_fused_workspace_floats(M, hidden) floats (partial sums of panels shared by two CTAs: the persistent
 * CTAs take equal numbers of 128-hidden-unit chunks of the (panel, chunk) sequence, whatever the panel count). */
int64_t sdetr_ffn_fused_workspace_floats(int M, int hidden);
/* host-side view of the work decomposition (for tests): lo[i] = first item of CTA i in the panel-major (panel, chunk) sequence,
 * lo[G] = #panels * hidden / 128; returns the CTA count G (or -G when `capacity` < G + 1, 0 for bad sizes) */
int sdetr_ffn_fused_ranges(int M, int hidden, int64_t *lo, int capacity);
int sdetr_ffn_fused_layernorm(const float *x, int64_t ldx, const void *W1_hi, const void *W1_lo, float w1_scale, const float *b1,
                              const void *W2_hi, const void *W2_lo, float w2_scale, const float *b2, const float *gamma,
                              const float *beta, float eps, int M, int hidden, float *workspace, int64_t workspace_floats, float *y,
                              sdetr_stream_t stream);
/* benchmarking knob: 1 (default) = equal chunk counts per CTA, 0 = whole panels per CTA (idle SMs when the panel count is not
 * near a multiple of the SM count).  Changes the workspace size: query it after setting. */
int sdetr_ffn_fused_set_balance(int enable);
/* debugging aid: CTA 0 of every sdetr_ffn_fused_layernorm launch records clock64() per pipeline event (8 x 256 int64) */
int sdetr_ffn_fused_set_trace(long long *device_buffer);
/* debugging / benchmarking knob: at most n persistent CTAs (0 = one per SM) */
int sdetr_ffn_fused_set_max_ctas(int n);

#ifdef __cplusplus
}
#endif
#endif /* SDETR_B200_H */
