"""ctypes binding of the C-ABI library declared in ``include/sdetr_b200.h``.

PyTorch is plumbing only here: it owns device memory and the stream; every wrapper validates its
tensors the way the reference op does (contiguous + CUDA, ``AT_ASSERTM`` in
models/bricks/ops/cuda/ms_deform_attn_cuda.cu:20-30), allocates the outputs with ``torch.empty`` and passes
raw pointers plus ``torch.cuda.current_stream().cuda_stream``.  A non-zero return code becomes a
``RuntimeError`` carrying ``sdetr_last_error()``.

There is NO fallback: if the shared library is missing the import of the product path fails loudly.
"""
from __future__ import annotations

import ctypes
import os
from typing import Optional, Sequence

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libsdetr_b200.so")
_lib = None

_vp, _i, _i64, _sz, _f = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_size_t, ctypes.c_float

# name -> (restype, argtypes); must list every symbol of include/sdetr_b200.h (tests check this)
SIGNATURES = {
    "sdetr_version": (_i, []),
    "sdetr_last_error": (ctypes.c_char_p, []),
    "sdetr_launch_count": (ctypes.c_ulonglong, []),
    "sdetr_set_persistent_ctas": (_i, [_i]),
    "sdetr_set_option": (_i, [ctypes.c_char_p, _i]),
    "sdetr_msda_forward": (_i, [_vp] * 6 + [_i] * 7 + [_vp]),
    "sdetr_msda_forward_ex": (_i, [_vp, _i64, _i64] + [_vp] * 5 + [_i] * 7 + [_vp, _i, _vp]),
    "sdetr_msda_fused_forward": (_i, [_vp, _i64, _i64, _vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp] + [_i] * 7 + [_vp, _i, _vp]),
    "sdetr_msda_fused_forward_boxes": (_i, [_vp, _i64, _i64, _vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp] + [_i] * 7 + [_vp, _i, _vp]),
    "sdetr_msda_set_host_shapes": (_i, [_i, _vp, _vp]),
    "sdetr_nms_topk_index": (_i, [_vp, _i, _i, _i, _i, _vp, _vp, _f, _vp, _vp, _vp, _vp]),
    "sdetr_msda_backward": (_i, [_vp] * 9 + [_i] * 7 + [_vp]),
    "sdetr_salience_select_workspace": (_sz, [_i, _i, _i]),
    "sdetr_salience_select": (_i, [_vp] * 7 + [_i] * 4 + [_vp] * 5 + [_sz, _vp]),
    "sdetr_order_prefixes": (_i, [_vp, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "sdetr_topk_workspace": (_sz, [_i, _i]),
    "sdetr_topk_desc": (_i, [_vp, _i, _i, _i, _vp, _vp, _sz, _vp]),
    "sdetr_token_gather": (_i, [_vp] * 5 + [_i64, _vp, _vp] + [_i] * 5 + [_vp] * 6),
    "sdetr_token_scatter": (_i, [_vp, _vp, _vp, _i64, _vp] + [_i] * 4 + [_vp]),
    "sdetr_background_embed": (_i, [_vp, _vp, _vp, _i64, _i, _vp, _vp, _i, _vp, _vp, _vp] + [_i] * 4 + [_vp, _vp]),
    "sdetr_score_modulate": (_i, [_vp, _i64, _vp, _i64, _vp] + [_i] * 7 + [_vp, _vp]),
    "sdetr_zero_masked_rows": (_i, [_vp, _i64, _i, _vp, _i64, _vp]),
    "sdetr_gelu_colmean_workspace": (_sz, [_i, _i, _i, _i]),
    "sdetr_gelu_colmean": (_i, [_vp, _i, _i, _i, _i, _vp, _sz, _vp]),
    "sdetr_class_max_times_fg": (_i, [_vp, _i64, _vp, _i64, _i, _vp, _vp]),
    "sdetr_add_layernorm": (_i, [_vp, _vp, _vp, _vp, _f, _i64, _i, _vp, _vp]),
    "sdetr_split_tf32": (_i, [_vp, _i64, _i64, _i, _i, _i, _i, _vp, _vp]),
    "sdetr_gemm_3xtf32_raw": (_i, [_vp, _i64, _vp, _vp, _vp, _i64, _i, _i, _i, _i, _vp]),
    "sdetr_gemm_set_trace": (_i, [_vp]),
    "sdetr_gemm_set_variant": (_i, [_i]),
    "sdetr_split_tf32_pair": (_i, [_vp, _i64, _vp, _vp, _vp]),
    "sdetr_gemm_3xtf32": (_i, [_vp, _i64, _vp, _vp, _vp, _vp, _i64, _i, _i, _i, _i, _vp]),
    "sdetr_gemm_3xtf32_pre": (_i, [_vp, _i64, _vp, _vp, _vp, _vp, _i64, _i, _i, _i, _i, _vp]),
    "sdetr_flatten_tokens_pos": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "sdetr_salience_targets": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "sdetr_token_map_transpose": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "sdetr_mask_plan": (_i, [_vp, _i, _i, _i, _vp, _vp, _vp, _f, _f, _f, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "sdetr_sine_pos_tokens": (_i, [_vp, _vp, _vp, _vp, _i64, _i, _vp, _vp]),
    "sdetr_split_f16_pair": (_i, [_vp, _i64, _f, _vp, _vp, _vp]),
    "sdetr_gemm_f16x3_set_as": (_i, [_i]),
    "sdetr_gemm_f16x3_set_cluster": (_i, [_i]),
    "sdetr_gemm_f16x3_set_epilogue": (_i, [_i]),
    "sdetr_gemm_f16x3_set_epilogue_warps": (_i, [_i]),
    "sdetr_gemm_f16x3_set_trace": (_i, [_vp]),
    "sdetr_gemm_f16x3_pre": (_i, [_vp, _i64, _vp, _vp, _f, _vp, _vp, _i64, _i, _i, _i, _i, _vp]),
    "sdetr_gemm_f16x3_scaled": (_i, [_vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i, _i, _i, _vp]),
    "sdetr_split_f16_pair_dev": (_i, [_vp, _i64, _vp, _vp, _vp, _vp]),
    "sdetr_pow2_scale": (_i, [_vp, _i64, _i, _vp, _vp, _vp]),
    "sdetr_mask_predictor_level_workspace_floats": (_i64, [_i, _i, _i]),
    "sdetr_mask_predictor_level": (_i, [_vp, _i64, _i, _i, _i, _i, _vp, _i64, _i, _i, _vp, _i, _vp, _vp, _f, _vp, _vp, _vp, _vp, _vp, _vp,
                                        _vp, _vp, _vp, _i64, _vp, _i64, _vp]),
    "sdetr_ffn_fused_workspace_floats": (_i64, [_i, _i]),
    "sdetr_ffn_fused_ranges": (_i, [_i, _i, _vp, _i]),
    "sdetr_ffn_fused_set_balance": (_i, [_i]),
    "sdetr_ffn_fused_layernorm": (_i, [_vp, _i64, _vp, _vp, _f, _vp, _vp, _vp, _f, _vp, _vp, _vp, _f, _i, _i, _vp, _i64, _vp, _vp]),
    "sdetr_ffn_fused_set_trace": (_i, [_vp]),
    "sdetr_flatten_set_vectorized": (_i, [_i]),
    "sdetr_ffn_fused_set_max_ctas": (_i, [_i]),
    "sdetr_flatten_tokens": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "sdetr_attention_small": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "sdetr_attention_qkv": (_i, [_vp, _vp, _i, _i, _i, _i, _vp]),
    "sdetr_mha_in_proj": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp]),
    "sdetr_mha_out_proj_ln_scatter": (_i, [_vp] * 6 + [_f, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "sdetr_rows_gather": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _vp]),
    "sdetr_rows_gather_add": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp]),
    "sdetr_rows_scatter": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _vp]),
}


def lib():
    """Load (once) the sm_100a library.  Raises if it has not been built (``__graft_entry__.build()``)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a). There is no CPU or PyTorch fallback for this path.")
        cdll = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(cdll, name)
            fn.restype, fn.argtypes = res, args
        _lib = cdll
    return _lib


def set_option(name: str, value: int):
    _check(lib().sdetr_set_option(name.encode(), int(value)), "sdetr_set_option")


def launch_count() -> int:
    return int(lib().sdetr_launch_count())


def _check(rc: int, what: str):
    if rc != 0:
        raise RuntimeError(f"{what} failed (rc={rc}): {lib().sdetr_last_error().decode()}")


# The eager path issues ~120 launches per forward from Python; `torch.cuda.current_stream()` / `current_device()` cost several
# microseconds each (device-index resolution, Stream object construction) and were a third of the host time per forward
# (tools/profile_eager_cpu.py).  The raw accessors below return the same values.
_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)
_raw_device = getattr(torch._C, "_cuda_getDevice", None)


def _stream() -> int:
    if _raw_stream is not None and _raw_device is not None:
        return _raw_stream(_raw_device())
    return torch.cuda.current_stream().cuda_stream


def _current_device() -> int:
    return _raw_device() if _raw_device is not None else torch.cuda.current_device()


def _req(t: torch.Tensor, name: str, dtype=None):
    # same contract as the reference's AT_ASSERTM checks -> RuntimeError
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor")
    if t.device.index != _current_device():
        # the call launches on the CURRENT device's current stream; a tensor of another device would be dereferenced there
        raise RuntimeError(f"{name} lives on {t.device} but cuda:{_current_device()} is current: "
                           f"wrap the call in `with torch.cuda.device({t.device.index}):`")
    if not t.is_contiguous():
        raise RuntimeError(f"{name} tensor has to be contiguous")
    if dtype is not None and t.dtype != dtype:
        raise RuntimeError(f"{name} must be {dtype}, got {t.dtype}")
    return t.data_ptr()


def _host_i32(xs: Sequence[int]):
    return (ctypes.c_int32 * len(xs))(*[int(x) for x in xs])


def _host_i64(xs: Sequence[int]):
    return (ctypes.c_int64 * len(xs))(*[int(x) for x in xs])


# ---- MSDA core ------------------------------------------------------------------------------------------
def msda_forward(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, query_order=None, schedule=0):
    """`_C.ms_deform_attn_forward` contract: value (b,Nv,M,D) f32 -> (b,Nq,M*D), fresh tensor."""
    b, nv, m, d = value.shape
    nq, L, P = sampling_loc.shape[1], sampling_loc.shape[3], sampling_loc.shape[4]
    out = torch.empty(b, nq, m * d, device=value.device, dtype=torch.float32)
    rc = lib().sdetr_msda_forward_ex(
        _req(value, "value", torch.float32), nv * m * d, m * d, _req(spatial_shapes, "spatial_shapes", torch.int64),
        _req(level_start_index, "level_start_index", torch.int64), _req(sampling_loc, "sampling_loc", torch.float32),
        _req(attn_weight, "attn_weight", torch.float32), out.data_ptr(), b, nv, m, d, L, nq, P,
        _req(query_order, "query_order", torch.int32) if query_order is not None else None, schedule, _stream())
    _check(rc, "sdetr_msda_forward_ex")
    return out


def msda_forward_plain(value, spatial_shapes, level_start_index, sampling_loc, attn_weight):
    """The plain entry point (no strides / order), same result as :func:`msda_forward`."""
    b, nv, m, d = value.shape
    nq, L, P = sampling_loc.shape[1], sampling_loc.shape[3], sampling_loc.shape[4]
    out = torch.empty(b, nq, m * d, device=value.device, dtype=torch.float32)
    rc = lib().sdetr_msda_forward(
        _req(value, "value", torch.float32), _req(spatial_shapes, "spatial_shapes", torch.int64),
        _req(level_start_index, "level_start_index", torch.int64), _req(sampling_loc, "sampling_loc", torch.float32),
        _req(attn_weight, "attn_weight", torch.float32), out.data_ptr(), b, nv, m, d, L, nq, P, _stream())
    _check(rc, "sdetr_msda_forward")
    return out


KERNEL_TIMERS = None  # measurement aid (bench.py): {"msda": [(event_before, event_after), ...]} filled per launch


def _timer_event():
    """CUDA event that can be recorded inside a stream capture AND timed after the graph has run (external event node)."""
    ev = torch.cuda.Event(enable_timing=True, external=torch.cuda.is_current_stream_capturing())
    ev.record()
    return ev


def msda_fused_forward(value_buf, value_batch_stride, value_token_stride, value_offset, spatial_shapes,
                       level_start_index, ref_points, proj, heads, head_dim, levels, points, num_value,
                       query_order=None, schedule=0, want_loc_attn=False):
    """softmax + sampling locations + core in one launch.  ``value_buf`` may be a wide projection buffer;
    ``value_offset`` (floats) selects this layer's column slice.  proj: (b,Nq,>=3*M*L*P) last-dim contiguous."""
    b, nq = proj.shape[0], proj.shape[1]
    if not (proj.is_cuda and proj.stride(2) == 1 and proj.stride(0) == nq * proj.stride(1)):
        raise RuntimeError("proj must be a CUDA tensor with contiguous rows")
    out = torch.empty(b, nq, heads * head_dim, device=proj.device, dtype=torch.float32)
    loc = attn = None
    if want_loc_attn:
        loc = torch.empty(b, nq, heads, levels, points, 2, device=proj.device, dtype=torch.float32)
        attn = torch.empty(b, nq, heads, levels, points, device=proj.device, dtype=torch.float32)
    if not value_buf.is_cuda or value_buf.dtype != torch.float32:
        raise RuntimeError("value must be a CUDA float32 tensor")
    t0 = _timer_event() if KERNEL_TIMERS is not None else None
    if ref_points.shape[-1] not in (2, 4):
        raise ValueError(f"Last dim of reference_points must be 2 or 4, but get {ref_points.shape[-1]} instead.")
    fn = lib().sdetr_msda_fused_forward if ref_points.shape[-1] == 2 else lib().sdetr_msda_fused_forward_boxes
    rc = fn(
        value_buf.data_ptr() + 4 * value_offset, value_batch_stride, value_token_stride,
        _req(spatial_shapes, "spatial_shapes", torch.int64), _req(level_start_index, "level_start_index", torch.int64),
        _req(ref_points, "reference_points", torch.float32), proj.data_ptr(), proj.stride(1), out.data_ptr(),
        loc.data_ptr() if loc is not None else None, attn.data_ptr() if attn is not None else None, b, num_value,
        heads, head_dim, levels, nq, points,
        _req(query_order, "query_order", torch.int32) if query_order is not None else None, schedule, _stream())
    _check(rc, "sdetr_msda_fused_forward")
    if t0 is not None:
        KERNEL_TIMERS.setdefault("msda", []).append((t0, _timer_event()))
    return (out, loc, attn) if want_loc_attn else out


def msda_set_host_shapes(shapes):
    """Level shapes as host integers for the TMA-staged variant's tensor maps (option "msda_tma")."""
    _check(lib().sdetr_msda_set_host_shapes(len(shapes), _host_i32([h for h, _ in shapes]), _host_i32([w for _, w in shapes])),
           "sdetr_msda_set_host_shapes")


def nms_topk_index(topk_index, shapes, iou_threshold=0.3):
    """topk_index (b,k) int64, candidates in descending score order -> (kept_index (b,k), kept_count (b,) int32, keep_flag (b,k) u8)."""
    b, k = topk_index.shape
    dev = topk_index.device
    kept = torch.empty(b, k, device=dev, dtype=torch.int64)
    count = torch.empty(b, device=dev, dtype=torch.int32)
    flag = torch.empty(b, k, device=dev, dtype=torch.uint8)
    rc = lib().sdetr_nms_topk_index(_req(topk_index, "topk_index", torch.int64), b, k, sum(h * w for h, w in shapes), len(shapes),
                                    _host_i32([h for h, _ in shapes]), _host_i32([w for _, w in shapes]), float(iou_threshold),
                                    kept.data_ptr(), count.data_ptr(), flag.data_ptr(), _stream())
    _check(rc, "sdetr_nms_topk_index")
    return kept, count, flag


def msda_backward(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, grad_output):
    """`_C.ms_deform_attn_backward` contract -> (grad_value, grad_sampling_loc, grad_attn_weight)."""
    b, nv, m, d = value.shape
    nq, L, P = sampling_loc.shape[1], sampling_loc.shape[3], sampling_loc.shape[4]
    gv = torch.empty_like(value)
    gl = torch.empty_like(sampling_loc)
    ga = torch.empty_like(attn_weight)
    rc = lib().sdetr_msda_backward(
        _req(value, "value", torch.float32), _req(spatial_shapes, "spatial_shapes", torch.int64),
        _req(level_start_index, "level_start_index", torch.int64), _req(sampling_loc, "sampling_loc", torch.float32),
        _req(attn_weight, "attn_weight", torch.float32), _req(grad_output, "grad_output", torch.float32),
        gv.data_ptr(), gl.data_ptr(), ga.data_ptr(), b, nv, m, d, L, nq, P, _stream())
    _check(rc, "sdetr_msda_backward")
    return gv, gl, ga


# ---- salience filter --------------------------------------------------------------------------------------
def salience_select(raw_score, mask_u8, level_start: Sequence[int], level_size: Sequence[int],
                    level_k: Sequence[int], level_width: Optional[Sequence[int]] = None,
                    level_stride: Optional[Sequence[int]] = None, cell_px: int = 0, workspace=None):
    """-> selected_inds (b,K) i64, selected_score (b,K), foreground_score (b,Nv), tile_order (b,K) i32 | None."""
    b, nv = raw_score.shape
    L = len(level_k)
    K = int(sum(level_k))
    dev = raw_score.device
    inds = torch.empty(b, K, device=dev, dtype=torch.int64)
    score = torch.empty(b, K, device=dev, dtype=torch.float32)
    fg = torch.empty(b, nv, device=dev, dtype=torch.float32)
    want_order = level_width is not None and K > 0
    order = torch.empty(b, K, device=dev, dtype=torch.int32) if want_order else None
    need = lib().sdetr_salience_select_workspace(b, nv, L)
    if workspace is None or workspace.numel() < need:
        workspace = torch.empty(need, device=dev, dtype=torch.uint8)
    rc = lib().sdetr_salience_select(
        _req(raw_score, "raw_score", torch.float32), _req(mask_u8, "mask", torch.uint8), _host_i32(level_start),
        _host_i32(level_size), _host_i32(level_k), _host_i32(level_width) if want_order else None,
        _host_i32(level_stride) if want_order else None, cell_px, b, nv, L, inds.data_ptr(), score.data_ptr(),
        fg.data_ptr(), order.data_ptr() if want_order else None, workspace.data_ptr(), workspace.numel(), _stream())
    _check(rc, "sdetr_salience_select")
    return inds, score, fg, order


def order_prefixes(tile_order, nq_list: Sequence[int]):
    """-> list of (b, nq_j) int32 processing orders, one per encoder layer."""
    b, K = tile_order.shape
    offs, tot = [], 0
    for n in nq_list:
        offs.append(tot)
        tot += b * int(n)
    out = torch.empty(max(tot, 1), device=tile_order.device, dtype=torch.int32)
    rc = lib().sdetr_order_prefixes(_req(tile_order, "tile_order", torch.int32), b, K, len(nq_list),
                                    _host_i32(nq_list), _host_i64(offs), out.data_ptr(), _stream())
    _check(rc, "sdetr_order_prefixes")
    return [out[o:o + b * int(n)].view(b, int(n)) for o, n in zip(offs, nq_list)]


def topk_desc(score, k: int, workspace=None):
    """(segments, n) -> (segments, k) int64 positions of the k largest (ties: smaller position first)."""
    seg, n = score.shape
    out = torch.empty(seg, k, device=score.device, dtype=torch.int64)
    need = lib().sdetr_topk_workspace(seg, n)
    if workspace is None or workspace.numel() < need:
        workspace = torch.empty(need, device=score.device, dtype=torch.uint8)
    rc = lib().sdetr_topk_desc(_req(score, "score", torch.float32), seg, n, k, out.data_ptr(), workspace.data_ptr(),
                               workspace.numel(), _stream())
    _check(rc, "sdetr_topk_desc")
    return out


# ---- token movement -----------------------------------------------------------------------------------------
def _inds_view(inds, num_query):
    if not (inds.is_cuda and inds.dtype == torch.int64 and inds.stride(1) == 1 and inds.shape[1] >= num_query):
        raise RuntimeError("inds must be a CUDA int64 tensor with unit column stride")
    return inds.data_ptr(), inds.stride(0)


def token_gather(tokens, pos, fg, valid_ratios, inds, spatial_shapes, level_start_index, num_query, want_sum=False):
    """-> (query, query_pos, fg_q, ref_q[, query + query_pos])"""
    b, nv, c = tokens.shape
    L = spatial_shapes.shape[0]
    dev = tokens.device
    q = torch.empty(b, num_query, c, device=dev, dtype=torch.float32)
    qp = torch.empty(b, num_query, c, device=dev, dtype=torch.float32)
    fq = torch.empty(b, num_query, device=dev, dtype=torch.float32)
    rq = torch.empty(b, num_query, L, 2, device=dev, dtype=torch.float32)
    qs = torch.empty(b, num_query, c, device=dev, dtype=torch.float32) if want_sum else None
    ip, istride = _inds_view(inds, num_query)
    rc = lib().sdetr_token_gather(
        _req(tokens, "tokens", torch.float32), _req(pos, "pos", torch.float32), _req(fg, "fg", torch.float32),
        _req(valid_ratios, "valid_ratios", torch.float32), ip, istride, _req(spatial_shapes, "spatial_shapes", torch.int64),
        _req(level_start_index, "level_start_index", torch.int64), b, nv, c, L, num_query, q.data_ptr(), qp.data_ptr(),
        fq.data_ptr(), rq.data_ptr(), qs.data_ptr() if want_sum else None, _stream())
    _check(rc, "sdetr_token_gather")
    return (q, qp, fq, rq, qs) if want_sum else (q, qp, fq, rq)


def token_scatter_(tokens, query, inds, focus_token_nums):
    b, nv, c = tokens.shape
    nq = query.shape[1]
    ip, istride = _inds_view(inds, nq)
    rc = lib().sdetr_token_scatter(_req(tokens, "tokens", torch.float32), _req(query, "query", torch.float32), ip,
                                   istride, _req(focus_token_nums, "focus_token_nums", torch.int32), b, nv, c, nq,
                                   _stream())
    _check(rc, "sdetr_token_scatter")
    return tokens


def background_embed_(tokens, mask_u8, last_inds, row_embed, col_embed, spatial_shapes, level_start_index, flags=None,
                      shapes_host: Optional[Sequence[Sequence[int]]] = None):
    """``shapes_host``: the [(H, W), ...] of the levels as host ints; a map larger than the embedding tables then raises
    like the reference's ``nn.Embedding`` lookup.  Without it the shapes are fetched from the device (one sync)."""
    b, nv, c = tokens.shape
    num_last = last_inds.shape[1]
    ip, istride = _inds_view(last_inds, num_last)
    if flags is None:
        flags = torch.empty(b, nv, device=tokens.device, dtype=torch.uint8)
    L = spatial_shapes.shape[0]
    if shapes_host is None:
        shapes_host = spatial_shapes.tolist()
    host = (ctypes.c_int64 * (2 * L))(*[int(v) for hw in shapes_host for v in hw])
    rc = lib().sdetr_background_embed(
        _req(tokens, "tokens", torch.float32), _req(mask_u8, "mask", torch.uint8), ip, istride, num_last,
        _req(row_embed, "row_embed", torch.float32), _req(col_embed, "col_embed", torch.float32),
        min(row_embed.shape[0], col_embed.shape[0]),
        _req(spatial_shapes, "spatial_shapes", torch.int64), ctypes.cast(host, ctypes.c_void_p),
        _req(level_start_index, "level_start_index", torch.int64),
        b, nv, c, L, _req(flags, "flags", torch.uint8), _stream())
    _check(rc, "sdetr_background_embed")
    return tokens


def score_modulate(mem_all, level_start: int, H: int, W: int, coarse_score, Hc: int, Wc: int, alpha, alpha_index: int):
    """mem_all (b,Nv,C) contiguous; the level slice starts at token ``level_start``.  coarse_score (b,Hc*Wc) rows
    (any batch stride).  -> (b,H*W,C)."""
    b, nv, c = mem_all.shape
    out = torch.empty(b, H * W, c, device=mem_all.device, dtype=torch.float32)
    if not (coarse_score.is_cuda and coarse_score.stride(1) == 1 and coarse_score.dtype == torch.float32):
        raise RuntimeError("coarse_score rows must be contiguous CUDA float32")
    if not 0 <= alpha_index < alpha.numel():  # the reference raises IndexError on alpha[lvl] (salience_transformer.py:143)
        raise IndexError(f"alpha has {alpha.numel()} entries, level {alpha_index} asked for")
    rc = lib().sdetr_score_modulate(
        _req(mem_all, "mem", torch.float32) + 4 * level_start * c, nv * c, coarse_score.data_ptr(),
        coarse_score.stride(0), _req(alpha, "alpha", torch.float32), alpha_index, b, H, W, Hc, Wc, c, out.data_ptr(),
        _stream())
    _check(rc, "sdetr_score_modulate")
    return out


def mask_predictor_level(mem_all, level_start: int, H: int, W: int, coarse_score, Hc: int, Wc: int, alpha, alpha_index: int,
                         ln_gamma, ln_beta, eps: float, w1_t, b1, w2a_t, b2a, w2b_t, b2b, w2c, b2c, raw, out_start: int):
    """Score modulation + MaskPredictor of one level in two launches.  mem_all (b,Nv,256) contiguous, the level's tokens start at
    ``level_start``; coarse_score: (b,Hc*Wc) rows of the next coarser level's raw scores (any batch stride) or None; weights
    transposed to (in, out); writes raw[:, out_start:out_start + H*W] (raw (b,Nv) contiguous)."""
    b, nv, c = mem_all.shape
    if coarse_score is not None:
        if not (coarse_score.is_cuda and coarse_score.stride(1) == 1 and coarse_score.dtype == torch.float32):
            raise RuntimeError("coarse_score rows must be contiguous CUDA float32")
        if not 0 <= alpha_index < alpha.numel():
            raise IndexError(f"alpha has {alpha.numel()} entries, level {alpha_index} asked for")
    if not (raw.is_cuda and raw.dtype == torch.float32 and raw.is_contiguous() and tuple(raw.shape) == (b, nv)):
        raise RuntimeError("raw must be a contiguous CUDA float32 (b, Nv) tensor")
    ws = torch.empty(int(lib().sdetr_mask_predictor_level_workspace_floats(b, H, W)), device=mem_all.device, dtype=torch.float32)
    rc = lib().sdetr_mask_predictor_level(
        _req(mem_all, "mem", torch.float32) + 4 * level_start * c, nv * c, b, H, W, c,
        coarse_score.data_ptr() if coarse_score is not None else None, coarse_score.stride(0) if coarse_score is not None else 0,
        Hc, Wc, _req(alpha, "alpha", torch.float32) if coarse_score is not None else None, alpha_index,
        _req(ln_gamma, "ln_gamma", torch.float32), _req(ln_beta, "ln_beta", torch.float32), float(eps),
        _req(w1_t, "w1_t", torch.float32), _req(b1, "b1", torch.float32), _req(w2a_t, "w2a_t", torch.float32),
        _req(b2a, "b2a", torch.float32), _req(w2b_t, "w2b_t", torch.float32), _req(b2b, "b2b", torch.float32),
        _req(w2c, "w2c", torch.float32), _req(b2c, "b2c", torch.float32), ws.data_ptr(), ws.numel(),
        raw.data_ptr() + 4 * out_start, nv, _stream())
    _check(rc, "sdetr_mask_predictor_level")
    return raw


def zero_masked_rows_(buf, row_stride: int, row_floats: int, mask_u8, num_rows: int, offset_floats: int = 0):
    rc = lib().sdetr_zero_masked_rows(buf.data_ptr() + 4 * offset_floats, row_stride, row_floats,
                                      _req(mask_u8, "mask", torch.uint8), num_rows, _stream())
    _check(rc, "sdetr_zero_masked_rows")
    return buf


def class_max_times_fg(logits, fg):
    """logits (..., num_classes) with unit last stride and a uniform row pitch (a padded GEMM output is fine)."""
    rows = fg.numel()
    out = torch.empty_like(fg)
    if not (logits.is_cuda and logits.dtype == torch.float32 and logits.stride(-1) == 1):
        raise RuntimeError("logits must be CUDA float32 with unit last stride")
    pitch = logits.stride(-2)
    if logits.dim() == 3 and logits.stride(0) != logits.shape[1] * pitch:
        logits, pitch = logits.contiguous(), logits.shape[-1]
    rc = lib().sdetr_class_max_times_fg(logits.data_ptr(), pitch, _req(fg, "fg", torch.float32), rows,
                                        logits.shape[-1], out.data_ptr(), _stream())
    _check(rc, "sdetr_class_max_times_fg")
    return out


def add_layernorm(x, r, gamma, beta, eps: float = 1e-5, out=None):
    """LayerNorm(x + r); r may be None; out may alias x."""
    c = x.shape[-1]
    rows = x.numel() // c
    if out is None:
        out = torch.empty_like(x)
    rc = lib().sdetr_add_layernorm(_req(x, "x", torch.float32), _req(r, "r", torch.float32) if r is not None else None,
                                   _req(gamma, "gamma", torch.float32), _req(beta, "beta", torch.float32), eps, rows, c,
                                   out.data_ptr(), _stream())
    _check(rc, "sdetr_add_layernorm")
    return out


def split_tf32(x, layout_b: bool = False, relu=0, chunk: int = 0):
    """(..., K) fp32 with contiguous rows -> (rows, 3K) 3xTF32 operand: per K-chunk [hi|hi|lo] (weights: [hi|lo|hi])."""
    K = x.shape[-1]
    chunk = chunk or K
    if not (x.is_cuda and x.dtype == torch.float32 and x.stride(-1) == 1):
        raise RuntimeError("split_tf32 needs a CUDA float32 tensor with unit last stride")
    x2 = x if x.dim() == 2 else x.reshape(-1, K)  # view when rows are uniformly strided
    rows = x2.shape[0]
    out = torch.empty(rows, 3 * K, device=x.device, dtype=torch.float32)
    rc = lib().sdetr_split_tf32(x2.data_ptr(), x2.stride(0) if rows > 1 else K, rows, K, chunk, int(layout_b), int(relu),
                                out.data_ptr(), _stream())
    _check(rc, "sdetr_split_tf32")
    return out


def rows_gather(src, index):
    """src (b,n,C), index (b,k) int64 -> (b,k,C)."""
    b, n, c = src.shape
    k = index.shape[1]
    out = torch.empty(b, k, c, device=src.device, dtype=torch.float32)
    rc = lib().sdetr_rows_gather(_req(src, "src", torch.float32), _req(index, "index", torch.int64), b, n, k, c,
                                 out.data_ptr(), _stream())
    _check(rc, "sdetr_rows_gather")
    return out


def rows_scatter_(dst, index, src):
    """dst (b,n,C) <- src (b,k,C) at index (b,k) int64, in place."""
    b, n, c = dst.shape
    k = index.shape[1]
    rc = lib().sdetr_rows_scatter(_req(dst, "dst", torch.float32), _req(index, "index", torch.int64), b, n, k, c,
                                  _req(src, "src", torch.float32), _stream())
    _check(rc, "sdetr_rows_scatter")
    return dst


def split_tf32_pair(w):
    """(N,K) fp32 -> (W_hi, W_lo), both TF32-representable fp32 tensors of the same shape."""
    w = w.contiguous()
    hi, lo = torch.empty_like(w), torch.empty_like(w)
    _check(lib().sdetr_split_tf32_pair(_req(w, "w", torch.float32), w.numel(), hi.data_ptr(), lo.data_ptr(), _stream()),
           "sdetr_split_tf32_pair")
    return hi, lo


def gemm_3xtf32(x, w_hi, w_lo, bias=None, relu_input=0):
    """y = act(x) @ W.T + bias on the tcgen05 tensor cores; x (..., K) with unit last stride and uniform row pitch."""
    K = x.shape[-1]
    N = w_hi.shape[0]
    if not (x.is_cuda and x.dtype == torch.float32 and x.stride(-1) == 1):
        raise RuntimeError("gemm_3xtf32 needs a CUDA float32 input with unit last stride")
    x2 = x if x.dim() == 2 else x.reshape(-1, K)
    M = x2.shape[0]
    y = torch.empty(M, N, device=x.device, dtype=torch.float32)
    rc = lib().sdetr_gemm_3xtf32(x2.data_ptr(), x2.stride(0) if M > 1 else K, _req(w_hi, "w_hi", torch.float32),
                                 _req(w_lo, "w_lo", torch.float32),
                                 _req(bias, "bias", torch.float32) if bias is not None else None, y.data_ptr(), N, M, N, K,
                                 int(relu_input), _stream())
    _check(rc, "sdetr_gemm_3xtf32")
    return y.view(*x.shape[:-1], N)


def flatten_tokens(feats, pos, level_embeds, keep):
    """Per-level (b,C,H,W) feats / pos -> (feat_tok, lpos_tok, x_tok), each (b,Nv,C); keep (b,Nv) float."""
    L = len(feats)
    b, c = feats[0].shape[:2]
    sizes = [f.shape[2] * f.shape[3] for f in feats]
    nv = sum(sizes)
    dev = feats[0].device
    fp = (ctypes.c_void_p * L)(*[_req(f, "feat", torch.float32) for f in feats])
    pp = (ctypes.c_void_p * L)(*[_req(p, "pos", torch.float32) for p in pos])
    out = [torch.empty(b, nv, c, device=dev, dtype=torch.float32) for _ in range(3)]
    rc = lib().sdetr_flatten_tokens(fp, pp, _req(level_embeds, "level_embeds", torch.float32), _req(keep, "keep", torch.float32),
                                    _host_i32(sizes), b, c, L, out[0].data_ptr(), out[1].data_ptr(), out[2].data_ptr(), _stream())
    _check(rc, "sdetr_flatten_tokens")
    return out


def flatten_tokens_pos(feats, pos_tokens, level_embeds, keep):
    """Per-level (b,C,H,W) feats + token-layout position embedding (b,Nv,C) -> (feat_tok, lpos_tok, x_tok)."""
    L = len(feats)
    b, c = feats[0].shape[:2]
    sizes = [f.shape[2] * f.shape[3] for f in feats]
    nv = sum(sizes)
    fp = (ctypes.c_void_p * L)(*[_req(f, "feat", torch.float32) for f in feats])
    out = [torch.empty(b, nv, c, device=feats[0].device, dtype=torch.float32) for _ in range(3)]
    if tuple(pos_tokens.shape) != (b, nv, c):
        raise RuntimeError(f"pos_tokens must be {(b, nv, c)}, got {tuple(pos_tokens.shape)}")
    rc = lib().sdetr_flatten_tokens_pos(fp, _req(pos_tokens, "pos_tokens", torch.float32),
                                        _req(level_embeds, "level_embeds", torch.float32), _req(keep, "keep", torch.float32),
                                        _host_i32(sizes), b, c, L, out[0].data_ptr(), out[1].data_ptr(), out[2].data_ptr(), _stream())
    _check(rc, "sdetr_flatten_tokens_pos")
    return out


def salience_targets(boxes_xyxy, num_boxes, shapes, strides, limit_range):
    """boxes (b,max_boxes,4) xyxy pixels (padded), num_boxes (b,) int32 -> (b,Nv) salience targets (noise-free)."""
    b, mb = boxes_xyxy.shape[:2]
    nv = sum(h * w for h, w in shapes)
    out = torch.empty(b, nv, device=num_boxes.device, dtype=torch.float32)
    f = lambda xs: (ctypes.c_float * len(xs))(*[float(x) for x in xs])  # noqa: E731
    rc = lib().sdetr_salience_targets(_req(boxes_xyxy, "boxes", torch.float32) if mb else None, _req(num_boxes, "num_boxes", torch.int32),
                                      mb, b, nv, len(shapes), _host_i32([h for h, _ in shapes]), _host_i32([w for _, w in shapes]),
                                      f([s[0] for s in strides]), f([s[1] for s in strides]), f([r[0] for r in limit_range]),
                                      f([r[1] for r in limit_range]), out.data_ptr(), _stream())
    _check(rc, "sdetr_salience_targets")
    return out


def tokens_to_maps(tokens, shapes):
    """(b,Nv,C) tokens -> list of per-level (b,C,H,W) maps (one launch; salience_transformer.py:186-190)."""
    b, nv, c = tokens.shape
    maps = [torch.empty(b, c, h, w, device=tokens.device, dtype=torch.float32) for h, w in shapes]
    ptrs = (ctypes.c_void_p * len(shapes))(*[m.data_ptr() for m in maps])
    rc = lib().sdetr_token_map_transpose(_req(tokens, "tokens", torch.float32), ptrs, _host_i32([h * w for h, w in shapes]), b, c,
                                         len(shapes), 1, _stream())
    _check(rc, "sdetr_token_map_transpose")
    return maps


def maps_to_tokens(maps):
    """list of per-level (b,C,H,W) maps -> (b,Nv,C) tokens (one launch; salience_transformer.py:192)."""
    b, c = maps[0].shape[:2]
    sizes = [m.shape[2] * m.shape[3] for m in maps]
    tokens = torch.empty(b, sum(sizes), c, device=maps[0].device, dtype=torch.float32)
    ptrs = (ctypes.c_void_p * len(maps))(*[_req(m, "map", torch.float32) for m in maps])
    rc = lib().sdetr_token_map_transpose(tokens.data_ptr(), ptrs, _host_i32(sizes), b, c, len(maps), 0, _stream())
    _check(rc, "sdetr_token_map_transpose")
    return tokens


def mask_plan(mask_u8, shapes, level_filter_ratio, pos_offset=-0.5, pos_eps=1e-6, pos_scale=2 * 3.141592653589793):
    """mask_u8 (b,Nv) -> dict(valid (b,L) i32, focus (b,L) i32, valid_ratios (b,L,2), keep (b,Nv), ynorm, xnorm (b,Nv))."""
    b, nv = mask_u8.shape
    L = len(shapes)
    dev = mask_u8.device
    f32 = dict(device=dev, dtype=torch.float32)
    out = dict(valid=torch.empty(b, L, device=dev, dtype=torch.int32), focus=torch.empty(b, L, device=dev, dtype=torch.int32),
               valid_ratios=torch.empty(b, L, 2, **f32), keep=torch.empty(b, nv, **f32), ynorm=torch.empty(b, nv, **f32),
               xnorm=torch.empty(b, nv, **f32))
    ratios = (ctypes.c_float * L)(*[float(r) for r in level_filter_ratio])
    rc = lib().sdetr_mask_plan(_req(mask_u8, "mask", torch.uint8), b, nv, L, _host_i32([h for h, _ in shapes]),
                               _host_i32([w for _, w in shapes]), ratios, float(pos_offset), float(pos_eps), float(pos_scale),
                               out["ynorm"].data_ptr(), out["xnorm"].data_ptr(), out["valid"].data_ptr(), out["focus"].data_ptr(),
                               out["valid_ratios"].data_ptr(), out["keep"].data_ptr(), _stream())
    _check(rc, "sdetr_mask_plan")
    return out


def sine_pos_tokens(ynorm, xnorm, dim_ty, dim_tx):
    """(b,Nv) normalised coordinates -> (b,Nv,2F) sine position embedding in token layout."""
    b, nv = ynorm.shape
    f = dim_ty.numel()
    pos = torch.empty(b, nv, 2 * f, device=ynorm.device, dtype=torch.float32)
    rc = lib().sdetr_sine_pos_tokens(_req(ynorm, "ynorm", torch.float32), _req(xnorm, "xnorm", torch.float32),
                                     _req(dim_ty, "dim_ty", torch.float32), _req(dim_tx, "dim_tx", torch.float32), b * nv, f,
                                     pos.data_ptr(), _stream())
    _check(rc, "sdetr_sine_pos_tokens")
    return pos


def gemm_3xtf32_raw(x, w, bias=None, act=0):
    """y = act(x) @ w.T + bias, both operands split inside the kernel (w: raw fp32 (N,K) contiguous)."""
    K = x.shape[-1]
    N = w.shape[0]
    if not (x.is_cuda and x.dtype == torch.float32 and x.stride(-1) == 1):
        raise RuntimeError("gemm_3xtf32_raw needs a CUDA float32 input with unit last stride")
    x2 = x if x.dim() == 2 else x.reshape(-1, K)
    M = x2.shape[0]
    ldc = (N + 3) // 4 * 4  # 16-byte row pitch lets the epilogue leave through TMA bulk stores (e.g. N = 91 -> 92)
    y = torch.empty(M, ldc, device=x.device, dtype=torch.float32)
    rc = lib().sdetr_gemm_3xtf32_raw(x2.data_ptr(), x2.stride(0) if M > 1 else K, _req(w, "w", torch.float32),
                                     _req(bias, "bias", torch.float32) if bias is not None else None, y.data_ptr(), ldc, M, N, K,
                                     int(act), _stream())
    _check(rc, "sdetr_gemm_3xtf32_raw")
    y = y if ldc == N else y[:, :N]
    return y.reshape(*x.shape[:-1], N) if ldc == N else y.unflatten(0, x.shape[:-1])


def gemm_3xtf32_pre(x, w_hi, w_lo, bias=None, act=0):
    """y = act(x) @ (w_hi + w_lo).T + bias on the persistent kernel with a pre-split weight (sdetr_split_tf32_pair)."""
    K = x.shape[-1]
    N = w_hi.shape[0]
    if not (x.is_cuda and x.dtype == torch.float32 and x.stride(-1) == 1):
        raise RuntimeError("gemm_3xtf32_pre needs a CUDA float32 input with unit last stride")
    x2 = x if x.dim() == 2 else x.reshape(-1, K)
    M = x2.shape[0]
    ldc = (N + 3) // 4 * 4
    y = torch.empty(M, ldc, device=x.device, dtype=torch.float32)
    rc = lib().sdetr_gemm_3xtf32_pre(x2.data_ptr(), x2.stride(0) if M > 1 else K, _req(w_hi, "w_hi", torch.float32),
                                     _req(w_lo, "w_lo", torch.float32),
                                     _req(bias, "bias", torch.float32) if bias is not None else None, y.data_ptr(), ldc, M, N, K,
                                     int(act), _stream())
    _check(rc, "sdetr_gemm_3xtf32_pre")
    y = y if ldc == N else y[:, :N]
    return y.reshape(*x.shape[:-1], N) if ldc == N else y.unflatten(0, x.shape[:-1])


def split_f16_pair(w):
    """(N,K) fp32 -> (W_hi, W_lo, scale): fp16 pair of scale * W with scale = 2^s such that max|scale * W| is in
    [2^13, 2^14) (one host read of max|W|; done once per parameter version by gemm.split_weight_f16)."""
    import math
    w = w.contiguous()
    amax = float(w.abs().max())
    if not math.isfinite(amax):
        raise RuntimeError("split_f16_pair: weight has non-finite entries")
    scale = 2.0 ** (13 - math.frexp(amax)[1] + 1) if amax > 0 else 1.0  # frexp: amax = m * 2^e, 0.5 <= m < 1
    hi = torch.empty(w.shape, device=w.device, dtype=torch.float16)
    lo = torch.empty_like(hi)
    _check(lib().sdetr_split_f16_pair(_req(w, "w", torch.float32), w.numel(), scale, hi.data_ptr(), lo.data_ptr(), _stream()),
           "sdetr_split_f16_pair")
    return hi, lo, scale


def gemm_f16x3_pre(x, w_hi, w_lo, w_scale, bias=None, act=0, out=None):
    """y = act(x) @ W.T + bias on the persistent tcgen05.mma.kind::f16 kernel (3xFP16, fp32-class accuracy).
    ``out``: optional (M, N) fp32 destination with unit column stride (e.g. a row slice of a larger buffer)."""
    K = x.shape[-1]
    N = w_hi.shape[0]
    if not (x.is_cuda and x.dtype == torch.float32 and x.stride(-1) == 1):
        raise RuntimeError("gemm_f16x3_pre needs a CUDA float32 input with unit last stride")
    x2 = x if x.dim() == 2 else x.reshape(-1, K)
    M = x2.shape[0]
    if out is not None:
        if not (out.is_cuda and out.dtype == torch.float32 and out.dim() == 2 and tuple(out.shape) == (M, N) and out.stride(1) == 1):
            raise RuntimeError(f"gemm_f16x3_pre: out must be a CUDA float32 ({M}, {N}) tensor with unit column stride")
        rc = lib().sdetr_gemm_f16x3_pre(x2.data_ptr(), x2.stride(0) if M > 1 else K, _req(w_hi, "w_hi", torch.float16),
                                        _req(w_lo, "w_lo", torch.float16), float(w_scale),
                                        _req(bias, "bias", torch.float32) if bias is not None else None, out.data_ptr(),
                                        out.stride(0) if M > 1 else N, M, N, K, int(act), _stream())
        _check(rc, "sdetr_gemm_f16x3_pre")
        return out
    ldc = (N + 3) // 4 * 4
    y = torch.empty(M, ldc, device=x.device, dtype=torch.float32)
    rc = lib().sdetr_gemm_f16x3_pre(x2.data_ptr(), x2.stride(0) if M > 1 else K, _req(w_hi, "w_hi", torch.float16),
                                    _req(w_lo, "w_lo", torch.float16), float(w_scale),
                                    _req(bias, "bias", torch.float32) if bias is not None else None, y.data_ptr(), ldc, M, N, K,
                                    int(act), _stream())
    _check(rc, "sdetr_gemm_f16x3_pre")
    y = y if ldc == N else y[:, :N]
    return y.reshape(*x.shape[:-1], N) if ldc == N else y.unflatten(0, x.shape[:-1])


def ffn_fused_layernorm(x, w1, b1, w2, b2, gamma=None, beta=None, eps: float = 1e-5, out=None):
    """y = LayerNorm(x + linear2(ReLU(linear1(x)))) by the fused tcgen05 kernel (gamma None: the bare FFN, no residual / norm).
    ``w1`` / ``w2``: (W_hi, W_lo, scale) triples of ``split_f16_pair``; x (..., 256) fp32 with unit last stride."""
    if not (x.is_cuda and x.dtype == torch.float32 and x.shape[-1] == 256 and x.stride(-1) == 1):
        raise RuntimeError("ffn_fused_layernorm needs a CUDA float32 (..., 256) input with unit last stride")
    x2 = x if x.dim() == 2 else x.reshape(-1, 256)
    M = x2.shape[0]
    hidden = w1[0].shape[0]
    if tuple(w1[0].shape) != (hidden, 256) or tuple(w2[0].shape) != (256, hidden):
        raise RuntimeError("ffn_fused_layernorm: W1 must be (hidden, 256) and W2 (256, hidden)")
    y = out if out is not None else torch.empty(M, 256, device=x.device, dtype=torch.float32)
    if not (y.is_cuda and y.dtype == torch.float32 and y.is_contiguous() and y.numel() == M * 256):
        raise RuntimeError("ffn_fused_layernorm: out must be a contiguous CUDA float32 tensor of the input's size")
    ws = torch.empty(max(int(lib().sdetr_ffn_fused_workspace_floats(M, hidden)), 4), device=x.device, dtype=torch.float32)
    rc = lib().sdetr_ffn_fused_layernorm(
        x2.data_ptr(), x2.stride(0) if M > 1 else 256, _req(w1[0], "w1_hi", torch.float16), _req(w1[1], "w1_lo", torch.float16),
        float(w1[2]), _req(b1, "b1", torch.float32), _req(w2[0], "w2_hi", torch.float16), _req(w2[1], "w2_lo", torch.float16),
        float(w2[2]), _req(b2, "b2", torch.float32) if b2 is not None else None,
        _req(gamma, "gamma", torch.float32) if gamma is not None else None,
        _req(beta, "beta", torch.float32) if beta is not None else None, float(eps), M, hidden, ws.data_ptr(), ws.numel(),
        y.data_ptr(), _stream())
    _check(rc, "sdetr_ffn_fused_layernorm")
    return y if out is not None else y.reshape(x.shape)


_pow2_state = {}


def pow2_scale(x, target_log2: int = 12):
    """Device scalar 2^s with max|x| * 2^s in [2^(target_log2 - 1), 2^target_log2); no host synchronisation."""
    if not (x.is_cuda and x.dtype == torch.float32 and x.is_contiguous()):
        raise RuntimeError("pow2_scale needs a contiguous CUDA float32 tensor")
    key = (x.device.index, _stream())
    st = _pow2_state.get(key)
    if st is None:
        st = _pow2_state[key] = torch.zeros(2, device=x.device, dtype=torch.int32)  # per (device, stream): calls are stream-ordered
    out = torch.empty(1, device=x.device, dtype=torch.float32)
    _check(lib().sdetr_pow2_scale(x.data_ptr(), x.numel(), int(target_log2), st.data_ptr(), out.data_ptr(), _stream()), "sdetr_pow2_scale")
    return out


def split_f16_pair_dev(w, scale=None):
    """(W_hi, W_lo, scale) with the power-of-two scale computed and kept ON THE DEVICE (training: the weights change every step)."""
    w = w if w.is_contiguous() else w.contiguous()
    scale = pow2_scale(w, 14) if scale is None else scale
    hi = torch.empty(w.shape, device=w.device, dtype=torch.float16)
    lo = torch.empty_like(hi)
    _check(lib().sdetr_split_f16_pair_dev(_req(w, "w", torch.float32), w.numel(), scale.data_ptr(), hi.data_ptr(), lo.data_ptr(), _stream()),
           "sdetr_split_f16_pair_dev")
    return hi, lo, scale


def gemm_f16x3_scaled(x, a_scale, w_hi, w_lo, w_scale, bias=None):
    """y = x @ W.T + bias; ``a_scale`` / ``w_scale``: device scalars (pow2_scale / split_f16_pair_dev) -- no host synchronisation."""
    K = x.shape[-1]
    N = w_hi.shape[0]
    if not (x.is_cuda and x.dtype == torch.float32 and x.stride(-1) == 1):
        raise RuntimeError("gemm_f16x3_scaled needs a CUDA float32 input with unit last stride")
    x2 = x if x.dim() == 2 else x.reshape(-1, K)
    M = x2.shape[0]
    ldc = (N + 3) // 4 * 4
    # the result must not be a view (autograd Functions hand it out, and in-place ops on it are legal): final shape when N % 4 == 0
    y = torch.empty(*x.shape[:-1], N, device=x.device, dtype=torch.float32) if ldc == N else torch.empty(M, ldc, device=x.device, dtype=torch.float32)
    rc = lib().sdetr_gemm_f16x3_scaled(x2.data_ptr(), x2.stride(0) if M > 1 else K, _req(a_scale, "a_scale", torch.float32),
                                       _req(w_hi, "w_hi", torch.float16), _req(w_lo, "w_lo", torch.float16),
                                       _req(w_scale, "w_scale", torch.float32),
                                       _req(bias, "bias", torch.float32) if bias is not None else None, y.data_ptr(), ldc, M, N, K,
                                       _stream())
    _check(rc, "sdetr_gemm_f16x3_scaled")
    return y if ldc == N else y[:, :N].contiguous().reshape(*x.shape[:-1], N)


def rows_gather_add(src, pos, index):
    """t = src[b, index], x = t + pos[b, index]  (the pre-attention's two gathers and the `with_pos_embed` add)."""
    b, n, c = src.shape
    k = index.shape[1]
    t = torch.empty(b, k, c, device=src.device, dtype=torch.float32)
    x = torch.empty_like(t)
    rc = lib().sdetr_rows_gather_add(_req(src, "src", torch.float32), _req(pos, "pos", torch.float32),
                                     _req(index, "index", torch.int64), b, n, k, c, t.data_ptr(), x.data_ptr(), _stream())
    _check(rc, "sdetr_rows_gather_add")
    return t, x


def attention_small(qk, v):
    """qk (b,n,2,h,d) [queries | keys], v (b,n,h,d) -> (b,n,h*d) = softmax(QK^T/sqrt(d)) V per head."""
    b, n, _, h, d = qk.shape
    out = torch.empty(b, n, h * d, device=qk.device, dtype=torch.float32)
    rc = lib().sdetr_attention_small(_req(qk, "qk", torch.float32), _req(v, "v", torch.float32), out.data_ptr(), b, n, h, d,
                                     _stream())
    _check(rc, "sdetr_attention_small")
    return out


def attention_qkv(qkv, heads: int):
    """qkv (b,n,3C) packed [q | k | v] projections -> (b,n,C) = softmax(QK^T/sqrt(d)) V per head (d = C/heads = 32)."""
    b, n, c3 = qkv.shape
    c = c3 // 3
    out = torch.empty(b, n, c, device=qkv.device, dtype=torch.float32)
    rc = lib().sdetr_attention_qkv(_req(qkv, "qkv", torch.float32), out.data_ptr(), b, n, heads, c // heads, _stream())
    _check(rc, "sdetr_attention_qkv")
    return out


def mha_in_proj(tokens, pos, index, w_in_t, b_in):
    """tokens/pos (b,nq,C), index (b,k) int64, w_in_t (C,3C) = in_proj_weight.T -> (t (b,k,C), qkv (b,k,3C))."""
    b, nq, c = tokens.shape
    k = index.shape[1]
    t = torch.empty(b, k, c, device=tokens.device, dtype=torch.float32)
    qkv = torch.empty(b, k, 3 * c, device=tokens.device, dtype=torch.float32)
    if tuple(w_in_t.shape) != (c, 3 * c) or b_in.numel() != 3 * c:
        raise RuntimeError("w_in_t must be (C, 3C) and b_in (3C)")
    rc = lib().sdetr_mha_in_proj(_req(tokens, "tokens", torch.float32), _req(pos, "pos", torch.float32),
                                 _req(index, "index", torch.int64), b, nq, k, c, _req(w_in_t, "w_in_t", torch.float32),
                                 _req(b_in, "b_in", torch.float32), t.data_ptr(), qkv.data_ptr(), _stream())
    _check(rc, "sdetr_mha_in_proj")
    return t, qkv


def mha_out_proj_ln_scatter_(dst, attn, t, w_out_t, b_out, gamma, beta, eps: float, index, pos=None, dst_sum=None):
    """dst (b,nq,C) <- LayerNorm(t + attn @ w_out_t + b_out) at index (b,k), in place; with pos / dst_sum (b,nq,C) the
    same rows of dst_sum receive the new row + pos."""
    b, nq, c = dst.shape
    k = index.shape[1]
    if tuple(w_out_t.shape) != (c, c):
        raise RuntimeError("w_out_t must be (C, C)")
    rc = lib().sdetr_mha_out_proj_ln_scatter(
        _req(attn, "attn", torch.float32), _req(t, "t", torch.float32), _req(w_out_t, "w_out_t", torch.float32),
        _req(b_out, "b_out", torch.float32), _req(gamma, "gamma", torch.float32), _req(beta, "beta", torch.float32),
        float(eps), _req(index, "index", torch.int64), _req(dst, "dst", torch.float32),
        _req(pos, "pos", torch.float32) if pos is not None else None,
        _req(dst_sum, "dst_sum", torch.float32) if dst_sum is not None else None, b, nq, k, c, _stream())
    _check(rc, "sdetr_mha_out_proj_ln_scatter")
    return dst


def gelu_colmean_(z, half: int):
    """z (b,n,C) in place: GELU everywhere, then columns [half, C) replaced by their per-image token mean."""
    b, n, c = z.shape
    need = lib().sdetr_gelu_colmean_workspace(b, n, c, half)
    ws = torch.empty(need, device=z.device, dtype=torch.uint8)
    rc = lib().sdetr_gelu_colmean(_req(z, "z", torch.float32), b, n, c, half, ws.data_ptr(), ws.numel(), _stream())
    _check(rc, "sdetr_gelu_colmean")
    return z
