"""ctypes binding of libpa_b200.so (include/pa_b200.h).

torch tensors only supply device pointers (`.data_ptr()`) and the current stream; every
computation on the path happens inside the hand-written sm_100a kernels.  There is no
fallback: a missing or unloadable library raises at first use.
"""
import ctypes as C
import os

_PKG = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.path.join(_PKG, "libpa_b200.so")

PA_OK = 0
PA_ERR_UNSUPPORTED = -2
ACT = {"": 0, None: 0, "none": 0, "relu": 1, "gelu": 2}

_lib = None

_vp, _i32, _i64, _f32, _sz = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_size_t

_DECODE_COMMON = [_vp, _i32, _i32, _i32, _i32, _vp, _vp, _i32, _i32, _i32, _i32, _f32, _vp]

_SIGS = {
    "pa_version": ([], _i32),
    "pa_error_string": ([_i32], C.c_char_p),
    "pa_device_info": ([C.POINTER(_i32)] * 3, _i32),
    "pa_page_table_clear": ([_vp, _i64, _vp], _i32),
    "pa_page_table_update": ([_vp, _i64, _vp, _vp, _i32, _vp], _i32),
    "pa_page_table_lookup": ([_vp, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _i32, _vp], _i32),
    "pa_kv_gather": ([_vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _i32, _i32, _vp], _i32),
    "pa_kv_append_f16": ([_vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _i32, _vp], _i32),
    "pa_kv_append_f32": ([_vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _i32, _vp], _i32),
    "pa_kv_append_f32_f16": ([_vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _i32, _vp], _i32),
    "pa_kv_append_f32_i8": ([_vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _i32, _vp], _i32),
    "pa_kv_copy_pages": ([_vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp], _i32),
    "pa_decode_workspace_bytes": ([_i32, _i32, _i32, _i32, _i32], _sz),
    "pa_paged_decode_f16": ([_vp, _vp, _vp, _vp] + _DECODE_COMMON + [_vp, _vp, _sz, _vp], _i32),
    "pa_paged_decode_f16_overlap": ([_vp, _vp, _vp, _vp] + _DECODE_COMMON + [_vp, _vp, _sz, _vp], _i32),
    "pa_paged_decode_i8": ([_vp, _vp, _vp, _vp, _vp, _vp] + _DECODE_COMMON + [_vp, _vp, _sz, _vp], _i32),
    "pa_paged_decode_i8_overlap": ([_vp, _vp, _vp, _vp, _vp, _vp] + _DECODE_COMMON + [_vp, _vp, _sz, _vp], _i32),
    "pa_paged_decode_f32": ([_vp, _vp, _vp, _vp] + _DECODE_COMMON + [_vp, _vp, _sz, _vp], _i32),
    "pa_paged_decode_f32_overlap": ([_vp, _vp, _vp, _vp] + _DECODE_COMMON + [_vp, _vp, _sz, _vp], _i32),
    "pa_attention_filtered_workspace_bytes": ([_i32, _i32, _i32], _sz),
    "pa_paged_attention_filtered": ([_vp, _vp, _vp, _vp, _vp, _vp, _i32] + _DECODE_COMMON + [_i32, _f32, _vp, _vp, _vp, _sz, _vp], _i32),
    "pa_paged_decode_f16_partial": ([_vp, _vp, _vp, _vp, _vp, _vp] + _DECODE_COMMON + [_vp, _sz, _vp], _i32),
    "pa_paged_decode_i8_partial": ([_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp] + _DECODE_COMMON + [_vp, _sz, _vp], _i32),
    "pa_paged_decode_f16_splitkv": ([_vp, _vp, _vp, _vp] + _DECODE_COMMON + [_vp, _vp, _sz, _vp, _i32, _i32, _vp, _vp, _vp], _i32),
    "pa_paged_decode_i8_splitkv": ([_vp, _vp, _vp, _vp, _vp, _vp] + _DECODE_COMMON + [_vp, _vp, _sz, _vp, _i32, _i32, _vp, _vp, _vp], _i32),
    "pa_paged_decode_f16_group": ([_vp, _vp, _vp, _vp] + _DECODE_COMMON + [_i32, _vp, _vp, _sz, _vp], _i32),
    "pa_paged_decode_i8_group": ([_vp, _vp, _vp, _vp, _vp, _vp] + _DECODE_COMMON + [_i32, _vp, _vp, _sz, _vp], _i32),
    "pa_softmax_lut_i32": ([_vp, _i32, _i32, _f32, _vp, _i32, _vp, _vp], _i32),
    "pa_softmax_temperature": ([_vp, _i32, _i32, _f32, _vp, _vp], _i32),
    "pa_topk_topp_filter": ([_vp, _i32, _i32, _i32, _f32, _i32, _f32, _vp], _i32),
    "pa_sample_from_probs": ([_vp, _i32, _i32, _vp, _vp, _vp], _i32),
    "pa_prefill_workspace_bytes": ([_i32, _i32, _i32, _i32, _i32, _i32], _sz),
    "pa_paged_prefill_f16": ([_vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp, _vp, _i32, _i32, _i32, _i32, _f32, _vp, _sz, _vp], _i32),
    "pa_paged_prefill_i8": ([_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp, _vp, _i32, _i32, _i32, _i32, _f32, _vp, _sz, _vp], _i32),
    "pa_paged_prefill_f16_tokmajor": ([_vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp, _vp, _i32, _i32, _i32, _i32, _f32, _vp], _i32),
    "pa_paged_prefill_i8_tokmajor": ([_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp, _vp, _i32, _i32, _i32, _i32, _f32, _vp], _i32),
    "pa_lse_combine": ([_vp, _vp, _vp, _i32, _i32, _i32, _vp, _vp, _vp], _i32),
    "pa_quantize_i8": ([_vp, _i64, _f32, _vp, _vp], _i32),
    "pa_batch_quantize_i8": ([_vp, _vp, _i32, _i32, _vp, _vp], _i32),
    "pa_absmax": ([_vp, _i64, _vp, _vp], _i32),
    "pa_minmax_scale": ([_vp, _i64, _vp, _vp], _i32),
    "pa_batch_minmax_scale": ([_vp, _i32, _i32, _vp, _vp], _i32),
    "pa_dequantize_i8": ([_vp, _i64, _f32, _vp, _vp], _i32),
    "pa_batch_dequantize_i8": ([_vp, _vp, _i32, _i32, _vp, _vp], _i32),
    "pa_gemm_i8_workspace_bytes": ([_i32, _i32, _i32, _i32], _sz),
    "pa_gemm_i8": ([_vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _f32, _f32, _f32, _vp, _i32, _vp, _sz, _vp], _i32),
    "pa_gemm_i8_dequant": ([_vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp, _f32, _vp, _i32, _vp, _sz, _vp], _i32),
    "pa_gemm_i8_dynquant_workspace_bytes": ([_i32, _i32, _i32], _sz),
    "pa_gemm_i8_dynquant": ([_vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp, _f32, _vp, _i32, _vp, _sz, _vp], _i32),
    "pa_embedding_f32": ([_vp, _vp, _i32, _i32, _i32, _vp, _vp], _i32),
    "pa_embedding_i8": ([_vp, _f32, _vp, _i32, _i32, _i32, _vp, _vp], _i32),
    "pa_layer_norm_f32": ([_vp, _vp, _vp, _i32, _i32, _f32, _vp, _vp], _i32),
    "pa_linear_workspace_bytes": ([_i32, _i32, _i32], _sz),
    "pa_linear_f32": ([_vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp, _vp, _sz, _vp], _i32),
    "pa_linear_pack_bytes": ([_i32, _i32], _sz),
    "pa_linear_pack_f32": ([_vp, _vp, _i32, _i32, _vp], _i32),
    "pa_linear_f32_packed": ([_vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp, _vp, _sz, _vp], _i32),
    "pa_logits_f32": ([_vp, _vp, _i32, _i32, _i32, _vp, _vp], _i32),
    "pa_logits_i8": ([_vp, _vp, _f32, _i32, _i32, _i32, _vp, _vp], _i32),
    "pa_argmax_f32": ([_vp, _i32, _i32, _f32, _i32, _vp, _vp], _i32),
    "pa_logits_argmax": ([_vp, _vp, _i32, _f32, _i32, _i32, _i32, _f32, _i32, _vp, _vp, _vp, _vp], _i32),
    "pa_layer_norm_quantize_i8": ([_vp, _vp, _vp, _i32, _i32, _f32, _vp, _vp, _vp, _vp], _i32),
    "pa_row_quantize_dynamic_i8": ([_vp, _i32, _i32, _vp, _vp, _vp], _i32),
    "pa_apply_rope_f32": ([_vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp], _i32),
    "pa_advance_positions": ([_vp, _vp, _i32, _vp], _i32),
    "pa_splitkv_exchange_bytes": ([_i32, _i32, _i32], _sz),
    "pa_p2p_alloc": ([_sz, C.POINTER(_vp), C.c_char_p], _i32),
    "pa_p2p_open": ([C.c_char_p, C.POINTER(_vp)], _i32),
    "pa_p2p_close": ([_vp], _i32),
    "pa_p2p_free": ([_vp], _i32),
    "pa_splitkv_exchange_combine": ([_vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp], _i32),
    "pa_splitkv_exchange_send": ([_vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp, _vp], _i32),
    "pa_splitkv_exchange_recv": ([_vp, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp], _i32),
    "pa_nccl_unique_id": ([C.c_char_p], _i32),
    "pa_nccl_init": ([C.c_char_p, _i32, _i32, C.POINTER(_vp)], _i32),
    "pa_nccl_destroy": ([_vp], _i32),
    "pa_nccl_gather_bytes": ([_i32, _i32, _i32], _sz),
    "pa_nccl_allgather_combine": ([_vp, _i32, _vp, _vp, _vp, _i32, _i32, _vp, _sz, _vp, _vp, _vp], _i32),
}

EXPORTS = tuple(_SIGS)


class PAError(RuntimeError):
    """A non-zero status from libpa_b200.so (mirrors the reference's C++ exception ->
    pybind11 -> RuntimeError path, SURVEY 8b 'Errors')."""


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise PAError(
                f"{LIB_PATH} is missing: build it with `python build.py` in {_PKG} "
                "(there is no CPU or PyTorch fallback for this path)")
        handle = C.CDLL(LIB_PATH)
        for name, (args, res) in _SIGS.items():
            fn = getattr(handle, name)  # AttributeError if the header and the .so disagree
            fn.argtypes = args
            fn.restype = res
        _lib = handle
    return _lib


def check(status, what=""):
    if status != PA_OK:
        msg = lib().pa_error_string(status).decode()
        raise PAError(f"{what or 'pa_b200'} failed: {msg} (status {status})")


def ptr(t):
    """Device (or host) pointer of a torch tensor / None."""
    if t is None:
        return None
    return t.data_ptr()


def stream():
    import torch
    return torch.cuda.current_stream().cuda_stream


def device_info():
    sm, maj, mnr = _i32(), _i32(), _i32()
    check(lib().pa_device_info(C.byref(sm), C.byref(maj), C.byref(mnr)), "pa_device_info")
    return sm.value, maj.value, mnr.value
