"""int8_quant / dnnl_matmul_int8 -- device versions of attention_cpu/int8_quant.hpp:5-18 and
attention_cpu/dnnl_matmul_int8.hpp:6-14 (same names, argument order and meaning).

Inputs and outputs are CUDA tensors (the reference takes std::vector / raw pointers).
Scale convention (int8_quant.cpp): multiply on quantise, divide on dequantise.
"""
import torch

from . import _cabi


def _chk(st, what):
    _cabi.check(st, what)


def quantize_to_int8(input, scale):
    """int8_quant.cpp:5-13."""
    x = input.contiguous()
    assert x.is_cuda and x.dtype == torch.float32
    q = torch.empty(x.shape, dtype=torch.int8, device=x.device)
    with torch.cuda.device(x.device):
        _chk(_cabi.lib().pa_quantize_i8(x.data_ptr(), x.numel(), float(scale), q.data_ptr(), _cabi.stream()),
             "pa_quantize_i8")
    return q


def batch_quantize(input, scales, dim):
    """int8_quant.cpp:15-28: input [B*dim], one scale per row of `dim`."""
    x = input.contiguous()
    s = scales.contiguous()
    assert x.is_cuda and x.dtype == torch.float32 and s.dtype == torch.float32
    assert x.numel() == s.numel() * dim
    q = torch.empty(x.shape, dtype=torch.int8, device=x.device)
    with torch.cuda.device(x.device):
        _chk(_cabi.lib().pa_batch_quantize_i8(x.data_ptr(), s.data_ptr(), s.numel(), dim, q.data_ptr(),
                                              _cabi.stream()), "pa_batch_quantize_i8")
    return q


def compute_absmax(input):
    """int8_quant.cpp:30-36 -> python float."""
    x = input.contiguous()
    out = torch.empty(1, dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        _chk(_cabi.lib().pa_absmax(x.data_ptr(), x.numel(), out.data_ptr(), _cabi.stream()), "pa_absmax")
    return float(out.item())


def dequantize_from_int8(input, scale):
    """int8_quant.cpp:38-44."""
    q = input.contiguous()
    assert q.is_cuda and q.dtype == torch.int8
    x = torch.empty(q.shape, dtype=torch.float32, device=q.device)
    with torch.cuda.device(q.device):
        _chk(_cabi.lib().pa_dequantize_i8(q.data_ptr(), q.numel(), float(scale), x.data_ptr(), _cabi.stream()),
             "pa_dequantize_i8")
    return x


def batch_dequantize(input, scales, dim):
    """int8_quant.cpp:46-57."""
    q = input.contiguous()
    s = scales.contiguous()
    assert q.numel() == s.numel() * dim
    x = torch.empty(q.shape, dtype=torch.float32, device=q.device)
    with torch.cuda.device(q.device):
        _chk(_cabi.lib().pa_batch_dequantize_i8(q.data_ptr(), s.data_ptr(), s.numel(), dim, x.data_ptr(),
                                                _cabi.stream()), "pa_batch_dequantize_i8")
    return x


def compute_minmax_scale(input):
    """int8_quant.cpp:59-64 -> python float."""
    x = input.contiguous()
    out = torch.empty(1, dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        _chk(_cabi.lib().pa_minmax_scale(x.data_ptr(), x.numel(), out.data_ptr(), _cabi.stream()),
             "pa_minmax_scale")
    return float(out.item())


def batch_minmax_scale(input, dim):
    """compute_minmax_scale per row of `dim` (the INT8 KV granularity)."""
    x = input.contiguous()
    rows = x.numel() // dim
    s = torch.empty(rows, dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        _chk(_cabi.lib().pa_batch_minmax_scale(x.data_ptr(), rows, dim, s.data_ptr(), _cabi.stream()),
             "pa_batch_minmax_scale")
    return s


def dnnl_matmul_int8(A, B, C, BATCH, M, N, K, scaleA, scaleB, scaleC=1.0, bias=None, activation="",
                     acc_out=None):
    """attention_cpu/dnnl_matmul_int8.cpp:7-76 on the tcgen05 kind::i8 tensor cores.

    A [BATCH,M,K], B [BATCH,K,N], C [BATCH,M,N] int8 CUDA tensors; bias [N] f32 or None;
    activation in {"", "relu", "gelu"}.  Returns True on success, False on failure (the
    reference swallows every exception into `false`, cpp:72-75).  `acc_out`, when given,
    receives the raw int32 accumulators.
    """
    if activation not in _cabi.ACT:
        return False
    try:
        with torch.cuda.device(A.device):
            lib = _cabi.lib()
            # split-K scratch is the caller's (stream-ordered torch allocation, released after the call is enqueued)
            need = lib.pa_gemm_i8_workspace_bytes(BATCH, M, N, K)
            ws = torch.empty(need, dtype=torch.uint8, device=A.device) if need else None
            st = lib.pa_gemm_i8(A.data_ptr(), B.data_ptr(), _cabi.ptr(C), _cabi.ptr(acc_out), BATCH, M,
                                N, K, float(scaleA), float(scaleB), float(scaleC), _cabi.ptr(bias),
                                _cabi.ACT[activation], _cabi.ptr(ws), need, _cabi.stream())
        return st == _cabi.PA_OK
    except Exception:
        return False
