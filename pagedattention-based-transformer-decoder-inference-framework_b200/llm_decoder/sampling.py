"""Device-side sampling over vocabulary logits: the reference's CPU filters
(attention_cpu/softmax_lut.cpp:203-256) and its GPU sketch (attention/top_k_top_p_filter.cuh:55-111)
as three kernels of libpa_b200.so.  Tensors are CUDA float32 [rows, vocab]."""
import torch

from . import _cabi


def build_exp_lut(resolution=1024, max_x=10.0, device=None):
    """build_exp_lut (softmax_lut.cpp:11-18): lut[i] = exp(-max_x + 2*max_x*i/(resolution-1)) in float32 with
    the C library's expf (what std::exp(float) calls), so the table is the reference's bit for bit."""
    import ctypes
    import ctypes.util

    import numpy as np
    libm = ctypes.CDLL(ctypes.util.find_library("m") or "libm.so.6")
    libm.expf.restype = ctypes.c_float
    libm.expf.argtypes = [ctypes.c_float]
    mx = np.float32(max_x)
    lut = np.empty(resolution, dtype=np.float32)
    for i in range(resolution):
        x = -mx + np.float32(2) * mx * np.float32(i) / np.float32(resolution - 1)
        lut[i] = libm.expf(float(np.float32(x)))
    t = torch.from_numpy(lut)
    return t.to(device) if device is not None else t


def softmax_lut(logits_i32, scale, lut, out=None):
    """fused_softmax_lut_inplace / softmax_batch_parallel (softmax_lut.cpp:60-100) over int32 logits
    [rows, n] -> f32 probabilities, bit-exact (pa_softmax_lut_i32)."""
    x = logits_i32.contiguous()
    assert x.dtype == torch.int32 and lut.dtype == torch.float32 and lut.is_cuda
    rows, n = x.shape
    p = torch.empty((rows, n), dtype=torch.float32, device=x.device) if out is None else out
    with torch.cuda.device(x.device):
        _cabi.check(_cabi.lib().pa_softmax_lut_i32(x.data_ptr(), rows, n, float(scale), lut.data_ptr(), lut.numel(),
                                                   p.data_ptr(), _cabi.stream()), "pa_softmax_lut_i32")
    return p


def softmax_temperature(logits, temperature=1.0, out=None):
    """softmax_lut_vec (softmax_lut.cpp:203-231): exp((x - max)/T) / (sum + 1e-6) per row."""
    x = logits.contiguous()
    rows, V = x.shape
    p = torch.empty_like(x) if out is None else out
    with torch.cuda.device(x.device):
        _cabi.check(_cabi.lib().pa_softmax_temperature(x.data_ptr(), rows, V, float(temperature), p.data_ptr(),
                                                       _cabi.stream()), "pa_softmax_temperature")
    return p


def apply_topk_topp_filter(probs, top_k, top_p, eos_token_id=-1, eos_thresh=0.0):
    """apply_topk_topp_filter (softmax_lut.cpp:233-256), in place on probs [rows, vocab]."""
    assert probs.is_contiguous() and probs.dtype == torch.float32
    rows, V = probs.shape
    with torch.cuda.device(probs.device):
        _cabi.check(_cabi.lib().pa_topk_topp_filter(probs.data_ptr(), rows, V, int(top_k), float(top_p),
                                                    int(eos_token_id), float(eos_thresh), _cabi.stream()),
                    "pa_topk_topp_filter")
    return probs


def sample_from_probs(probs, uniform, out=None):
    """First index whose inclusive prefix sum exceeds uniform[row] * sum(probs[row])."""
    rows, V = probs.shape
    ids = torch.empty(rows, dtype=torch.int32, device=probs.device) if out is None else out
    with torch.cuda.device(probs.device):
        _cabi.check(_cabi.lib().pa_sample_from_probs(probs.data_ptr(), rows, V, uniform.data_ptr(), ids.data_ptr(),
                                                     _cabi.stream()), "pa_sample_from_probs")
    return ids


def sample(logits, temperature=1.0, top_k=0, top_p=1.0, eos_token_id=-1, eos_thresh=0.0, generator=None):
    """logits [rows, vocab] -> sampled token ids [rows] (temperature softmax -> top-k/top-p/EOS filter ->
    inverse-CDF draw).  The uniform numbers come from a torch generator (plumbing, not math)."""
    probs = softmax_temperature(logits, temperature)
    apply_topk_topp_filter(probs, top_k, top_p, eos_token_id, eos_thresh)
    u = torch.rand(probs.shape[0], device=probs.device, generator=generator)
    return sample_from_probs(probs, u)
