"""AttentionCUDA / AttentionTileLauncher -- the reference's attention entry points over
libpa_b200.so.

Mirrors attention/attention_config.hpp:5-26 (`AttentionCUDA::forward`, 17 positional
arguments, same names and order) and attention/attention_tile_launcher.hpp:35-90.
Differences that are decisions: the KV storage dtype comes from the
`kv_cache` object, not from runtime heuristics on temperature/top_k (D-a4); softmax is
global over the context (D1); `out` is overwritten (D9); in-attention top-k/top-p is
rejected unless disabled (D6; the launcher default top_k=1 would zero every key but one);
`rerank_scores`, when given, receives the per-(b,h) log-sum-exp (D10).

q / out may be CUDA tensors (zero copy) or HOST buffers (numpy arrays or CPU tensors); host
buffers are staged through pinned memory and copied inside the call, which is what the
end-to-end bench measures.
"""
import os

import numpy as np
import torch

from . import _cabi


class _HostStage:
    """Pinned + device staging buffers for pageable host inputs, reused across calls.  One pair per
    (key, shape, dtype, device, STREAM): calls on different streams never share a buffer, and the event recorded
    after each H2D is waited for before the pinned buffer is overwritten by the next call (the DMA of call n may
    not have run yet when call n+1 stages its input)."""

    def __init__(self):
        self.bufs = {}

    def get(self, key, shape, dtype, device):
        stream = torch.cuda.current_stream(device).cuda_stream
        k = (key, tuple(shape), dtype, str(device), stream)
        if k not in self.bufs:
            self.bufs[k] = [torch.empty(shape, dtype=dtype).pin_memory(),
                            torch.empty(shape, dtype=dtype, device=device), None]
        return self.bufs[k]


_stage = _HostStage()


class HostPipe:
    """Page-locked host buffers <-> device, off the compute stream.

    One H2D stream and one D2H stream per device (the two copy engines) and small rings of device staging
    buffers, so that the upload of call n+1 and the download of call n run under the kernels of calls n / n+1
    instead of in line with them.  Ordering is by events only: the compute stream waits for an upload, a
    staging buffer is reused once the kernel that read it (or the download that drained it) has completed."""
    SLOTS = 4
    _pipes = {}

    @classmethod
    def get(cls, device):
        device = torch.device(device)
        key = (device.type, device.index if device.index is not None else torch.cuda.current_device())
        if key not in cls._pipes:
            cls._pipes[key] = cls(torch.device("cuda", key[1]))
        return cls._pipes[key]

    def __init__(self, device):
        self.device = device
        self.h2d = torch.cuda.Stream(device)
        self.d2h = torch.cuda.Stream(device)
        self.rings = {}
        self.pending = []   # completion events of downloads not yet waited for

    def _ring(self, key, shape, dtype):
        k = (key, tuple(shape), dtype)
        r = self.rings.get(k)
        if r is None:
            r = self.rings[k] = {"bufs": [torch.empty(shape, dtype=dtype, device=self.device) for _ in range(self.SLOTS)],
                                 "free": [None] * self.SLOTS, "idx": 0}
        i = r["idx"]
        r["idx"] = (i + 1) % self.SLOTS
        return r, i

    def upload(self, src, key):
        """Async H2D of a pinned tensor.  Returns (device tensor, release): call release() once the kernel that
        reads the tensor has been enqueued on the current stream."""
        r, i = self._ring(key, src.shape, src.dtype)
        buf = r["bufs"][i]
        main = torch.cuda.current_stream(self.device)
        if r["free"][i] is not None:
            self.h2d.wait_event(r["free"][i])
        with torch.cuda.stream(self.h2d):
            buf.copy_(src, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.h2d)
        main.wait_event(ev)

        def release():
            e = torch.cuda.Event()
            e.record(torch.cuda.current_stream(self.device))
            r["free"][i] = e
        return buf, release

    def out_slot(self, key, shape, dtype):
        """Device buffer a kernel on the current stream may write (its previous download has been ordered first)."""
        r, i = self._ring(key, shape, dtype)
        if r["free"][i] is not None:
            torch.cuda.current_stream(self.device).wait_event(r["free"][i])
        return r["bufs"][i], (r, i)

    def download(self, buf, slot, dst, sync=True):
        """D2H of `buf` (written on the current stream) into the pinned tensor `dst` on the download stream."""
        r, i = slot
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        self.d2h.wait_event(ev)
        with torch.cuda.stream(self.d2h):
            dst.copy_(buf, non_blocking=True)
            done = torch.cuda.Event()
            done.record(self.d2h)
        r["free"][i] = done
        if sync:
            done.synchronize()
        else:
            self.pending.append(done)
            if len(self.pending) > 4 * self.SLOTS:
                self.pending.pop(0).synchronize()
        return done

    def synchronize(self):
        """Host results of every sync=False call are valid after this."""
        for ev in self.pending:
            ev.synchronize()
        self.pending.clear()


def _is_host(x):
    return isinstance(x, np.ndarray) or (isinstance(x, torch.Tensor) and not x.is_cuda)


def _is_pinned(x):
    return isinstance(x, torch.Tensor) and not x.is_cuda and x.is_pinned()


def _to_device(x, key, device, dtype):
    """Host buffer -> device tensor (async H2D on the current stream).  Page-locked torch tensors are
    DMA'd directly; pageable memory (numpy arrays, ordinary CPU tensors) goes through a pinned staging
    buffer first."""
    src = torch.from_numpy(x) if isinstance(x, np.ndarray) else x
    slot = _stage.get(key, src.shape, dtype, device)
    pinned, dev = slot[0], slot[1]
    if _is_pinned(src) and src.dtype == dtype and src.is_contiguous():
        dev.copy_(src, non_blocking=True)
    else:
        if slot[2] is not None:
            slot[2].synchronize()  # the previous call's H2D out of this pinned buffer has completed
        pinned.copy_(src)
        dev.copy_(pinned, non_blocking=True)
        slot[2] = torch.cuda.Event()
        slot[2].record(torch.cuda.current_stream(device))
    return dev


class AttentionTileLauncher:
    """attention/attention_tile_launcher.hpp:35-90."""

    @staticmethod
    def launch(q, out, B, H, D, T, beam_ids=None, kv_cache=None, rotary_emb=None, is_prefill=True,
               use_fp16=False, use_overlap=False, temperature=1.0, top_k=0, top_p=1.0,
               rerank_scores=None, debug=False, ctx_lens=None, sync=True, logits=None, attention_weights=None):
        """sync=False (page-locked host q / out only): return once the work is enqueued; `out` is valid after
        AttentionCUDA.synchronize().  Copies then overlap the kernels of neighbouring calls (HostPipe).
        top_k > 0 / top_p < 1 (the reference's in-attention filter, softmax_lut.cpp:233-256) or the side outputs
        `logits` / `attention_weights` ([B, H, T] f32 CUDA tensors, cpu_attention_kernel.hpp:34-39) take the explicit
        three-stage kernels (pa_paged_attention_filtered); everything else the one-pass hot path."""
        if kv_cache is None:
            raise ValueError("AttentionTileLauncher.launch: kv_cache is required (paged path only)")
        if top_k not in (0, None) or top_p < 1.0 or logits is not None or attention_weights is not None:
            if _is_host(q) or _is_host(out) or rerank_scores is not None:
                raise NotImplementedError("filtered attention: device q / out only, no rerank_scores")
            return paged_attention_filtered(q, out, kv_cache, B, T, temperature, int(top_k or 0), float(top_p), beam_ids,
                                            ctx_lens, rotary_emb, logits, attention_weights)
        pt = kv_cache.page_table_
        dev = kv_cache.key_buffer_.device
        assert H == pt.num_heads_ and D == kv_cache.head_dim_
        if is_prefill and not _is_host(q) and T > 1 and q.numel() == B * H * T * D and kv_cache.dtype != "f32":
            # the reference's prefill form: q/out [B, H, T, D], causal self-attention over the T cached tokens
            if top_k not in (0, None) or top_p < 1.0 or rotary_emb is not None or rerank_scores is not None:
                raise NotImplementedError("prefill: top-k/top-p, RoPE table and rerank_scores are decode-only here")
            return paged_prefill(q, out, kv_cache, B, T, temperature, beam_ids)
        host_io = _is_host(q)
        host_out = _is_host(out)
        # page-locked host tensors take the copy-engine pipeline, anything else the in-line staging path
        pipe_q = host_io and _is_pinned(q) and q.dtype == torch.float32 and q.is_contiguous()
        pipe_out = host_out and _is_pinned(out) and out.dtype == torch.float32 and out.is_contiguous()
        pipe = HostPipe.get(dev) if (pipe_q or pipe_out) else None
        release_q = out_slot = None
        if pipe_q:
            with torch.cuda.device(dev):
                d_q, release_q = pipe.upload(q.view(B, H, D), "q")
        else:
            d_q = _to_device(q, "q", dev, torch.float32) if host_io else q
        if pipe_out:
            with torch.cuda.device(dev):
                d_out, out_slot = pipe.out_slot("out", (B, H, D), torch.float32)
        else:
            d_out = _stage.get("out", (B, H, D), torch.float32, dev)[1] if host_out else out
        assert d_q.dtype == torch.float32 and d_q.is_contiguous() and d_q.numel() == B * H * D
        assert d_out.dtype == torch.float32 and d_out.is_contiguous() and d_out.numel() == B * H * D
        d_beam = None
        if beam_ids is not None:
            d_beam = beam_ids if (isinstance(beam_ids, torch.Tensor) and beam_ids.is_cuda) else \
                torch.as_tensor(np.asarray(beam_ids), dtype=torch.int32).to(dev)
            assert d_beam.dtype == torch.int32
        d_ctx = None
        if ctx_lens is not None:
            d_ctx = ctx_lens if (isinstance(ctx_lens, torch.Tensor) and ctx_lens.is_cuda) else \
                torch.as_tensor(np.asarray(ctx_lens), dtype=torch.int32).to(dev)
        d_rope = None
        if rotary_emb is not None:
            d_rope = rotary_emb if (isinstance(rotary_emb, torch.Tensor) and rotary_emb.is_cuda) else \
                torch.as_tensor(np.asarray(rotary_emb), dtype=torch.float32).to(dev)
        d_lse = None
        host_lse = rerank_scores is not None and _is_host(rerank_scores)
        if rerank_scores is not None:
            d_lse = torch.empty((B, H), dtype=torch.float32, device=dev) if host_lse else rerank_scores
        table = pt.device_data()
        ws = kv_cache.workspace(B)
        lib = _cabi.lib()
        common = (table.data_ptr(), pt.num_beams_, pt.num_heads_, pt.num_tiles_, kv_cache.total_pages_,
                  _cabi.ptr(d_beam), _cabi.ptr(d_ctx), B, T, D, kv_cache.tile_size_, float(temperature),
                  _cabi.ptr(d_rope), _cabi.ptr(d_lse), ws.data_ptr(), ws.numel(), None)
        with torch.cuda.device(dev):
            common = common[:-1] + (_cabi.stream(),)
            if kv_cache.dtype == "f16":
                fn = lib.pa_paged_decode_f16_overlap if use_overlap else lib.pa_paged_decode_f16
                st = fn(d_q.data_ptr(), d_out.data_ptr(), kv_cache.key_buffer_.data_ptr(),
                        kv_cache.value_buffer_.data_ptr(), *common)
            elif kv_cache.dtype == "f32":  # KVTileCache<float>
                fn = lib.pa_paged_decode_f32_overlap if use_overlap else lib.pa_paged_decode_f32
                st = fn(d_q.data_ptr(), d_out.data_ptr(), kv_cache.key_buffer_.data_ptr(),
                        kv_cache.value_buffer_.data_ptr(), *common)
            else:
                fn = lib.pa_paged_decode_i8_overlap if use_overlap else lib.pa_paged_decode_i8
                st = fn(d_q.data_ptr(), d_out.data_ptr(), kv_cache.key_buffer_.data_ptr(),
                        kv_cache.value_buffer_.data_ptr(), kv_cache.k_scales_.data_ptr(),
                        kv_cache.v_scales_.data_ptr(), *common)
            _cabi.check(st, "pa_paged_decode")
            if debug:  # attention_tile_launcher.hpp:84-88
                torch.cuda.synchronize(dev)
            if release_q is not None:
                release_q()
            if host_out:
                if pipe_out:
                    pipe.download(d_out, out_slot, out.view(B, H, D), sync=sync)  # D2H straight into the caller's buffer
                else:
                    pinned = _stage.get("out", (B, H, D), torch.float32, dev)[0]
                    pinned.copy_(d_out, non_blocking=True)
                    torch.cuda.current_stream().synchronize()
                    dst = torch.from_numpy(out) if isinstance(out, np.ndarray) else out
                    dst.view(B, H, D).copy_(pinned)
            if host_lse:
                dst = torch.from_numpy(rerank_scores) if isinstance(rerank_scores, np.ndarray) else rerank_scores
                dst.view(B, H).copy_(d_lse.cpu())
        return out


class AttentionCUDA:
    """attention/attention_config.hpp:5-26 / attention_cuda.cu:41-95."""

    @staticmethod
    def forward(q, out, B, H, D, T, beam_ids=None, kv_cache=None, rotary_emb=None, is_prefill=True,
                use_fp16=False, use_overlap=False, temperature=1.0, top_k=0, top_p=1.0,
                rerank_scores=None, debug=False, ctx_lens=None, sync=True, logits=None, attention_weights=None):
        return AttentionTileLauncher.launch(q, out, B, H, D, T, beam_ids, kv_cache, rotary_emb,
                                            is_prefill, use_fp16, use_overlap, temperature, top_k,
                                            top_p, rerank_scores, debug, ctx_lens, sync, logits, attention_weights)

    @staticmethod
    def synchronize(device=None):
        """Completes every forward(..., sync=False) issued on `device` (default: the current one)."""
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        HostPipe.get(dev).synchronize()
        torch.cuda.current_stream(dev).synchronize()


def apply_rotary_embedding(q, k, rotary_emb, positions):
    """attention/attention_kernel_utils.cuh:20-35 for batches: q and/or k [rows, H, D] f32 CUDA tensors rotated
    in place with rotary_emb [T, D] (interleaved cos, sin) at token positions[row] (int32 CUDA tensor)."""
    ref = q if q is not None else k
    rows, H, D = ref.shape
    with torch.cuda.device(ref.device):
        _cabi.check(_cabi.lib().pa_apply_rope_f32(_cabi.ptr(q), _cabi.ptr(k), rotary_emb.data_ptr(), positions.data_ptr(),
                                                  rows, H, D, rotary_emb.shape[0], _cabi.stream()), "pa_apply_rope_f32")
    return q, k


def paged_prefill(q, out, kv_cache, B, Tq, temperature=1.0, beam_ids=None, ctx_start=None, token_major=False):
    """Causal multi-query attention of Tq new tokens per row over the paged cache
    (pa_paged_prefill_f16/_i8).  q/out: [B, H, Tq, D] f32 CUDA tensors (the reference's prefill layout,
    attention_config.hpp:8-9); ctx_start: [B] int32 device tensor of tokens cached before this chunk.
    token_major=True: q/out are [B, Tq, H, D] (the decoders' activation layout) and the strides are folded
    into the tcgen05 kernel (pa_paged_prefill_*_tokmajor); returns None instead of out where that kernel does
    not apply, and the caller permutes."""
    pt = kv_cache.page_table_
    H, D = pt.num_heads_, kv_cache.head_dim_
    assert q.is_contiguous() and out.is_contiguous() and q.numel() == B * H * Tq * D == out.numel()
    lib = _cabi.lib()
    if token_major:
        common = (pt.device_data().data_ptr(), pt.num_beams_, H, pt.num_tiles_, kv_cache.total_pages_, _cabi.ptr(beam_ids),
                  _cabi.ptr(ctx_start), B, Tq, D, kv_cache.tile_size_, float(temperature))
        with torch.cuda.device(kv_cache.key_buffer_.device):
            if kv_cache.dtype == "f16":
                st = lib.pa_paged_prefill_f16_tokmajor(q.data_ptr(), out.data_ptr(), kv_cache.key_buffer_.data_ptr(),
                                                       kv_cache.value_buffer_.data_ptr(), *common, _cabi.stream())
            elif kv_cache.dtype == "i8":
                st = lib.pa_paged_prefill_i8_tokmajor(q.data_ptr(), out.data_ptr(), kv_cache.key_buffer_.data_ptr(),
                                                      kv_cache.value_buffer_.data_ptr(), kv_cache.k_scales_.data_ptr(),
                                                      kv_cache.v_scales_.data_ptr(), *common, _cabi.stream())
            else:
                return None
        if st == _cabi.PA_ERR_UNSUPPORTED:
            return None
        _cabi.check(st, "pa_paged_prefill_tokmajor")
        return out
    # head_dim 128 takes a tensor-core flash-attention kernel (fp16 or int8 pages), which needs no scratch; head_dim 64
    # takes the tcgen05 kernel too but keeps the scratch for the row-per-query fallback (odd page sizes)
    fa = (D == 128 and kv_cache.tile_size_ % 16 == 0 and
          kv_cache.key_buffer_.data_ptr() % 128 == 0 and kv_cache.value_buffer_.data_ptr() % 128 == 0 and
          q.data_ptr() != out.data_ptr() and os.environ.get("PA_PREFILL_FA", "1") != "0")
    ws_ptr, ws_bytes = None, 0
    if not fa:
        need = lib.pa_prefill_workspace_bytes(B, Tq, H, D, pt.num_tiles_, kv_cache.tile_size_)
        ws = getattr(kv_cache, "_prefill_ws", None)
        if ws is None or ws.numel() < need:
            ws = kv_cache._prefill_ws = torch.empty(need, dtype=torch.uint8, device=kv_cache.key_buffer_.device)
        ws_ptr, ws_bytes = ws.data_ptr(), ws.numel()
    common = (pt.device_data().data_ptr(), pt.num_beams_, H, pt.num_tiles_, kv_cache.total_pages_, _cabi.ptr(beam_ids),
              _cabi.ptr(ctx_start), B, Tq, D, kv_cache.tile_size_, float(temperature), ws_ptr, ws_bytes)
    with torch.cuda.device(kv_cache.key_buffer_.device):
        if kv_cache.dtype == "f16":
            st = lib.pa_paged_prefill_f16(q.data_ptr(), out.data_ptr(), kv_cache.key_buffer_.data_ptr(),
                                          kv_cache.value_buffer_.data_ptr(), *common, _cabi.stream())
        else:
            st = lib.pa_paged_prefill_i8(q.data_ptr(), out.data_ptr(), kv_cache.key_buffer_.data_ptr(),
                                         kv_cache.value_buffer_.data_ptr(), kv_cache.k_scales_.data_ptr(),
                                         kv_cache.v_scales_.data_ptr(), *common, _cabi.stream())
    _cabi.check(st, "pa_paged_prefill")
    return out


def _dev_i32(x, dev):
    if x is None:
        return None
    if isinstance(x, torch.Tensor) and x.is_cuda:
        return x
    return torch.as_tensor(np.asarray(x), dtype=torch.int32).to(dev)


def paged_attention_filtered(q, out, kv_cache, B, T, temperature=1.0, top_k=0, top_p=1.0, beam_ids=None, ctx_lens=None,
                             rotary_emb=None, logits=None, attention_weights=None):
    """pa_paged_attention_filtered: the CPU kernel's K pass -> softmax -> top-k / top-p filter -> V pass
    (cpu_attention_kernel.cpp:61-126) with the optional side outputs logits / attention_weights [B, H, T]."""
    pt = kv_cache.page_table_
    dev = kv_cache.key_buffer_.device
    H, D = pt.num_heads_, kv_cache.head_dim_
    lib = _cabi.lib()
    kind = {"f16": 0, "i8": 1, "f32": 2}[kv_cache.dtype]
    need = lib.pa_attention_filtered_workspace_bytes(B, H, T)
    ws = torch.empty(need, dtype=torch.uint8, device=dev)
    d_beam, d_ctx = _dev_i32(beam_ids, dev), _dev_i32(ctx_lens, dev)
    d_rope = None if rotary_emb is None else (rotary_emb if (isinstance(rotary_emb, torch.Tensor) and rotary_emb.is_cuda)
                                              else torch.as_tensor(np.asarray(rotary_emb), dtype=torch.float32).to(dev))
    for t in (logits, attention_weights):
        assert t is None or (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous() and t.numel() == B * H * T)
    with torch.cuda.device(dev):
        st = lib.pa_paged_attention_filtered(
            q.data_ptr(), out.data_ptr(), kv_cache.key_buffer_.data_ptr(), kv_cache.value_buffer_.data_ptr(),
            _cabi.ptr(kv_cache.k_scales_), _cabi.ptr(kv_cache.v_scales_), kind, pt.device_data().data_ptr(), pt.num_beams_, H,
            pt.num_tiles_, kv_cache.total_pages_, _cabi.ptr(d_beam), _cabi.ptr(d_ctx), B, T, D, kv_cache.tile_size_,
            float(temperature), _cabi.ptr(d_rope), int(top_k), float(top_p), _cabi.ptr(logits), _cabi.ptr(attention_weights),
            ws.data_ptr(), ws.numel(), _cabi.stream())
    _cabi.check(st, "pa_paged_attention_filtered")
    return out


def paged_decode_group(q, out, kv_cache, B, T, beam_width, temperature=1.0, beam_ids=None, rotary_emb=None,
                       lse_out=None, ctx_lens=None):
    """Beam-aware decode (pa_paged_decode_f16_group): rows [g*W, (g+1)*W) are the W beams of group g;
    pages with equal ids within a group are read once.  q/out [B, H, D] f32 CUDA tensors."""
    pt = kv_cache.page_table_
    H, D = pt.num_heads_, kv_cache.head_dim_
    if kv_cache.dtype == "i8":  # per-row streaming kernel through the beam indirection; shared pages served from L2
        ws = kv_cache.workspace(B)
        with torch.cuda.device(kv_cache.key_buffer_.device):
            st = _cabi.lib().pa_paged_decode_i8_group(
                q.data_ptr(), out.data_ptr(), kv_cache.key_buffer_.data_ptr(), kv_cache.value_buffer_.data_ptr(),
                kv_cache.k_scales_.data_ptr(), kv_cache.v_scales_.data_ptr(), pt.device_data().data_ptr(), pt.num_beams_, H,
                pt.num_tiles_, kv_cache.total_pages_, _cabi.ptr(beam_ids), _cabi.ptr(ctx_lens), B, T, D, kv_cache.tile_size_,
                float(temperature), _cabi.ptr(rotary_emb), int(beam_width), _cabi.ptr(lse_out), ws.data_ptr(), ws.numel(),
                _cabi.stream())
        _cabi.check(st, "pa_paged_decode_i8_group")
        return out
    if kv_cache.dtype != "f16":
        raise NotImplementedError("paged_decode_group: fp16 or int8 KV pages")
    ws = kv_cache.workspace(B)
    with torch.cuda.device(kv_cache.key_buffer_.device):
        st = _cabi.lib().pa_paged_decode_f16_group(
            q.data_ptr(), out.data_ptr(), kv_cache.key_buffer_.data_ptr(), kv_cache.value_buffer_.data_ptr(),
            pt.device_data().data_ptr(), pt.num_beams_, H, pt.num_tiles_, kv_cache.total_pages_, _cabi.ptr(beam_ids),
            _cabi.ptr(ctx_lens), B, T, D, kv_cache.tile_size_, float(temperature), _cabi.ptr(rotary_emb), int(beam_width),
            _cabi.ptr(lse_out), ws.data_ptr(), ws.numel(), _cabi.stream())
    _cabi.check(st, "pa_paged_decode_f16_group")
    return out


def paged_decode_partial(q, kv_cache, B, T, temperature=1.0, beam_ids=None, ctx_lens=None, rotary_emb=None):
    """Un-normalised (m, l, O) of this rank's pages (multi-GPU split-KV): pa_paged_decode_{f16,i8}_partial."""
    pt = kv_cache.page_table_
    dev = kv_cache.key_buffer_.device
    H, D = pt.num_heads_, kv_cache.head_dim_
    pm = torch.empty((B, H), dtype=torch.float32, device=dev)
    pl = torch.empty((B, H), dtype=torch.float32, device=dev)
    po = torch.empty((B, H, D), dtype=torch.float32, device=dev)
    ws = kv_cache.workspace(B)
    lib = _cabi.lib()
    pools = (kv_cache.key_buffer_.data_ptr(), kv_cache.value_buffer_.data_ptr())
    if kv_cache.dtype == "i8":
        fn, name = lib.pa_paged_decode_i8_partial, "pa_paged_decode_i8_partial"
        pools += (kv_cache.k_scales_.data_ptr(), kv_cache.v_scales_.data_ptr())
    elif kv_cache.dtype == "f16":
        fn, name = lib.pa_paged_decode_f16_partial, "pa_paged_decode_f16_partial"
    else:
        raise NotImplementedError("paged_decode_partial: fp16 or int8 KV pages")
    with torch.cuda.device(dev):
        st = fn(q.data_ptr(), pm.data_ptr(), pl.data_ptr(), po.data_ptr(), *pools, pt.device_data().data_ptr(),
                pt.num_beams_, H, pt.num_tiles_, kv_cache.total_pages_, _cabi.ptr(beam_ids), _cabi.ptr(ctx_lens), B, T, D,
                kv_cache.tile_size_, float(temperature), _cabi.ptr(rotary_emb), ws.data_ptr(), ws.numel(), _cabi.stream())
    _cabi.check(st, name)
    return pm, pl, po


def lse_combine(part_m, part_l, part_o, out=None, lse_out=None):
    """part_m/l [n_parts, rows], part_o [n_parts, rows, D] -> out [rows, D]."""
    n_parts, rows, D = part_o.shape
    if out is None:
        out = torch.empty((rows, D), dtype=torch.float32, device=part_o.device)
    with torch.cuda.device(part_o.device):
        _cabi.check(_cabi.lib().pa_lse_combine(part_m.data_ptr(), part_l.data_ptr(), part_o.data_ptr(),
                                               n_parts, rows, D, out.data_ptr(), _cabi.ptr(lse_out),
                                               _cabi.stream()), "pa_lse_combine")
    return out
