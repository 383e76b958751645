"""CUDADecoder / INT8Decoder -- drop-in for the two classes the reference binds
(src/bindings.cpp:5-29): same constructor (six ints), `load_weights` /
`load_quantized_weights` / `quantize_weights`, and `generate(input_ids, max_len, temperature)`
returning prompt + generated ids.

The layer loop is the reference's (decoder/decoder_block.hpp:41-62): LN1 -> attention (q := LN1
output, B=1 per sequence, no QKV/O projections, no residuals) -> LN2 -> MLP(ReLU, 4x hidden,
decoder/mlp.hpp:23-41), greedy argmax sampling (cuda_decoder.cu:7-14 / int8_decoder.cpp:97-104).
Decisions where the reference is not well defined, stated here because
token-level parity with a program that never appends K/V and recomputes the whole sequence on
the host with aliased buffers is meaningless:
  * decode is incremental: each step appends the new token's K/V rows (K = V = the LN1 output
    split per head -- the only tensor the reference feeds to attention) to a per-layer paged
    KV cache (pa_kv_append_*), then attends over the cached context (pa_paged_decode_*);
  * logits = final hidden state . E^T (tied embedding); the reference reads the first
    vocab_size floats of the hidden buffer (cuda_decoder.cu:58);
  * INT8Decoder: int8 weights + int8 KV pages; activations are quantised per row with
    compute_minmax_scale/batch_quantize (int8_quant.cpp), the MLP runs on the tcgen05 kind::i8
    GEMM with the dequantising epilogue (pa_gemm_i8_dequant) -- the "int8_quant ->
    dnnl_matmul_int8 -> dequant" pipeline of attention_cpu/README.md:80-86 instead of
    MLP<int8_t>'s overflowing int8 accumulators (mlp.hpp:28).

Optional attention projections (weights/README.md:31-34): when a layer directory holds
attn_wq.bin / attn_wk.bin / attn_wv.bin / attn_wo.bin (each [hidden, hidden] row-major, column block h = head h's
[hidden, head_dim] matrix of the README), q = n.Wq, k = n.Wk, v = n.Wv feed the paged cache / attention and the
attention output is multiplied by Wo before LN2.  CUDADecoder: fp32 (pa_linear_f32); INT8Decoder: int8 weights on the
tcgen05 GEMM (pa_gemm_i8_dequant) with the LN1 output quantised per row.  Files absent = the reference block (q = k = v
= LN1 output, no output projection).

Every step runs on the device through libpa_b200.so (no host math, no per-step
synchronisation); the steady-state step is captured in a CUDA graph.
"""
import json
import os

import numpy as np
import torch

from . import _cabi
from .kv_tile_cache import KVTileCache

_LAYER_FILES = ("ln1.bin", "ln2.bin", "mlp_fc1.bin", "mlp_fc2.bin", "mlp_biases.bin")  # int8_decoder.cpp:66-70
_ATTN_FILES = ("attn_wq.bin", "attn_wk.bin", "attn_wv.bin", "attn_wo.bin")               # weights/README.md:31-34


def _read_bin(path, dtype, count, what):
    """load_vector_from_file (decoder/decoder_block.hpp:10-20): raw little-endian flat buffer."""
    if not os.path.isfile(path):
        raise RuntimeError(f"Failed to open weight file: {path}")
    a = np.fromfile(path, dtype=dtype)
    if a.size < count:
        raise RuntimeError(f"Failed to read from file: {path} ({what}: need {count} elements, file has {a.size})")
    return a[:count]


class _Layer:
    pass


class _Bufs:
    """Activation buffers for R rows (R = batch for a decode step, batch * prompt_len for the prefill)."""

    def __init__(self, R, hid, inter, dev):
        f32 = dict(dtype=torch.float32, device=dev)
        self.R = R
        self.x = torch.zeros((R, hid), **f32)
        self.n = torch.zeros((R, hid), **f32)
        self.a = torch.zeros((R, hid), **f32)
        self.h = torch.zeros((R, inter), **f32)
        self.xq = torch.zeros((R, hid), dtype=torch.int8, device=dev)
        self.hq = torch.zeros((R, inter), dtype=torch.int8, device=dev)
        self.xs = torch.ones(R, **f32)
        self.hs = torch.ones(R, **f32)
        self.ws = None  # GEMM / linear scratch for R rows (owned here, never by the library; _DecoderBase._scratch)
        self.qp = self.kp = self.vp = None  # q / k / v projections [R, hid] f32, allocated when a layer has them


class _DecoderBase:
    KV_DTYPE = "f16"
    ARGMAX_DIVIDE = 1  # cuda_decoder.cu:10-13 logits / temperature

    def __init__(self, num_layers, num_heads, head_dim, hidden_dim, vocab_size, max_seq_len, *, device=None,
                 tile_size=16, batch_size=1, attn_temperature=1.0, use_overlap=None, use_cuda_graph=True,
                 use_prefill=True):
        if min(num_layers, num_heads, head_dim, hidden_dim, vocab_size, max_seq_len) <= 0:
            raise ValueError("decoder dimensions must be positive")
        if hidden_dim != num_heads * head_dim:
            raise ValueError("hidden_dim must equal num_heads * head_dim: the block feeds the LN1 output "
                             "[hidden] to attention as q [H, D] (decoder_block.hpp:45-50)")
        if head_dim not in (64, 128):
            raise ValueError("head_dim must be 64 or 128 (kernels built for these)")
        self.num_layers_, self.num_heads_, self.head_dim_ = num_layers, num_heads, head_dim
        self.hidden_dim_, self.vocab_size_, self.max_seq_len_ = hidden_dim, vocab_size, max_seq_len
        self.inter_dim_ = hidden_dim * 4  # decoder_block.hpp:28
        self.tile_size_ = tile_size
        self.attn_temperature = float(attn_temperature)  # AttentionCUDA::forward default (attention_config.hpp:19)
        self.use_overlap = use_overlap
        self.use_cuda_graph = use_cuda_graph
        self.use_prefill = use_prefill
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self._lib = _cabi.lib()  # fails loudly when the CUDA library is missing
        self.eps = 1e-5  # layer_norm.hpp:10
        self.layers = [_Layer() for _ in range(num_layers)]
        self._alloc_default_weights()
        self._batch = 0
        self._graph = None
        self._setup_batch(batch_size)

    # ------------------------------------------------------------------ state
    def _setup_batch(self, B):
        if B == self._batch:
            return
        H, D, hid, dev = self.num_heads_, self.head_dim_, self.hidden_dim_, self.device
        nt = (self.max_seq_len_ + self.tile_size_ - 1) // self.tile_size_
        self._batch, self._num_tiles = B, nt
        table = np.arange(B * H * nt, dtype=np.int32).reshape(B, H, nt)
        self.kv_caches = []
        for _ in range(self.num_layers_):
            kvc = KVTileCache(self.KV_DTYPE, device=dev)
            kvc.init(B * H * nt, self.tile_size_, D)
            kvc.configure_table(B, H, nt)
            kvc.page_table_.load_host_table(table)
            self.kv_caches.append(kvc)
        f32 = dict(dtype=torch.float32, device=dev)
        self.ids = torch.zeros(B, dtype=torch.int32, device=dev)
        self.positions = torch.zeros(B, dtype=torch.int32, device=dev)
        self.ctx_lens = torch.ones(B, dtype=torch.int32, device=dev)
        self.bufs = _Bufs(B, hid, self.inter_dim_, dev)
        self.logits = torch.zeros((B, self.vocab_size_), **f32)
        self._best = torch.zeros(B, dtype=torch.int64, device=dev)  # argmax keys of pa_logits_argmax
        self._ws = self.kv_caches[0].workspace(B)
        self._graph = None
        self._step_warm = 0  # a new batch size gets one eager step before capture (lazy buffers, GEMM scratch)

    def reset(self):
        """Forget all cached context (positions back to 0; pages keep their assignment)."""
        self.positions.zero_()
        self.ctx_lens.fill_(1)

    # ------------------------------------------------------------------ device step
    def _chk(self, st, what):
        _cabi.check(st, what)

    def _scratch(self, bf, need):
        """Caller-owned scratch of the split-K / K-sliced matmuls for the rows of `bf` (pointer, bytes).  Lives as
        long as the activation buffers it belongs to, so a CUDA graph that captured the pointer stays valid; a
        request that does not fit re-allocates AND drops the captured graph."""
        if need == 0:
            return None, 0
        if bf.ws is None or bf.ws.numel() < need:
            bf.ws = torch.empty(need, dtype=torch.uint8, device=self.device)
            if bf is getattr(self, "bufs", None):
                self._graph = None
        return bf.ws.data_ptr(), bf.ws.numel()

    def _attention(self, layer_idx, q, out, R, ctx_lens, beam_ids, ws, prefill_shape=None):
        kvc = self.kv_caches[layer_idx]
        pt = kvc.page_table_
        lib = self._lib
        if prefill_shape is not None and self.head_dim_ in (64, 128):
            # prompt pass: the tcgen05 flash-attention prefill kernel ([B, H, n, D] layout)
            from .attention import paged_prefill
            B, n = prefill_shape
            H, D = self.num_heads_, self.head_dim_
            # q/out are [B, n, H, D] as the projection left them: the kernel takes the strides (no permute + copy)
            if q.data_ptr() != out.data_ptr() and \
                    paged_prefill(q, out, kvc, B, n, self.attn_temperature, token_major=True) is not None:
                return
            qb = q.view(B, n, H, D).permute(0, 2, 1, 3).contiguous()
            ob = torch.empty_like(qb)
            paged_prefill(qb, ob, kvc, B, n, self.attn_temperature)
            out.view(B, n, H, D).copy_(ob.permute(0, 2, 1, 3))
            return
        # the persistent kernel keeps a per-row chunk prefix in shared memory: very many rows (long prompts in
        # the row-per-query prefill) go through the split-KV grid kernel instead
        if self.use_overlap is None:
            # size rule (measured, paged_decode.cu): the persistent streaming kernel pays off above ~0.5 GB of K/V
            use_overlap = R * self.num_heads_ * ((self.max_seq_len_ + 15) // 16) > 65536
        else:
            use_overlap = self.use_overlap
        use_overlap = use_overlap and R <= 2048
        common = (pt.d_table_.data_ptr(), pt.num_beams_, pt.num_heads_, pt.num_tiles_, kvc.total_pages_,
                  _cabi.ptr(beam_ids), ctx_lens.data_ptr(), R, self.max_seq_len_, self.head_dim_, kvc.tile_size_,
                  self.attn_temperature, None, None, ws.data_ptr(), ws.numel(), _cabi.stream())
        if kvc.dtype == "f16":
            fn = lib.pa_paged_decode_f16_overlap if use_overlap else lib.pa_paged_decode_f16
            st = fn(q.data_ptr(), out.data_ptr(), kvc.key_buffer_.data_ptr(), kvc.value_buffer_.data_ptr(), *common)
        else:
            fn = lib.pa_paged_decode_i8_overlap if use_overlap else lib.pa_paged_decode_i8
            st = fn(q.data_ptr(), out.data_ptr(), kvc.key_buffer_.data_ptr(), kvc.value_buffer_.data_ptr(),
                    kvc.k_scales_.data_ptr(), kvc.v_scales_.data_ptr(), *common)
        self._chk(st, "pa_paged_decode")

    def _layers(self, bf, ids, positions, ctx_lens, beam_ids, ws, prefill_shape=None):
        """The decoder stack (decoder_block.hpp:41-62) for bf.R rows: row r is token ids[r] at position
        positions[r] of table row beam_ids[r] (None: r) attending ctx_lens[r] cached tokens (its own K/V
        included).  Leaves the final hidden states in bf.x."""
        lib, R, hid, s = self._lib, bf.R, self.hidden_dim_, _cabi.stream()
        self._embed(bf, ids)
        for li, L in enumerate(self.layers):
            self._chk(lib.pa_layer_norm_f32(bf.x.data_ptr(), L.ln1_g.data_ptr(), L.ln1_b.data_ptr(), R, hid,
                                            self.eps, bf.n.data_ptr(), s), "pa_layer_norm_f32")
            H, D = self.num_heads_, self.head_dim_
            if getattr(L, "wq", None) is None:
                nview = bf.n.view(R, H, D)
                self.kv_caches[li].append(nview, nview, positions, beam_ids)  # K = V = LN1 output (see module doc)
                self._attention(li, bf.n, bf.a, R, ctx_lens, beam_ids, ws, prefill_shape)
            else:
                if bf.qp is None:
                    bf.qp, bf.kp, bf.vp = (torch.empty((R, hid), dtype=torch.float32, device=self.device) for _ in range(3))
                self._qkv_proj(bf, L)                                   # bf.n -> bf.qp, bf.kp, bf.vp
                self.kv_caches[li].append(bf.kp.view(R, H, D), bf.vp.view(R, H, D), positions, beam_ids)
                self._attention(li, bf.qp, bf.kp, R, ctx_lens, beam_ids, ws, prefill_shape)   # kp reused as the output
                self._o_proj(bf, L, bf.kp, bf.a)                        # attention output . Wo -> bf.a
            self._ln2_mlp(bf, L)

    def _ln2_mlp(self, bf, L):
        """LN2 -> MLP (decoder_block.hpp:55-60): bf.a -> bf.x."""
        self._chk(self._lib.pa_layer_norm_f32(bf.a.data_ptr(), L.ln2_g.data_ptr(), L.ln2_b.data_ptr(), bf.R,
                                              self.hidden_dim_, self.eps, bf.n.data_ptr(), _cabi.stream()),
                  "pa_layer_norm_f32")
        self._mlp(bf, L)

    def _head(self, x_rows):
        """logits (tied embedding) of the B rows in x_rows, then the next ids -> self.ids: the reference's
        greedy argmax by default; temperature / top-k / top-p / EOS sampling (sampling.py) when requested."""
        sp = getattr(self, "_sampling", None)
        if sp is None:  # logits and the greedy sampler in one pass over the embedding table
            E, eb, qs = self._embedding_table()
            self._chk(self._lib.pa_logits_argmax(x_rows.data_ptr(), E.data_ptr(), eb, qs, self._batch,
                                                 self.hidden_dim_, self.vocab_size_, self._temperature,
                                                 self.ARGMAX_DIVIDE, self.logits.data_ptr(), self._best.data_ptr(),
                                                 self.ids.data_ptr(), _cabi.stream()), "pa_logits_argmax")
            return
        self._logits(x_rows)
        from . import sampling
        # CUDADecoder scales logits by 1/T (cuda_decoder.cu:10-13), INT8Decoder by T (int8_decoder.cpp:100)
        t_eff = self._temperature if self.ARGMAX_DIVIDE else 1.0 / self._temperature
        probs = sampling.softmax_temperature(self.logits, t_eff)
        sampling.apply_topk_topp_filter(probs, sp["top_k"], sp["top_p"], sp["eos_token_id"], sp["eos_thresh"])
        u = torch.rand(self._batch, device=self.device, generator=sp["generator"])
        sampling.sample_from_probs(probs, u, out=self.ids)

    def _step(self):
        """One decode step for the B tokens in self.ids at self.positions; writes the greedy next
        ids back into self.ids and advances positions.  Device work only (CUDA-graph capturable)."""
        self._layers(self.bufs, self.ids, self.positions, self.ctx_lens, None, self._ws)
        self._head(self.bufs.x)
        self._chk(self._lib.pa_advance_positions(self.positions.data_ptr(), self.ctx_lens.data_ptr(), self._batch,
                                                 _cabi.stream()), "pa_advance_positions")

    def _prefill(self, prompt):
        """All prompt tokens of all sequences in ONE pass through the stack: row
        (b, t) appends its K/V at position t of table row b and attends causally (ctx = t + 1) through
        the beam indirection of the decode kernels; LayerNorm and the MLP GEMMs run on B*n rows.
        Leaves the first sampled token in self.ids and positions = n."""
        B, n = prompt.shape
        dev = self.device
        R = B * n
        bf = _Bufs(R, self.hidden_dim_, self.inter_dim_, dev)
        t = torch.arange(n, dtype=torch.int32, device=dev)
        positions = t.repeat(B)
        ctx = positions + 1
        beam = torch.arange(B, dtype=torch.int32, device=dev).repeat_interleave(n)
        ws = torch.empty(self._lib.pa_decode_workspace_bytes(R, self.num_heads_, self.head_dim_, self._num_tiles,
                                                             self.tile_size_), dtype=torch.uint8, device=dev)
        self._layers(bf, prompt.reshape(-1).contiguous(), positions, ctx, beam, ws, prefill_shape=(B, n))
        last = bf.x.view(B, n, self.hidden_dim_)[:, n - 1, :].contiguous()
        self._head(last)
        self.positions.fill_(n)
        self.ctx_lens.fill_(n + 1)

    # ------------------------------------------------------------------ generate
    @staticmethod
    def normalize_generate_args(input_ids, max_len=None, temperature=1.0, *extra, max_gen_len=None, max_tokens=None):
        """The call forms of `generate` found in the reference, reduced to (input_ids, out_list, max_len, temperature):
          * pybind form        generate(input_ids, max_len, temperature)            src/bindings.cpp:8-15
          * C++ out-param form generate(input_ids, output_ids, max_tokens[, temp])  api/router.py:23, web/app.py:23,
                                                                                   cli/generate_batch.py:23
          * keyword forms      generate(input_ids, output_ids, max_gen_len=64[, temperature=1.0])
                                                                                   cli/chat_cli.py:24, cli/stream_cli.py:16
        Pure host logic (no device work), so the call-site compatibility tests run without a GPU."""
        out_list = None
        if isinstance(max_len, list):  # 4-arg out-param form: (input_ids, output_ids, max_tokens, temperature)
            out_list = max_len
            if isinstance(temperature, (int, float)) and (extra or max_gen_len is None and max_tokens is None):
                # positional max_tokens landed in `temperature`, the real temperature (if any) in extra[0]
                max_len, temperature = temperature, (extra[0] if extra else 1.0)
            else:
                max_len = None  # max_gen_len= / max_tokens= keyword follows; `temperature` is the real one
        elif extra:
            raise TypeError("generate() takes at most 4 positional arguments")
        if max_gen_len is not None:
            max_len = max_gen_len
        if max_tokens is not None:
            max_len = max_tokens
        if max_len is None:
            raise TypeError("generate() missing max_len")
        return list(input_ids), out_list, int(max_len), float(temperature)

    def generate(self, input_ids, max_len=None, temperature=1.0, *extra, max_gen_len=None, max_tokens=None,
                 **sampling_kw):
        """generate(input_ids, max_len, temperature) -> prompt + generated ids (bindings.cpp:8-15).
        Also accepted (callers in api/, cli/, web/; see normalize_generate_args): generate(input_ids,
        output_ids_list, max_tokens, temperature) -- the list is overwritten in place with prompt + generated
        ids (cuda_decoder.cu:49,59: `output_ids = input_ids` then push_back) and returned -- and max_gen_len=."""
        ids, out_list, max_len, temperature = self.normalize_generate_args(
            input_ids, max_len, temperature, *extra, max_gen_len=max_gen_len, max_tokens=max_tokens)
        seqs = self.generate_batch([ids], max_len, temperature, **sampling_kw)
        if out_list is not None:
            out_list[:] = seqs[0]
            return out_list
        return seqs[0]

    def generate_batch(self, prompts, max_len, temperature=1.0, top_k=0, top_p=1.0, eos_token_id=-1,
                       eos_thresh=0.0, seed=None):
        """Decode several sequences at once (equal prompt lengths): rows are independent.  Greedy argmax (the
        reference's sampler) unless top_k > 0 or top_p < 1, which switch to temperature / top-k / top-p / EOS
        sampling on the device with a torch generator seeded by `seed`."""
        B = len(prompts)
        n_prompt = len(prompts[0])
        if n_prompt == 0 or any(len(p) != n_prompt for p in prompts):
            raise ValueError("prompts must be non-empty and of equal length")
        if n_prompt + max_len - 1 > self.max_seq_len_:
            raise ValueError("prompt + max_len exceeds max_seq_len")
        if max_len <= 0:
            return [list(p) for p in prompts]
        self._setup_batch(B)
        self.reset()
        self._temperature = float(temperature)
        dev = self.device
        self._sampling = None
        if top_k > 0 or top_p < 1.0:
            gen = torch.Generator(device=dev)
            gen.manual_seed(0 if seed is None else int(seed))
            self._sampling = dict(top_k=int(top_k), top_p=float(top_p), eos_token_id=int(eos_token_id),
                                  eos_thresh=float(eos_thresh), generator=gen)
        with torch.cuda.device(dev):
            prompt = torch.tensor(prompts, dtype=torch.int32).pin_memory().to(dev, non_blocking=True)  # [B, n]
            gen = torch.empty((max_len, B), dtype=torch.int32, device=dev)
            if self.use_prefill:
                self._prefill(prompt)
            else:
                for t in range(n_prompt):  # prompt tokens one decode step each
                    self.ids.copy_(prompt[:, t])
                    self._step_or_replay()
            for i in range(max_len):  # self.ids holds the token sampled by the previous step
                gen[i].copy_(self.ids)
                if i + 1 < max_len:
                    self._step_or_replay()
            gen_host = gen.t().cpu().tolist()  # the only synchronisation of the call
        return [list(p) + g for p, g in zip(prompts, gen_host)]

    def forward_tokens(self, token_ids, temperature=1.0):
        """One eager decode step for `token_ids` ([B] ints) at the current positions: appends their K/V,
        returns the logits [B, vocab] (device tensor, valid until the next step).  The sampled next ids are
        in `self.ids`.  Call `reset()` first to start a new sequence."""
        self._setup_batch(len(token_ids))
        self._temperature = float(temperature)
        with torch.cuda.device(self.device):
            self.ids.copy_(torch.tensor(list(token_ids), dtype=torch.int32))
            self._step()
        return self.logits

    def _step_or_replay(self):
        if not self.use_cuda_graph or getattr(self, "_sampling", None) is not None:
            self._step()  # the sampling path draws from a torch generator: not captured
        elif self._graph is not None and self._graph_temp == self._temperature:
            self._graph.replay()
        else:
            self._step_warm = getattr(self, "_step_warm", 0)
            if self._step_warm < 1:
                self._step()  # one eager step first: lazy one-time init inside the library
                self._step_warm += 1
            else:
                # capture a graph of the step; capture itself does not run it, so replay once
                torch.cuda.current_stream().synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self._step()
                self._graph, self._graph_temp = g, self._temperature
                g.replay()

    # ------------------------------------------------------------------ helpers
    def _dev(self, a, dtype=torch.float32):
        return torch.from_numpy(np.ascontiguousarray(a)).to(self.device).to(dtype)


class CUDADecoder(_DecoderBase):
    """py::class_<CUDADecoder<float>>(m, "CUDADecoder") -- src/bindings.cpp:5-15;
    decoder/cuda_decoder.hpp:6-19, cuda_decoder.cu:23-61.  fp32 weights, fp16 KV pages."""
    KV_DTYPE = "f16"
    ARGMAX_DIVIDE = 1

    def _alloc_default_weights(self):
        hid, inter, dev = self.hidden_dim_, self.inter_dim_, self.device
        z = lambda *s: torch.zeros(s, dtype=torch.float32, device=dev)  # noqa: E731  std::vector<T>(n) zero-init
        self.embedding = z(self.vocab_size_, hid)
        for L in self.layers:
            L.ln1_g, L.ln1_b = torch.ones(hid, device=dev), z(hid)  # layer_norm.hpp:11 gamma=1, beta=0
            L.ln2_g, L.ln2_b = torch.ones(hid, device=dev), z(hid)
            L.fc1_w, L.fc1_b, L.fc2_w, L.fc2_b = z(hid, inter), z(inter), z(inter, hid), z(hid)
            L.wq = L.wk = L.wv = L.wo = None

    def load_weights(self, path):
        """cuda_decoder.cu:35-45: <path>/embedding.bin, <path>/layer_<i>/{ln1.bin, ln2.bin} (gamma then
        beta, layer_norm.hpp:13-18) and the MLP.  The reference's MLP::load_weights opens the layer
        DIRECTORY as a file (decoder_block.hpp:38, mlp.hpp:15), which cannot work; accepted here:
        <layer>/mlp.bin = fc1_w, fc1_b, fc2_w, fc2_b concatenated (the order mlp.hpp:17-20 reads), or the
        split files mlp_fc1.bin / mlp_fc2.bin / mlp_biases.bin of weights/README.md and quantize_weights."""
        hid, inter, V = self.hidden_dim_, self.inter_dim_, self.vocab_size_
        self.embedding = self._dev(_read_bin(os.path.join(path, "embedding.bin"), np.float32, V * hid,
                                             "embedding").reshape(V, hid))
        for i, L in enumerate(self.layers):
            lp = os.path.join(path, f"layer_{i}")
            ln1 = _read_bin(os.path.join(lp, "ln1.bin"), np.float32, 2 * hid, "ln1 gamma+beta")
            ln2 = _read_bin(os.path.join(lp, "ln2.bin"), np.float32, 2 * hid, "ln2 gamma+beta")
            L.ln1_g, L.ln1_b = self._dev(ln1[:hid]), self._dev(ln1[hid:])
            L.ln2_g, L.ln2_b = self._dev(ln2[:hid]), self._dev(ln2[hid:])
            packed = os.path.join(lp, "mlp.bin")
            if os.path.isfile(packed):
                m = _read_bin(packed, np.float32, hid * inter * 2 + inter + hid, "mlp")
                o = 0
                fc1 = m[o:o + hid * inter]; o += hid * inter
                b1 = m[o:o + inter]; o += inter
                fc2 = m[o:o + inter * hid]; o += inter * hid
                b2 = m[o:o + hid]
            elif os.path.isfile(os.path.join(lp, "mlp_fc1.bin")):
                fc1 = _read_bin(os.path.join(lp, "mlp_fc1.bin"), np.float32, hid * inter, "mlp_fc1")
                fc2 = _read_bin(os.path.join(lp, "mlp_fc2.bin"), np.float32, inter * hid, "mlp_fc2")
                bb = _read_bin(os.path.join(lp, "mlp_biases.bin"), np.float32, inter + hid, "mlp_biases")
                b1, b2 = bb[:inter], bb[inter:]
            else:
                raise RuntimeError("Cannot open MLP weights file")  # mlp.hpp:16
            L.fc1_w, L.fc1_b = self._dev(fc1.reshape(hid, inter)), self._dev(b1)
            L.fc2_w, L.fc2_b = self._dev(fc2.reshape(inter, hid)), self._dev(b2)
            L.wq = L.wk = L.wv = L.wo = None
            if os.path.isfile(os.path.join(lp, "attn_wq.bin")):  # weights/README.md:31-34 (optional)
                for nm in ("wq", "wk", "wv", "wo"):
                    setattr(L, nm, self._dev(_read_bin(os.path.join(lp, f"attn_{nm}.bin"), np.float32, hid * hid,
                                                       f"attn_{nm}").reshape(hid, hid)))
        self._graph = None
        self._emb_T = None
        self._pk = {}   # packed copies belong to the replaced tensors

    def _embedding_table(self):
        return self.embedding, 4, 1.0

    def _packed(self, W, K, N):
        """The tensor-core kernel's own weight order (pa_linear_pack_f32: every streamed block one contiguous 16 KB run,
        ~10 % faster than the [K, N] rows), made once per weight tensor and re-made when the tensor was written in place
        (torch's version counter).  None: PA_LINEAR_PACKED=0, K % 4 != 0, or a copy would have to be made during graph
        capture -- the caller then uses the [K, N] layout."""
        if K % 4 or os.environ.get("PA_LINEAR_PACKED", "1") == "0":
            return None
        cache = self.__dict__.setdefault("_pk", {})
        ent = cache.get(W.data_ptr())
        if ent is not None and ent[0] is W and ent[1] == W._version:   # (the entry keeps W alive: its address is not reused)
            return ent[2]
        if torch.cuda.is_current_stream_capturing():
            return None
        Wp = torch.empty(self._lib.pa_linear_pack_bytes(K, N) // 4, dtype=torch.float32, device=self.device)
        self._chk(self._lib.pa_linear_pack_f32(W.data_ptr(), Wp.data_ptr(), K, N, _cabi.stream()), "pa_linear_pack_f32")
        cache[W.data_ptr()] = (W, W._version, Wp)
        return Wp

    def _linear(self, x, W, bias, R, K, N, act, out, wp, wb):
        lib = self._lib
        Wp = self._packed(W, K, N) if R >= 16 else None   # few rows: the strip kernel on the [K, N] rows
        if Wp is not None:
            st = lib.pa_linear_f32_packed(x.data_ptr(), Wp.data_ptr(), _cabi.ptr(bias), R, K, N, _cabi.ACT[act],
                                          out.data_ptr(), wp, wb, _cabi.stream())
            if st != _cabi.PA_ERR_UNSUPPORTED:
                return self._chk(st, "pa_linear_f32_packed")
        self._chk(lib.pa_linear_f32(x.data_ptr(), W.data_ptr(), _cabi.ptr(bias), R, K, N, _cabi.ACT[act], out.data_ptr(),
                                    wp, wb, _cabi.stream()), "pa_linear_f32")

    def _lin(self, bf, x, W, out):
        R, hid = bf.R, self.hidden_dim_
        wp, wb = self._scratch(bf, self._lib.pa_linear_workspace_bytes(R, hid, hid))
        self._linear(x, W, None, R, hid, hid, "", out, wp, wb)

    def _qkv_proj(self, bf, L):
        self._lin(bf, bf.n, L.wq, bf.qp)
        self._lin(bf, bf.n, L.wk, bf.kp)
        self._lin(bf, bf.n, L.wv, bf.vp)

    def _o_proj(self, bf, L, x, out):
        self._lin(bf, x, L.wo, out)

    def _head(self, x_rows):
        """More than 8 rows: logits = x . E^T through the fp32 linear kernel on a transposed copy of the table
        (weights streamed once per 16 rows) followed by the argmax kernel; up to 8 rows: the GEMV with the
        sampler folded in."""
        B = self._batch
        if B <= 8 or getattr(self, "_sampling", None) is not None:
            return super()._head(x_rows)
        if getattr(self, "_emb_T", None) is None or self._emb_T.shape != (self.hidden_dim_, self.vocab_size_):
            self._emb_T = self.embedding.t().contiguous()
        lib, s = self._lib, _cabi.stream()
        wp, wb = self._scratch(self.bufs, lib.pa_linear_workspace_bytes(B, self.hidden_dim_, self.vocab_size_))
        self._linear(x_rows, self._emb_T, None, B, self.hidden_dim_, self.vocab_size_, "", self.logits, wp, wb)
        self._chk(lib.pa_argmax_f32(self.logits.data_ptr(), B, self.vocab_size_, self._temperature,
                                    self.ARGMAX_DIVIDE, self.ids.data_ptr(), s), "pa_argmax_f32")

    def _embed(self, bf, ids):
        self._chk(self._lib.pa_embedding_f32(self.embedding.data_ptr(), ids.data_ptr(), bf.R, self.hidden_dim_,
                                             self.vocab_size_, bf.x.data_ptr(), _cabi.stream()), "pa_embedding_f32")

    def _mlp(self, bf, L):
        lib, R, hid, inter, s = self._lib, bf.R, self.hidden_dim_, self.inter_dim_, _cabi.stream()
        wp, wb = self._scratch(bf, max(lib.pa_linear_workspace_bytes(R, hid, inter),
                                       lib.pa_linear_workspace_bytes(R, inter, hid)))
        self._linear(bf.n, L.fc1_w, L.fc1_b, R, hid, inter, "relu", bf.h, wp, wb)
        self._linear(bf.h, L.fc2_w, L.fc2_b, R, inter, hid, "", bf.x, wp, wb)

    def _logits(self, x_rows):
        self._chk(self._lib.pa_logits_f32(x_rows.data_ptr(), self.embedding.data_ptr(), self._batch,
                                          self.hidden_dim_, self.vocab_size_, self.logits.data_ptr(), _cabi.stream()),
                  "pa_logits_f32")


def quantize_file_reference(w):
    """INT8Quantizer / quantize_weights arithmetic (int8_decoder.cpp:52-56): scale = max_element
    (SIGNED max, not absmax), q = static_cast<int8_t>(w / scale * 127): truncation toward zero,
    no clamp (out-of-range values wrap as an x86 float->int32->int8 conversion does)."""
    w = np.ascontiguousarray(w, dtype=np.float32)
    scale = np.float32(w.max())
    with np.errstate(all="ignore"):
        v = (w / scale) * np.float32(127)
        q = np.trunc(v)
        q = np.where(np.isfinite(q), q, 0).astype(np.int64).astype(np.int32).astype(np.int8)
    return q, float(scale)


class INT8Decoder(_DecoderBase):
    """py::class_<INT8Decoder>(m, "INT8Decoder") -- src/bindings.cpp:18-29;
    decoder/int8_decoder.hpp:6-19, int8_decoder.cpp:34-119.  int8 weights, int8 KV pages."""
    KV_DTYPE = "i8"
    ARGMAX_DIVIDE = 0  # int8_decoder.cpp:100 logits * temperature
    SCALES_FILE = "quant_scales.json"

    def _alloc_default_weights(self):
        hid, inter, dev = self.hidden_dim_, self.inter_dim_, self.device
        zi = lambda *s: torch.zeros(s, dtype=torch.int8, device=dev)  # noqa: E731
        z = lambda *s: torch.zeros(s, dtype=torch.float32, device=dev)  # noqa: E731
        self.embedding = zi(self.vocab_size_, hid)
        self.emb_qscale = 127.0
        for L in self.layers:
            L.ln1_g, L.ln1_b = torch.ones(hid, device=dev), z(hid)
            L.ln2_g, L.ln2_b = torch.ones(hid, device=dev), z(hid)
            L.fc1_w, L.fc2_w = zi(hid, inter), zi(inter, hid)
            L.fc1_deq = L.fc2_deq = 1.0 / 127.0
            L.fc1_b, L.fc2_b = z(inter), z(hid)
            L.wq = L.wk = L.wv = L.wo = None

    def quantize_weights(self, path_fp32, path_int8):
        """int8_decoder.cpp:43-89: embedding.bin and, per layer, ln1/ln2/mlp_fc1/mlp_fc2/mlp_biases .bin are
        quantised file by file with the reference's arithmetic (quantize_file_reference) and written as raw
        int8.  The reference drops the per-file scale, which makes the int8 files undecodable; it is kept
        here in a sidecar <path_int8>/quant_scales.json (an addition, ignored by the reference)."""
        os.makedirs(path_int8, exist_ok=True)
        scales = {}

        def one(rel):
            src = os.path.join(path_fp32, rel)
            if not os.path.isfile(src):
                raise RuntimeError(f"Failed to open weight file: {src}")
            q, sc = quantize_file_reference(np.fromfile(src, dtype=np.float32))
            q.tofile(os.path.join(path_int8, rel))
            scales[rel] = sc

        one("embedding.bin")
        for i in range(self.num_layers_):
            os.makedirs(os.path.join(path_int8, f"layer_{i}"), exist_ok=True)
            for fname in _LAYER_FILES:
                one(f"layer_{i}/{fname}")
            for fname in _ATTN_FILES:  # optional projections (weights/README.md:31-34): quantised like every other file
                if os.path.isfile(os.path.join(path_fp32, f"layer_{i}", fname)):
                    one(f"layer_{i}/{fname}")
        with open(os.path.join(path_int8, self.SCALES_FILE), "w") as f:
            json.dump(scales, f, indent=1)
        print(f"[INT8Decoder] Quantized weights saved to {path_int8}")  # int8_decoder.cpp:88

    def load_quantized_weights(self, path_int8):
        """int8_decoder.cpp:91-95.  Dequantisation multiplier of a file = scale/127 (w ~ q*scale/127);
        scale comes from quant_scales.json, 1.0 when the sidecar is missing."""
        hid, inter, V = self.hidden_dim_, self.inter_dim_, self.vocab_size_
        scales = {}
        sp = os.path.join(path_int8, self.SCALES_FILE)
        if os.path.isfile(sp):
            with open(sp) as f:
                scales = json.load(f)
        deq = lambda rel: float(scales.get(rel, 1.0)) / 127.0  # noqa: E731
        emb = _read_bin(os.path.join(path_int8, "embedding.bin"), np.int8, V * hid, "embedding")
        self.embedding = self._dev(emb.reshape(V, hid), torch.int8)
        self.emb_qscale = 1.0 / deq("embedding.bin")
        for i, L in enumerate(self.layers):
            rel = lambda f: f"layer_{i}/{f}"  # noqa: E731
            rd = lambda f, n: _read_bin(os.path.join(path_int8, rel(f)), np.int8, n, f).astype(np.float32)  # noqa: E731
            ln1 = rd("ln1.bin", 2 * hid) * np.float32(deq(rel("ln1.bin")))
            ln2 = rd("ln2.bin", 2 * hid) * np.float32(deq(rel("ln2.bin")))
            L.ln1_g, L.ln1_b = self._dev(ln1[:hid]), self._dev(ln1[hid:])
            L.ln2_g, L.ln2_b = self._dev(ln2[:hid]), self._dev(ln2[hid:])
            L.fc1_w = self._dev(_read_bin(os.path.join(path_int8, rel("mlp_fc1.bin")), np.int8, hid * inter,
                                          "mlp_fc1").reshape(hid, inter), torch.int8)
            L.fc2_w = self._dev(_read_bin(os.path.join(path_int8, rel("mlp_fc2.bin")), np.int8, inter * hid,
                                          "mlp_fc2").reshape(inter, hid), torch.int8)
            L.fc1_deq, L.fc2_deq = deq(rel("mlp_fc1.bin")), deq(rel("mlp_fc2.bin"))
            bb = rd("mlp_biases.bin", inter + hid) * np.float32(deq(rel("mlp_biases.bin")))
            L.fc1_b, L.fc2_b = self._dev(bb[:inter]), self._dev(bb[inter:])
            L.wq = L.wk = L.wv = L.wo = None
            if os.path.isfile(os.path.join(path_int8, rel("attn_wq.bin"))):
                for nm in ("wq", "wk", "wv", "wo"):
                    f = f"attn_{nm}.bin"
                    setattr(L, nm, self._dev(_read_bin(os.path.join(path_int8, rel(f)), np.int8, hid * hid, f).reshape(hid, hid),
                                             torch.int8))
                    setattr(L, nm + "_deq", deq(rel(f)))
        self._graph = None
        self._emb_T = None

    def _gemm_deq(self, bf, xq, xs, W, w_deq, out):
        lib, R, hid = self._lib, bf.R, self.hidden_dim_
        wp, wb = self._scratch(bf, lib.pa_gemm_i8_workspace_bytes(1, R, hid, hid))
        self._chk(lib.pa_gemm_i8_dequant(xq.data_ptr(), W.data_ptr(), out.data_ptr(), 1, R, hid, hid, xs.data_ptr(),
                                         float(w_deq), None, _cabi.ACT[""], wp, wb, _cabi.stream()), "pa_gemm_i8_dequant")

    def _qkv_proj(self, bf, L):
        """int8_quant -> three tcgen05 GEMMs with the dequantising epilogue (the LN1 output is quantised once)."""
        self._quant_rows(bf.n, bf.xq, bf.xs, bf.R, self.hidden_dim_)
        self._gemm_deq(bf, bf.xq, bf.xs, L.wq, L.wq_deq, bf.qp)
        self._gemm_deq(bf, bf.xq, bf.xs, L.wk, L.wk_deq, bf.kp)
        self._gemm_deq(bf, bf.xq, bf.xs, L.wv, L.wv_deq, bf.vp)

    def _o_proj(self, bf, L, x, out):
        self._quant_rows(x, bf.xq, bf.xs, bf.R, self.hidden_dim_)
        self._gemm_deq(bf, bf.xq, bf.xs, L.wo, L.wo_deq, out)

    def _embedding_table(self):
        return self.embedding, 1, float(self.emb_qscale)

    GEMM_LOGITS_MIN_ROWS = 9  # more rows than one pass of the GEMV kernel: logits become an int8 GEMM

    def _head(self, x_rows):
        """Batches of more than 8 rows compute the logits on the tensor cores: rows are quantised like every
        other activation of this decoder (int8_quant.cpp:59-64, 15-28) and multiplied with the transposed int8
        embedding table by the tcgen05 GEMM (dequantising epilogue); up to 8 rows keep the exact f32 x int8
        GEMV with the sampler folded in."""
        B = self._batch
        if B < self.GEMM_LOGITS_MIN_ROWS:
            return super()._head(x_rows)
        V, hid, dev = self.vocab_size_, self.hidden_dim_, self.device
        Vp = (V + 15) // 16 * 16
        if getattr(self, "_emb_T", None) is None or self._emb_T.shape != (hid, Vp):
            self._emb_T = torch.zeros((hid, Vp), dtype=torch.int8, device=dev)
            self._emb_T[:, :V] = self.embedding.t()
            self._pad_bias = torch.zeros(Vp, dtype=torch.float32, device=dev)
            self._pad_bias[V:] = float("-inf")  # padded columns never win the argmax
        if getattr(self, "_logits_pad", None) is None or self._logits_pad.shape != (B, Vp):
            self._logits_pad = torch.empty((B, Vp), dtype=torch.float32, device=dev)
            self._lq = torch.empty((B, hid), dtype=torch.int8, device=dev)
            self._ls = torch.empty(B, dtype=torch.float32, device=dev)
        lib, s = self._lib, _cabi.stream()
        self._quant_rows(x_rows, self._lq, self._ls, B, hid)
        wp, wb = self._scratch(self.bufs, lib.pa_gemm_i8_workspace_bytes(1, B, Vp, hid))
        self._chk(lib.pa_gemm_i8_dequant(self._lq.data_ptr(), self._emb_T.data_ptr(), self._logits_pad.data_ptr(), 1,
                                         B, Vp, hid, self._ls.data_ptr(), 1.0 / float(self.emb_qscale),
                                         self._pad_bias.data_ptr(), _cabi.ACT[""], wp, wb, s), "pa_gemm_i8_dequant")
        self.logits = self._logits_pad[:, :V]
        sp = getattr(self, "_sampling", None)
        if sp is None:
            self._chk(lib.pa_argmax_f32(self._logits_pad.data_ptr(), B, Vp, self._temperature, self.ARGMAX_DIVIDE,
                                        self.ids.data_ptr(), s), "pa_argmax_f32")
            return
        from . import sampling
        probs = sampling.softmax_temperature(self.logits.contiguous(), 1.0 / self._temperature)
        sampling.apply_topk_topp_filter(probs, sp["top_k"], sp["top_p"], sp["eos_token_id"], sp["eos_thresh"])
        u = torch.rand(B, device=dev, generator=sp["generator"])
        sampling.sample_from_probs(probs, u, out=self.ids)

    def _embed(self, bf, ids):
        self._chk(self._lib.pa_embedding_i8(self.embedding.data_ptr(), float(self.emb_qscale), ids.data_ptr(), bf.R,
                                            self.hidden_dim_, self.vocab_size_, bf.x.data_ptr(), _cabi.stream()),
                  "pa_embedding_i8")

    def _quant_rows(self, x, q, scales, R, dim):
        """compute_minmax_scale + batch_quantize per row (int8_quant.cpp:59-64, 15-28), one kernel."""
        self._chk(self._lib.pa_row_quantize_dynamic_i8(x.data_ptr(), R, dim, scales.data_ptr(), q.data_ptr(),
                                                       _cabi.stream()), "pa_row_quantize_dynamic_i8")

    def _ln2_mlp(self, bf, L):
        """LN2 -> int8_quant -> fc1 (relu) -> int8_quant -> fc2 with both quantisations fused into their producers:
        pa_layer_norm_quantize_i8 (the normalised row never goes to memory as f32) and pa_gemm_i8_dynquant (fc1's
        accumulators wait in TMEM while the row maxima cross the grid; no 4 x inter bytes per row written and re-read).
        Same bits as the unfused kernels, which remain the path for shapes the fused GEMM does not take (few rows, more
        than one wave of tiles)."""
        lib, R, hid, inter, s = self._lib, bf.R, self.hidden_dim_, self.inter_dim_, _cabi.stream()
        if getattr(bf, "dq_ok", None) is None:  # decided once per buffer set (before any graph capture: warm-up step)
            bf.dq_ok = os.environ.get("PA_MLP_FUSED_QUANT", "1") != "0" and R > 128
        if not bf.dq_ok:
            return super()._ln2_mlp(bf, L)
        wp, wb = self._scratch(bf, max(lib.pa_gemm_i8_workspace_bytes(1, R, hid, inter), lib.pa_gemm_i8_workspace_bytes(1, R, inter, hid)))
        if getattr(bf, "dq_ws", None) is None:  # dedicated, zeroed once: the grid barrier's self-resetting counters live here
            bf.dq_ws = torch.zeros(lib.pa_gemm_i8_dynquant_workspace_bytes(1, R, inter), dtype=torch.uint8, device=self.device)
        self._chk(lib.pa_layer_norm_quantize_i8(bf.a.data_ptr(), L.ln2_g.data_ptr(), L.ln2_b.data_ptr(), R, hid, self.eps,
                                                None, bf.xs.data_ptr(), bf.xq.data_ptr(), s), "pa_layer_norm_quantize_i8")
        st = lib.pa_gemm_i8_dynquant(bf.xq.data_ptr(), L.fc1_w.data_ptr(), bf.hq.data_ptr(), bf.hs.data_ptr(), 1, R, inter,
                                     hid, bf.xs.data_ptr(), float(L.fc1_deq), L.fc1_b.data_ptr(), _cabi.ACT["relu"],
                                     bf.dq_ws.data_ptr(), bf.dq_ws.numel(), s)
        if st == -2:  # PA_ERR_UNSUPPORTED: this shape does not fit one wave -> unfused fc1 + quantise, same bits
            bf.dq_ok = False
            self._chk(lib.pa_gemm_i8_dequant(bf.xq.data_ptr(), L.fc1_w.data_ptr(), bf.h.data_ptr(), 1, R, inter, hid,
                                             bf.xs.data_ptr(), float(L.fc1_deq), L.fc1_b.data_ptr(), _cabi.ACT["relu"],
                                             wp, wb, s), "pa_gemm_i8_dequant")
            self._quant_rows(bf.h, bf.hq, bf.hs, R, inter)
        else:
            self._chk(st, "pa_gemm_i8_dynquant")
        self._chk(lib.pa_gemm_i8_dequant(bf.hq.data_ptr(), L.fc2_w.data_ptr(), bf.x.data_ptr(), 1, R, hid, inter,
                                         bf.hs.data_ptr(), float(L.fc2_deq), L.fc2_b.data_ptr(), _cabi.ACT[""], wp, wb, s),
                  "pa_gemm_i8_dequant")

    def _mlp(self, bf, L):
        lib, R, hid, inter, s = self._lib, bf.R, self.hidden_dim_, self.inter_dim_, _cabi.stream()
        wp, wb = self._scratch(bf, max(lib.pa_gemm_i8_workspace_bytes(1, R, inter, hid),
                                       lib.pa_gemm_i8_workspace_bytes(1, R, hid, inter)))
        self._quant_rows(bf.n, bf.xq, bf.xs, R, hid)
        self._chk(lib.pa_gemm_i8_dequant(bf.xq.data_ptr(), L.fc1_w.data_ptr(), bf.h.data_ptr(), 1, R, inter, hid,
                                         bf.xs.data_ptr(), float(L.fc1_deq), L.fc1_b.data_ptr(),
                                         _cabi.ACT["relu"], wp, wb, s), "pa_gemm_i8_dequant")
        self._quant_rows(bf.h, bf.hq, bf.hs, R, inter)
        self._chk(lib.pa_gemm_i8_dequant(bf.hq.data_ptr(), L.fc2_w.data_ptr(), bf.x.data_ptr(), 1, R, hid, inter,
                                         bf.hs.data_ptr(), float(L.fc2_deq), L.fc2_b.data_ptr(), _cabi.ACT[""], wp, wb, s),
                  "pa_gemm_i8_dequant")

    def _logits(self, x_rows):
        self._chk(self._lib.pa_logits_i8(x_rows.data_ptr(), self.embedding.data_ptr(), float(self.emb_qscale),
                                         self._batch, self.hidden_dim_, self.vocab_size_, self.logits.data_ptr(),
                                         _cabi.stream()), "pa_logits_i8")
