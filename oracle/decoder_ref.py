"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the decoder layer loop, built from the oracle's
pinned pieces.  Every STAGE is pinned bit-exact against the reference's own compiled code
(embedding lookup, LayerNorm<float>, MLP<float>, cpu_paged_attention_forward<float>, int8_quant:
tests/test_oracle_pinning.py); the WIRING between them is not: decoder_block.hpp / cuda_decoder.cu /
int8_decoder.cpp do not compile and never append K/V, so this file restates the
decisions llm_decoder/decoders.py documents, step by step, as an independent CPU path.

  decoder/token_embedding.hpp:19-26  embedding lookup
  decoder/decoder_block.hpp:41-62    LN1 -> attention(q = LN1 out) -> LN2 -> MLP, no residuals
  decoder/layer_norm.hpp:20-37       oracle.cpu.layer_norm
  decoder/mlp.hpp:23-41              oracle.cpu.mlp_f32 (float) / int8_quant + dnnl_matmul_int8 (int8)
  attention_cpu/cpu_attention_kernel.cpp:36-129  oracle.cpu.paged_attention
  decoder/cuda_decoder.cu:7-14, int8_decoder.cpp:97-104  greedy argmax (first maximum)
  weights/README.md:31-34            optional attn_wq / wk / wv / wo projections ([hidden, hidden], head-major columns)
"""
import numpy as np

from . import cpu


def _pages_from_rows(rows, H, D, tile, dtype):
    """rows: list of [H, D] arrays (one per cached token) -> pool [H*nt, tile, D], table [1,H,nt]."""
    t = len(rows)
    nt = max(1, (t + tile - 1) // tile)
    pool = np.zeros((H * nt, tile, D), dtype=dtype)
    for i, r in enumerate(rows):
        for h in range(H):
            pool[h * nt + i // tile, i % tile] = r[h]
    table = np.arange(H * nt, dtype=np.int32).reshape(1, H, nt)
    return pool, table, nt


class RefDecoder:
    """Teacher-forced CPU decoder: feed tokens one at a time, get the logits of each step."""

    def __init__(self, weights, H, D, tile=16, int8=False, attn_temperature=1.0, eps=1e-5):
        self.w, self.H, self.D, self.tile, self.int8 = weights, H, D, tile, int8
        self.attn_temperature, self.eps = attn_temperature, eps
        L = len(weights["layers"])
        self.k_rows = [[] for _ in range(L)]   # per layer: list of [H, D] (fp16-rounded f32, or int8)
        self.v_rows = [[] for _ in range(L)]
        self.k_scales = [[] for _ in range(L)]  # int8: list of [H] scales
        self.v_scales = [[] for _ in range(L)]

    def _lin_i8(self, x, wq, deq, bias=None, relu=False):
        """int8_quant -> exact int32 GEMM -> dequantising epilogue (the chain pa_gemm_i8_dequant implements)."""
        s = cpu.batch_minmax_scale(x, x.size)
        xq = cpu.batch_quantize(x, s, x.size).reshape(1, 1, -1)
        acc = cpu.gemm_s8s8s32(xq, wq[None])[0, 0].astype(np.float32)
        alpha = np.float32(deq) / np.float32(s[0])
        v = (alpha * acc).astype(np.float32) + (np.float32(0) if bias is None else bias.astype(np.float32))
        return np.maximum(v, np.float32(0)) if relu else v.astype(np.float32)

    def _proj(self, L, name, x):
        """Optional attention projection (weights/README.md:31-34): x . W, fp32 or through the int8 chain."""
        if self.int8:
            return self._lin_i8(x, L[name], L[name + "_deq"])
        return (x.astype(np.float64) @ L[name].astype(np.float64)).astype(np.float32)

    def _attend(self, li, n1):
        H, D, tile = self.H, self.D, self.tile
        L = self.w["layers"][li]
        proj = L.get("wq") is not None
        qv, kv_, vv = (self._proj(L, "wq", n1), self._proj(L, "wk", n1), self._proj(L, "wv", n1)) if proj else (n1, n1, n1)
        if not self.int8:
            self.k_rows[li].append(kv_.reshape(H, D).astype(np.float16).astype(np.float32))
            self.v_rows[li].append(vv.reshape(H, D).astype(np.float16).astype(np.float32))
            kpool, table, nt = _pages_from_rows(self.k_rows[li], H, D, tile, np.float32)
            vpool, _, _ = _pages_from_rows(self.v_rows[li], H, D, tile, np.float32)
            out = cpu.paged_attention(qv.reshape(1, H, D), kpool, vpool, table, num_beams=1, num_tiles=nt,
                                      tile_size=tile, T=len(self.k_rows[li]),
                                      temperature=self.attn_temperature).reshape(-1)
        else:
            pools = []
            for rows, scales, x in ((self.k_rows[li], self.k_scales[li], kv_), (self.v_rows[li], self.v_scales[li], vv)):
                sc = cpu.batch_minmax_scale(x, D)                   # one scale per (token, head) row
                rows.append(cpu.batch_quantize(x, sc, D).reshape(H, D))
                scales.append(sc.reshape(H))
                pool, table, nt = _pages_from_rows(rows, H, D, tile, np.int8)
                spool = np.ones((H * nt, tile), dtype=np.float32)
                for i, s in enumerate(scales):
                    for h in range(H):
                        spool[h * nt + i // tile, i % tile] = s[h]
                pools.append((pool, spool))
            out = cpu.paged_attention(qv.reshape(1, H, D), pools[0][0], pools[1][0], table, num_beams=1, num_tiles=nt,
                                      tile_size=tile, T=len(self.k_rows[li]), temperature=self.attn_temperature,
                                      k_scales=pools[0][1], v_scales=pools[1][1]).reshape(-1)
        return self._proj(L, "wo", out) if proj else out

    def _mlp_i8(self, n2, L):
        h = self._lin_i8(n2, L["fc1_w"], L["fc1_deq"], L["fc1_b"], True)
        return self._lin_i8(h.astype(np.float32), L["fc2_w"], L["fc2_deq"], L["fc2_b"], False)

    def step(self, token):
        w = self.w
        if self.int8:
            x = w["embedding"][token].astype(np.float32) / np.float32(w["emb_qscale"])
        else:
            x = w["embedding"][token].astype(np.float32)
        for li, L in enumerate(w["layers"]):
            n1 = cpu.layer_norm(x[None], L["ln1_g"], L["ln1_b"], self.eps)[0]
            a = self._attend(li, n1)
            n2 = cpu.layer_norm(a[None], L["ln2_g"], L["ln2_b"], self.eps)[0]
            if self.int8:
                x = self._mlp_i8(n2, L)
            else:
                x = cpu.mlp_f32(n2[None], L["fc1_w"], L["fc1_b"], L["fc2_w"], L["fc2_b"])[0]
        if self.int8:
            logits = (w["embedding"].astype(np.float32) @ x.astype(np.float32)) / np.float32(w["emb_qscale"])
        else:
            logits = w["embedding"].astype(np.float32) @ x.astype(np.float32)
        return logits.astype(np.float32)


def sample(logits, temperature, divide):
    """cuda_decoder.cu:7-14 (divide) / int8_decoder.cpp:97-104 (multiply); first maximum."""
    t = np.float32(temperature)
    v = logits / t if divide else logits * t
    return int(np.argmax(v))
