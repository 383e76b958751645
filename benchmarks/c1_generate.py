#!/usr/bin/env python
"""C1: GPT-2-small shape (12 layers, 12 heads, head_dim 64, hidden 768, vocab 50257), batch 1, 512-token
context, 16-token pages, random-init weights (BASELINE.json configs[0], the reference's own
CPU-runnable case).  Reports decode tokens/s of INT8Decoder / CUDADecoder.generate on the GPU (CUDA-graph
steps, one host synchronisation per call) next to the reference's CPU path on the host cores:
the oracle port of cpu_paged_attention_forward<int8> + oneDNN s8 GEMMs (torch._int_mm on CPU, the
only linkable oneDNN here) for the two MLP layers, per layer, at the same context length."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "pagedattention-based-transformer-decoder-inference-framework_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

L, H, D, HID, V, CTX = 12, 12, 64, 768, 50257, 512


def cpu_reference_layer_seconds(ctx):
    """One decode step of ONE layer on the host: int8 paged attention (oracle port, OpenMP over heads) +
    quantise + fc1 + fc2 as oneDNN s8 GEMMs (M = 1)."""
    import oracle
    oracle.cpu.build()
    c = oracle.cpu
    rng = np.random.default_rng(3)
    nt = ctx // 16
    P = H * nt
    kq = rng.integers(-127, 128, (P, 16, D), dtype=np.int8)
    vq = rng.integers(-127, 128, (P, 16, D), dtype=np.int8)
    ks = (rng.random((P, 16)) * 20 + 30).astype(np.float32)
    q = rng.standard_normal((1, H, D)).astype(np.float32)
    table = rng.permutation(P).astype(np.int32).reshape(1, H, nt)
    x8 = torch.randint(-127, 128, (1, HID), dtype=torch.int8)
    w1 = torch.randint(-127, 128, (HID, 4 * HID), dtype=torch.int8)
    w2 = torch.randint(-127, 128, (4 * HID, HID), dtype=torch.int8)
    h8 = torch.randint(-127, 128, (1, 4 * HID), dtype=torch.int8)

    def step():
        c.paged_attention(q, kq, vq, table, num_beams=1, num_tiles=nt, tile_size=16, T=ctx, temperature=1.0,
                          k_scales=ks, v_scales=ks)
        torch._int_mm(x8.expand(32, HID).contiguous()[:1] if False else x8, w1) if False else None
        # torch._int_mm on CPU requires M > 16: time M = 32 and divide (weight-streaming bound either way)
        torch._int_mm(x8.expand(32, HID).contiguous(), w1)
        torch._int_mm(h8.expand(32, 4 * HID).contiguous(), w2)

    for _ in range(3):
        step()
    t0 = time.perf_counter()
    n = 20
    for _ in range(n):
        step()
    return (time.perf_counter() - t0) / n, c.num_threads()


def main():
    import llm_decoder as ld
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    res = {}
    gen_tokens = 64
    prompt = [int(t) for t in np.random.default_rng(1).integers(0, V, CTX - gen_tokens)]
    for name, cls in (("INT8Decoder", ld.INT8Decoder), ("CUDADecoder", ld.CUDADecoder)):
        dec = cls(L, H, D, HID, V, CTX)
        g = torch.Generator(device=dev).manual_seed(11)
        # random-init weights directly on the device (the file loaders are exercised by the tests)
        if name == "CUDADecoder":
            dec.embedding.normal_(generator=g)
            for Ly in dec.layers:
                for w in (Ly.fc1_w, Ly.fc2_w):
                    w.normal_(generator=g)
                    w.mul_(0.03)
        else:
            dec.embedding.copy_(torch.randint(-127, 128, dec.embedding.shape, generator=g, device=dev, dtype=torch.int8))
            for Ly in dec.layers:
                Ly.fc1_w.copy_(torch.randint(-127, 128, Ly.fc1_w.shape, generator=g, device=dev, dtype=torch.int8))
                Ly.fc2_w.copy_(torch.randint(-127, 128, Ly.fc2_w.shape, generator=g, device=dev, dtype=torch.int8))
                Ly.fc1_deq = Ly.fc2_deq = 0.03 / 127
        dec.generate(prompt, 8, 1.0)  # warm-up: lazy init + graph capture
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = dec.generate(prompt, gen_tokens, 1.0)
        torch.cuda.synchronize()
        t_all = time.perf_counter() - t0
        t1 = time.perf_counter()
        dec.generate(prompt, 1, 1.0)   # prefill + first token only
        torch.cuda.synchronize()
        t_prefill = time.perf_counter() - t1
        assert len(out) == CTX
        res[name] = {"decode_tok_s": round((gen_tokens - 1) / max(t_all - t_prefill, 1e-9), 1),
                     "prefill_ms": round(t_prefill * 1e3, 2), "generate_64_tokens_ms": round(t_all * 1e3, 2)}
        # the same model serving 64 sequences at once (rows are independent; the step is latency bound at batch 1)
        NB = 64
        prompts = [[int(t) for t in np.random.default_rng(100 + i).integers(0, V, CTX - gen_tokens)] for i in range(NB)]
        dec.generate_batch(prompts, 8, 1.0)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        dec.generate_batch(prompts, gen_tokens, 1.0)
        torch.cuda.synchronize()
        tb_all = time.perf_counter() - t0
        t1 = time.perf_counter()
        dec.generate_batch(prompts, 1, 1.0)
        torch.cuda.synchronize()
        tb_prefill = time.perf_counter() - t1
        res[name]["batch64_decode_tok_s"] = round(NB * (gen_tokens - 1) / max(tb_all - tb_prefill, 1e-9), 1)
        res[name]["batch64_prefill_ms"] = round(tb_prefill * 1e3, 2)
    t_layer, threads = cpu_reference_layer_seconds(CTX)
    res["cpu_reference"] = {"decode_tok_s": round(1.0 / (t_layer * L), 1), "cores": threads, "kind": "port",
                            "sample": f"one layer-step (int8 paged attention over {CTX} tokens + two oneDNN s8 GEMMs), x{L} layers; "
                                      "embedding/LayerNorm/logits not counted"}
    print(json.dumps({"workload": "C1: GPT-2-small shape, batch 1, 512-token ctx, 16-token pages, random-init",
                      "results": res}))


if __name__ == "__main__":
    main()
