#!/usr/bin/env python
"""Prefill attention timing: one layer, Llama-7B head shape (32 heads, D = 128), fp16
and int8 pages, B prompts of Tq tokens, causal.  Reports ms per layer and achieved attention FLOP/s of the
tensor-core flash-attention kernel (PA_PREFILL_FA=0 times the row-per-query path instead)."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "pagedattention-based-transformer-decoder-inference-framework_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)


def main():
    import llm_decoder as ld
    dev = torch.device("cuda", 0)
    H, D, TILE = 32, int(os.environ.get("HEAD_DIM", "128")), 16   # HEAD_DIM=64: GPT-2-style heads
    res = []
    shapes = ((1, 512), (1, 2048), (4, 2048), (2, 8192))
    kvs = ("f16", "i8")
    if len(sys.argv) > 1:  # prefill.py B,Tq [kv]: one shape (for ncu captures)
        shapes = (tuple(int(v) for v in sys.argv[1].split(",")),)
        kvs = (sys.argv[2],) if len(sys.argv) > 2 else kvs
    for B, Tq in shapes:
        nt = Tq // TILE
        P = B * H * nt
        g = torch.Generator(device=dev).manual_seed(3)
        out_line = {"B": B, "Tq": Tq}
        for kv in kvs:
            kvc = ld.KVTileCache(kv, device=dev)
            if kv == "f16":
                kvc.adopt_buffers(torch.randn((P, TILE, D), generator=g, device=dev, dtype=torch.float16),
                                  torch.randn((P, TILE, D), generator=g, device=dev, dtype=torch.float16))
            else:
                kvc.adopt_buffers(torch.randint(-127, 128, (P, TILE, D), generator=g, device=dev, dtype=torch.int8),
                                  torch.randint(-127, 128, (P, TILE, D), generator=g, device=dev, dtype=torch.int8),
                                  torch.rand((P, TILE), generator=g, device=dev) * 20 + 30,
                                  torch.rand((P, TILE), generator=g, device=dev) * 20 + 30)
            kvc.configure_table(B, H, nt)
            kvc.page_table_.load_host_table(torch.randperm(P, generator=g, device=dev).to(torch.int32).cpu().numpy().reshape(B, H, nt))
            q = torch.randn((B, H, Tq, D), generator=g, device=dev)
            out = torch.empty_like(q)
            for _ in range(2):
                ld.paged_prefill(q, out, kvc, B, Tq, float(np.sqrt(D)))
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            n = 10
            for _ in range(n):
                ld.paged_prefill(q, out, kvc, B, Tq, float(np.sqrt(D)))
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / n
            flops = 4.0 * B * H * D * Tq * (Tq + 1) / 2
            tc = os.environ.get("PA_PREFILL_TC", "1") != "0"
            out_line[("tcgen05_" if tc else "flash_mma_") + kv] = {
                "ms": round(ms, 3), "tflops": round(flops / ms / 1e9, 1)}
        res.append(out_line)
    print(json.dumps({"workload": f"prefill attention, one layer, 32 heads x D={D}, causal", "results": res}))


if __name__ == "__main__":
    main()
