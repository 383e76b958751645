#!/usr/bin/env python
"""One GPU's share of C5 (32 heads x 16384 tokens = 268 MB of fp16 K/V) on ONE GPU: the streaming part of the
split-KV step without the exchange (pa_paged_decode_f16_partial), for kernel experiments.

    [PA_DECODE_STATIC=0|1] [PA_DECODE_MERGE_KERNEL=0|1] [PA_PARTIAL_DIRECT=1] python benchmarks/c5_share.py
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "pagedattention-based-transformer-decoder-inference-framework_b200"), os.path.join(ROOT, "benchmarks")):
    if p not in sys.path:
        sys.path.insert(0, p)
import extras  # noqa: E402
import llm_decoder as ld  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    H, D, TILE = 32, 128, 16
    T = int(os.environ.get("C5_T", 16384))
    nt = T // TILE
    P = H * nt
    g = torch.Generator(device=dev).manual_seed(77)
    caches = []
    for _ in range(4):
        k = torch.randn((P, TILE, D), generator=g, device=dev, dtype=torch.float16)
        v = torch.randn((P, TILE, D), generator=g, device=dev, dtype=torch.float16)
        kvc = ld.KVTileCache("f16", device=dev)
        kvc.adopt_buffers(k, v)
        kvc.configure_table(1, H, nt)
        kvc.page_table_.load_host_table(torch.randperm(P, generator=g, device=dev).to(torch.int32).cpu().numpy().reshape(1, H, nt))
        caches.append(kvc)
    q = torch.randn((1, H, D), device=dev)
    temp = float(np.sqrt(D))
    us, us_min = extras.graph_time([lambda c=c: ld.paged_decode_partial(q, c, 1, T, temp) for c in caches], 20, dev, per=4)
    from llm_decoder import dist as pd
    ex = pd.PeerExchange(H, D)   # a world of one rank: the exchange tail without NVLink
    us_f, us_f_min = extras.graph_time([lambda c=c: pd.split_kv_decode(q, c, 1, T, temp, exchange=ex, fused=True) for c in caches],
                                       20, dev, per=4)
    ex.check()
    kv_bytes = H * T * D * 2 * 2
    env = {k: os.environ[k] for k in ("PA_DECODE_STATIC", "PA_DECODE_MERGE_KERNEL", "PA_PARTIAL_DIRECT") if k in os.environ}
    print(json.dumps({"env": env, "T": T, "partial_us": round(us, 2), "partial_us_min": round(us_min, 2), "fused_world1_us": round(us_f, 2), "fused_world1_us_min": round(us_f_min, 2), "gbs": round(kv_bytes / us / 1e3, 1)}))


if __name__ == "__main__":
    main()
