#!/usr/bin/env python
"""One short invocation of each hot kernel at its BASELINE.json shape, for ncu (profiles/r02_*):

    python benchmarks/profile_targets.py c2|c4|c3|gemm|c5

Each target warms up, then launches the kernel(s) 3 times on rotating data; ncu picks launches with -k / -s / -c."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "pagedattention-based-transformer-decoder-inference-framework_b200"), os.path.join(ROOT, "benchmarks")):
    if p not in sys.path:
        sys.path.insert(0, p)
import llm_decoder as ld  # noqa: E402
from llm_decoder import _cabi  # noqa: E402

H, D, TILE = 32, 128, 16
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
g = torch.Generator(device=dev).manual_seed(7)
temp = float(np.sqrt(D))


def cache(kind, B, T, nsets=2):
    nt = T // TILE
    P = B * H * nt
    out = []
    for _ in range(nsets):
        if kind == "i8":
            k = torch.randint(-127, 128, (P, TILE, D), generator=g, device=dev, dtype=torch.int8)
            v = torch.randint(-127, 128, (P, TILE, D), generator=g, device=dev, dtype=torch.int8)
            sc = [torch.rand((P, TILE), generator=g, device=dev) * 20 + 30 for _ in range(2)]
        else:
            k = torch.randn((P, TILE, D), generator=g, device=dev, dtype=torch.float16)
            v = torch.randn((P, TILE, D), generator=g, device=dev, dtype=torch.float16)
            sc = []
        c = ld.KVTileCache(kind, device=dev)
        c.adopt_buffers(k, v, *sc)
        c.configure_table(B, H, nt)
        c.page_table_.load_host_table(torch.randperm(P, generator=g, device=dev).to(torch.int32).cpu().numpy().reshape(B, H, nt))
        out.append(c)
    return out


def run_decode(kind, B, T):
    cs = cache(kind, B, T)
    q = torch.randn((B, H, D), generator=g, device=dev)
    o = torch.empty_like(q)
    for i in range(5):
        ld.AttentionCUDA.forward(q, o, B, H, D, T, None, cs[i % 2], None, False, kind == "f16", True, temp)
    torch.cuda.synchronize()


def run_c3():
    import extras
    print(extras.c3_group(dev, 6547.2, iters=2, sets=2))


def run_gemm():
    import extras
    print(extras.c4_gemm_pair(dev, iters=2))


def run_c5():
    from llm_decoder import dist as pd
    cs = cache("f16", 1, 16384, nsets=3)
    q = torch.randn((1, H, D), generator=g, device=dev)
    ex = pd.PeerExchange(H, D)
    for i in range(6):
        pd.split_kv_decode(q, cs[i % 3], 1, 16384, temp, exchange=ex, fused=True)
    torch.cuda.synchronize()
    ex.check()
    ex.close()


if __name__ == "__main__":
    t = sys.argv[1]
    {"c2": lambda: run_decode("f16", 64, 4096), "c4": lambda: run_decode("i8", 256, 4096), "c3": run_c3, "gemm": run_gemm,
     "c5": run_c5}[t]()
    print("done", t)
