#!/usr/bin/env python
"""C3: beam width 4 with shared-prefix pages, 32 groups (128 rows), 2K ctx = 1792 shared + 256
private tokens per beam (BASELINE.json configs[2]).  Reports achieved GB/s on UNIQUE
bytes (the algorithmic figure) and on logical bytes for each kernel."""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "pagedattention-based-transformer-decoder-inference-framework_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)


def build_case(dev, groups, W, H, D, T, shared, tile, seed):
    g = torch.Generator(device=dev).manual_seed(seed)
    nt, pt = T // tile, shared // tile
    B = groups * W
    unique = groups * H * pt + B * H * (nt - pt)
    perm = torch.randperm(unique, generator=g, device=dev).to(torch.int32)
    table = torch.empty((B, H, nt), dtype=torch.int32, device=dev)
    sh = perm[:groups * H * pt].reshape(groups, 1, H, pt).expand(groups, W, H, pt).reshape(B, H, pt)
    table[:, :, :pt] = sh
    table[:, :, pt:] = perm[groups * H * pt:].reshape(B, H, nt - pt)
    k = torch.randn((unique, tile, D), generator=g, device=dev, dtype=torch.float16)
    v = torch.randn((unique, tile, D), generator=g, device=dev, dtype=torch.float16)
    return k, v, table, unique


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=30)
    ap.add_argument("--sets", type=int, default=3)
    args = ap.parse_args()
    import llm_decoder as ld
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    groups, W, H, D, T, shared, tile = 32, 4, 32, 128, 2048, 1792, 16
    B = groups * W
    temp = float(np.sqrt(D))
    caches = []
    for i in range(args.sets):
        k, v, table, unique = build_case(dev, groups, W, H, D, T, shared, tile, 1237 + i)
        kvc = ld.KVTileCache("f16", device=dev)
        kvc.adopt_buffers(k, v)
        kvc.configure_table(B, H, T // tile)
        kvc.page_table_.load_host_table(table.cpu().numpy())
        caches.append(kvc)
    q = torch.randn((B, H, D), device=dev)
    out = torch.empty_like(q)
    # beam_ids: identity permutation given explicitly = the reference's beam indirection path
    beam_ids = torch.arange(B, dtype=torch.int32, device=dev)
    unique_bytes = unique * tile * D * 2 * 2 + 2 * B * H * D * 4 + B * H * (T // tile) * 4
    logical_bytes = B * H * T * D * 2 * 2 + 2 * B * H * D * 4 + B * H * (T // tile) * 4
    res = {}
    variants = [("fused(direct)", dict(use_overlap=False)), ("overlap", dict(use_overlap=True))]
    if hasattr(ld, "paged_decode_group"):
        variants.append(("group", None))
    outs = {}
    for name, kw in variants:
        def run(c):
            if kw is None:
                ld.paged_decode_group(q, out, c, B, T, W, temp, beam_ids=beam_ids)
            else:
                ld.AttentionCUDA.forward(q, out, B, H, D, T, beam_ids, c, None, False, True, kw["use_overlap"], temp)
        for c in caches:
            run(c)
        torch.cuda.synchronize()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.iters)]
        for i, (a, b) in enumerate(ev):
            a.record()
            run(caches[i % len(caches)])
            b.record()
        torch.cuda.synchronize()
        ts = sorted(a.elapsed_time(b) for a, b in ev)
        med = ts[len(ts) // 2]
        run(caches[0])
        outs[name] = out.clone()
        res[name] = {"ms_median": round(med, 4), "ms_min": round(ts[0], 4),
                     "gbs_unique": round(unique_bytes / med / 1e6, 1), "gbs_logical": round(logical_bytes / med / 1e6, 1)}
    names = list(outs)
    diff = max(float((outs[names[0]] - outs[n]).abs().max()) for n in names[1:])
    print(json.dumps({"workload": "C3: 32 groups x 4 beams, 32 heads, D=128, ctx 2048 (1792 shared + 256 private), fp16",
                      "unique_bytes": unique_bytes, "logical_bytes": logical_bytes, "kernels": res,
                      "max_abs_diff_between_kernels": diff,
                      "l2": f"{args.sets} page sets of {unique_bytes >> 20} MiB used round-robin"}))


if __name__ == "__main__":
    main()
