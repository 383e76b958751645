#!/usr/bin/env python
"""C4: INT8 KV cache + INT8 MLP GEMMs, Llama-7B shape (32 heads, D=128, hidden 4096, MLP 4x), batch
256 decode rows in total sharded over N GPUs (BASELINE.json configs[3]).

    python benchmarks/c4_int8.py                       # 1 GPU, 256 rows
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29535 benchmarks/c4_int8.py      # strong scaling: 256 / N rows per GPU
    ... c4_int8.py --weak                              # weak scaling: 256 rows per GPU

One layer-step = fused-quantise append (pa_kv_append_f32_i8) + int8-KV paged decode over 4096 tokens
(pa_paged_decode_i8_overlap) + per-row activation quantise + fc1 (relu, s8 out) + fc2 (s8 out) on the
tcgen05 kind::i8 GEMM.  Each phase is timed as a CUDA-graph replay over rotating page/weight sets.
"""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "pagedattention-based-transformer-decoder-inference-framework_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)


def graph_time(fns, iters, dev, world):
    """fns: list of callables (one per data set); captured round-robin into one graph of 2*len(fns) calls."""
    per = 2 * len(fns)
    for i in range(per):
        fns[i % len(fns)]()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(per):
            fns[i % len(fns)]()
    g.replay()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / per)
    t = torch.tensor(ts, device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.median().item())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=256, help="total decode rows (strong) or rows per GPU (--weak)")
    ap.add_argument("--ctx", type=int, default=4096)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--weak", action="store_true")
    args = ap.parse_args()
    world, rank, local = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    import llm_decoder as ld
    from llm_decoder import _cabi
    from llm_decoder.dist import shard_range

    H, D, TILE, HID, INTER = 32, 128, 16, 4096, 16384
    T = args.ctx
    if args.weak:
        B = args.rows
    else:
        r0, r1 = shard_range(args.rows, world, rank)
        B = r1 - r0
    nt = T // TILE
    P = B * H * nt
    g = torch.Generator(device=dev).manual_seed(1238 + rank)
    pool_bytes = 2 * P * TILE * D + 2 * P * TILE * 4
    nsets = max(2, min(4, int((torch.cuda.mem_get_info()[0] - (24 << 30)) // pool_bytes)))
    if B * H * T * D * 2 >= (3 * 126 << 20):
        nsets = min(nsets, 2)  # one set already exceeds L2 several times over
    caches = []
    for i in range(nsets):
        k = torch.randint(-127, 128, (P, TILE, D), generator=g, device=dev, dtype=torch.int8)
        v = torch.randint(-127, 128, (P, TILE, D), generator=g, device=dev, dtype=torch.int8)
        ks = torch.rand((P, TILE), generator=g, device=dev) * 20 + 30
        vs = torch.rand((P, TILE), generator=g, device=dev) * 20 + 30
        kvc = ld.KVTileCache("i8", device=dev)
        kvc.adopt_buffers(k, v, ks, vs)
        kvc.configure_table(B, H, nt)
        kvc.page_table_.load_host_table(torch.randperm(P, generator=g, device=dev).to(torch.int32).cpu().numpy().reshape(B, H, nt))
        caches.append(kvc)
    q = torch.randn((B, H, D), generator=g, device=dev)
    out = torch.empty_like(q)
    nk = torch.randn((B, H, D), generator=g, device=dev)
    nv = torch.randn((B, H, D), generator=g, device=dev)
    pos = torch.full((B,), T - 1, dtype=torch.int32, device=dev)
    temp = float(np.sqrt(D))
    res = {}
    res["append_quant_us"] = 1e3 * graph_time([lambda c=c: c.append(nk, nv, pos) for c in caches], args.iters, dev, world)
    res["decode_i8_us"] = 1e3 * graph_time(
        [lambda c=c: ld.AttentionCUDA.forward(q, out, B, H, D, T, None, c, None, False, False, True, temp) for c in caches],
        args.iters, dev, world)
    # MLP: per-row quantise -> fc1 (relu, s8) -> fc2 (s8)
    nw = 4
    W1 = [torch.randint(-127, 128, (1, HID, INTER), generator=g, device=dev, dtype=torch.int8) for _ in range(nw)]
    W2 = [torch.randint(-127, 128, (1, INTER, HID), generator=g, device=dev, dtype=torch.int8) for _ in range(nw)]
    b1, b2 = torch.randn(INTER, device=dev), torch.randn(HID, device=dev)
    x = torch.randn((B, HID), generator=g, device=dev)
    xq = torch.empty((1, B, HID), dtype=torch.int8, device=dev)
    xs = torch.empty(B, device=dev)
    h8 = torch.empty((1, B, INTER), dtype=torch.int8, device=dev)
    y8 = torch.empty((1, B, HID), dtype=torch.int8, device=dev)
    lib = _cabi.lib()

    def quant():
        s = _cabi.stream()
        _cabi.check(lib.pa_batch_minmax_scale(x.data_ptr(), B, HID, xs.data_ptr(), s))
        _cabi.check(lib.pa_batch_quantize_i8(x.data_ptr(), xs.data_ptr(), B, HID, xq.data_ptr(), s))

    def mlp(i):
        assert ld.dnnl_matmul_int8(xq, W1[i], h8, 1, B, INTER, HID, 1 / 16, 1 / 16, 8.0, b1, "relu")
        assert ld.dnnl_matmul_int8(h8, W2[i], y8, 1, B, HID, INTER, 1 / 16, 1 / 16, 32.0, b2, "")

    res["act_quant_us"] = 1e3 * graph_time([quant], args.iters, dev, world)
    res["mlp_gemm_pair_us"] = 1e3 * graph_time([lambda i=i: mlp(i) for i in range(nw)], args.iters, dev, world)
    if rank == 0:
        kv_bytes = B * H * T * D * 2 + B * H * T * 2 * 4 + 2 * B * H * D * 4 + B * H * nt * 4
        ops = 2 * 2.0 * B * HID * INTER
        layer_us = sum(res.values())
        rows_total = B * world if args.weak else args.rows
        line = {"workload": f"C4: int8 KV decode + int8 MLP GEMMs, Llama-7B shape, {rows_total} rows over {world} GPU(s), ctx {T}",
                "scaling": "weak" if args.weak else "strong", "n_gpus": world, "rows_per_gpu": B,
                "timing": "CUDA-graph replay, events, max over ranks; rotating page / weight sets (no L2 reuse)",
                "us_per_layer": {k: round(v, 2) for k, v in res.items()}, "layer_us": round(layer_us, 2),
                "decode_alg_bytes_per_gpu": kv_bytes,
                "decode_gbs_per_gpu": round(kv_bytes / res["decode_i8_us"] / 1e3, 1),
                "gemm_pair_tops_per_gpu": round(ops / res["mlp_gemm_pair_us"] / 1e6, 1),
                "gemm_frac_of_4.5POPS": round(ops / res["mlp_gemm_pair_us"] / 1e6 / 4500, 3),
                "tok_s_32_layers_whole_job": round(rows_total / (32 * layer_us * 1e-6), 1)}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
