#!/usr/bin/env python
"""C4 through the drop-in surface: INT8Decoder (Llama-7B shape: 32 layers, 32 heads, D = 128, hidden 4096,
MLP 4x, int8 weights + int8 KV pages) decoding B sequences at ~4K context, greedy, CUDA-graph steps.
The KV pools are zero-initialised pages at positions ~4080 (the bytes are streamed all the same); weights
are random int8.  Reports decode tokens/s of the whole loop (embedding, LayerNorm, append+quantise, paged
int8 attention, dynamic quantise, two tcgen05 GEMMs, logits+argmax per step)."""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "pagedattention-based-transformer-decoder-inference-framework_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=96)
    ap.add_argument("--layers", type=int, default=32)
    ap.add_argument("--ctx", type=int, default=4096)
    ap.add_argument("--steps", type=int, default=12)
    ap.add_argument("--eager", action="store_true", help="no CUDA graph (for ncu launch lists)")
    ap.add_argument("--decoder", default="int8", choices=["int8", "cuda"],
                    help="int8: INT8Decoder (C4); cuda: CUDADecoder, fp32 weights + fp16 KV pages (C2 shape)")
    args = ap.parse_args()
    import llm_decoder as ld
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    L, H, D, HID, V = args.layers, 32, 128, 4096, 32000
    g = torch.Generator(device=dev).manual_seed(5)
    if args.decoder == "int8":
        dec = ld.INT8Decoder(L, H, D, HID, V, args.ctx, batch_size=args.batch, use_cuda_graph=not args.eager)
        dec.embedding.copy_(torch.randint(-127, 128, dec.embedding.shape, generator=g, device=dev, dtype=torch.int8))
        for Ly in dec.layers:
            Ly.fc1_w.copy_(torch.randint(-127, 128, Ly.fc1_w.shape, generator=g, device=dev, dtype=torch.int8))
            Ly.fc2_w.copy_(torch.randint(-127, 128, Ly.fc2_w.shape, generator=g, device=dev, dtype=torch.int8))
            Ly.fc1_deq = Ly.fc2_deq = 0.02 / 127
    else:
        dec = ld.CUDADecoder(L, H, D, HID, V, args.ctx, batch_size=args.batch, use_cuda_graph=not args.eager)
        dec.embedding.normal_(generator=g)
        for Ly in dec.layers:
            for w in (Ly.fc1_w, Ly.fc2_w):
                w.normal_(generator=g)
                w.mul_(0.02)
    B = args.batch
    start = args.ctx - args.steps - 4
    dec._temperature = 1.0
    dec._sampling = None
    dec.ids.copy_(torch.randint(0, V, (B,), generator=g, device=dev, dtype=torch.int32))
    dec.positions.fill_(start)
    dec.ctx_lens.fill_(start + 1)
    for _ in range(3):          # eager warm-up, then graph capture inside _step_or_replay
        dec._step_or_replay()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        dec._step_or_replay()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / args.steps
    tok = start + args.steps // 2
    kv_bytes = B * H * tok * D * 2 + B * H * tok * 8 if args.decoder == "int8" else B * H * tok * D * 2 * 2
    name = "C4 via INT8Decoder" if args.decoder == "int8" else "C2 shape via CUDADecoder (fp32 weights, 3xTF32 tcgen05 MLP)"
    print(json.dumps({"workload": f"{name}: {L} layers, batch {B}, ctx ~{args.ctx}",
                      "ms_per_step": round(dt * 1e3, 3), "decode_tok_s": round(B / dt, 1),
                      "attention_kv_gbs_lower_bound": round(L * kv_bytes / dt / 1e9, 1),
                      "kv_pool_gib": round(L * 2 * B * H * (args.ctx // 16) * 16 * D * (1 if args.decoder == "int8" else 2) / 2**30, 1)}))


if __name__ == "__main__":
    main()
