#!/usr/bin/env python
"""In-kernel timeline of the tcgen05 prefill kernel (csrc/prefill_tc.cu built with -DPA_PTC_PROBE).

    python benchmarks/prefill_probe.py --build      # here (no GPU): nvcc -> libpa_b200_probe.so, in-tree
    gpurun -- python benchmarks/prefill_probe.py    # on the B200: run B = 2, Tq = 8192 and print the medians

clock64 stamps of warp 0 (softmax group A), warp 4 (group B) and the UMMA warp over 64 steady-state KV tiles
of CTA 5.  Softmax segments: wait S | tcgen05.ld | mask + max + exp2 + pack | wait P.V(i-1) | rescale + P
store to TMEM | fence + arrive.  UMMA issuer of tile A (default two-issuer kernel): wait K/V(i+1) + issue S(i+1) |
wait P(i) | issue P.V(i).  With PA_PREFILL_MW=1 (single issuer) the stamps are: loop top | S issued | A's P seen |
A's P.V issued | B's P seen | B's P.V issued.  The product library is not touched (the probes compile to nothing
there: identical SASS)."""
import ctypes
import glob
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "pagedattention-based-transformer-decoder-inference-framework_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)
PROBE_SO = os.path.join(PKG, "libpa_b200_probe.so")


def build():
    import importlib.util
    spec = importlib.util.spec_from_file_location("pa_b200_build", os.path.join(PKG, "build.py"))
    b = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(b)
    out_dir = os.path.join(PKG, "build", "probe")
    os.makedirs(out_dir, exist_ok=True)
    objs = []
    for src in b.sources():
        obj = os.path.join(out_dir, os.path.basename(src) + ".o")
        subprocess.check_call([b.NVCC, *b.FLAGS, "-DPA_PTC_PROBE", "-c", src, "-o", obj])
        objs.append(obj)
    subprocess.check_call([b.NVCC, "-shared", "-o", PROBE_SO, *objs, "-cudart", "static", "-ccbin", "/usr/bin/g++"])
    print(PROBE_SO)


def main():
    import torch
    from llm_decoder import _cabi
    _cabi.LIB_PATH = PROBE_SO
    import llm_decoder as ld
    dev = torch.device("cuda", 0)
    H, D, TILE, B, Tq = 32, 128, 16, 2, 8192
    nt = Tq // TILE
    P = B * H * nt
    g = torch.Generator(device=dev).manual_seed(3)
    kvc = ld.KVTileCache("f16", device=dev)
    kvc.adopt_buffers(torch.randn((P, TILE, D), generator=g, device=dev, dtype=torch.float16),
                      torch.randn((P, TILE, D), generator=g, device=dev, dtype=torch.float16))
    kvc.configure_table(B, H, nt)
    kvc.page_table_.load_host_table(torch.randperm(P, generator=g, device=dev).to(torch.int32).cpu().numpy().reshape(B, H, nt))
    q = torch.randn((B, H, Tq, D), generator=g, device=dev)
    out = torch.empty_like(q)
    for _ in range(3):
        ld.paged_prefill(q, out, kvc, B, Tq, float(np.sqrt(D)))
    torch.cuda.synchronize()
    buf = (ctypes.c_longlong * (3 * 64 * 8))()
    fn = _cabi.lib().pa_debug_ptc_probe
    fn.argtypes = [ctypes.POINTER(ctypes.c_longlong)]
    assert fn(buf) == 0
    a = np.array(buf[:]).reshape(3, 64, 8)
    cta_life(_cabi)
    if "--raw" in sys.argv:  # stamps of 8 consecutive tiles relative to the first
        t0 = a[:, 20, :].min()
        for role, name in ((0, "softmax A"), (1, "softmax B"), (2, "umma")):
            for it in range(20, 28):
                print(name, "tile", it + 16, (a[role, it, :7] - t0).tolist())
    for role, name, nseg in ((0, "softmax A", 6), (1, "softmax B", 6), (2, "umma", 3)):
        r = a[role]
        per = np.diff(r[:, 0])
        seg = np.median(np.diff(r[:, :nseg + 1], axis=1), axis=0)
        print(f"{name}: period median {np.median(per):.0f} cycles (min {per.min()}, max {per.max()}); "
              f"segment medians {[int(v) for v in seg]}")


def cta_life(_cabi):
    buf = (ctypes.c_longlong * 8)()
    fn = _cabi.lib().pa_debug_ptc_probe_cta
    fn.argtypes = [ctypes.POINTER(ctypes.c_longlong)]
    assert fn(buf) == 0
    v = np.array(buf[:])
    names = ["setup (barriers, TMEM alloc, sync)", "Q staged", "first S seen", "KV loop", "last P.V seen", "O stored", "exit sync"]
    print("CTA 5 life (cycles):", {n: int(d) for n, d in zip(names, np.diff(v))}, "total", int(v[7] - v[0]))


if __name__ == "__main__":
    build() if "--build" in sys.argv else main()
