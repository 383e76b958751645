#!/usr/bin/env python
"""bench.py -- headline benchmark of the paged-decode hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (config C2 of BASELINE.json, the one the metric is quoted on): CUDADecoder-style fp16
paged decode, Llama-7B shape (32 layers, 32 heads, head_dim 128), batch 64 rows PER GPU, 4096
tokens of context, 16-token pages, synthetic random-init data, page ids a random permutation
of the pool.  One STEP = one decode step of the whole batch = for each of the 32 layers: append
the new token's K/V rows to their pages (pa_kv_append_f32_f16) and run paged attention over the
4096-token context (pa_paged_decode_f16_overlap).  Each layer has its own 4 GiB K+V pool
(128 GiB of KV per GPU; inputs of one launch are 4 GiB >> 126 MB L2, so no flush is needed).

value  = tokens/s of the whole job (rows of all ranks / step time), inputs resident in HBM.
e2e    = the same step through the reference-facing call AttentionCUDA.forward with HOST q /
         new-K / new-V / out buffers (pinned H2D + D2H inside the timed region, every layer).
roofline = achieved algorithmic GB/s of the decode launch vs the measured HBM copy peak.
cpu_baseline / --impl reference = the reference's CPU attention path (oracle port of
         cpu_paged_attention_forward, which does not compile as shipped) on the host cores,
         on a bounded sample (a few rows of one layer), scaled to the same metric.
extra  = the other BASELINE.json configurations under the same clock (benchmarks/extras.py), each with its
         achieved rate, fraction of the measured peak and a max-abs-err against the CPU oracle on a bounded
         sample: c4_int8_decode, c4_gemm_pair (+ an int8 library peak measured in the same run), c3_group, c2_fp32_mlp,
         c5_splitkv (one 128K-token sequence split over the N ranks: decode + merge + exchange in ONE launch, the
         stand-alone peer-memory exchange, and the NCCL all-gather form).  --no-extra skips them.
N > 1: one process per GPU (torchrun), rows sharded across ranks, no data-path collective in the headline
         step; c5_splitkv is the mode with an exchange.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "pagedattention-based-transformer-decoder-inference-framework_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

METRIC = "decode tok/s (paged attention + KV append, Llama-7B shape 32L/32H/d128, batch 64/GPU, 4K ctx, 16-token pages)"
UNIT = "tok/s"
LAYERS, HEADS, HDIM, CTX, TILE = 32, 32, 128, 4096, 16


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=64, help="rows per GPU")
    ap.add_argument("--layers", type=int, default=LAYERS)
    ap.add_argument("--ctx", type=int, default=CTX)
    ap.add_argument("--cpu-rows", type=int, default=16, help="rows in the CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--kernel", default="overlap", choices=["overlap", "fused"])
    ap.add_argument("--no-extra", action="store_true", help="skip the C3 / C4 / C5 extra block")
    return ap.parse_args()


def workload_config(args):
    """The workload both arms measure (identical dict in both JSON lines)."""
    return {"workload": "C2: fp16 paged decode, Llama-7B shape, batch 64/GPU, 4K ctx, 16-token pages",
            "layers": args.layers, "heads": HEADS, "head_dim": HDIM, "ctx": args.ctx, "page_tokens": TILE,
            "rows_per_gpu": args.batch, "parallelism": f"batch-sharded x{args.gpus}, no collective"}


def host_threads():
    """Cores this process may run on (the CPU arm uses all of them, whatever OMP_NUM_THREADS torchrun exported)."""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


# --------------------------------------------------------------------------- CPU baseline
def cpu_reference_sample(rows, ctx, repeats):
    """Time the oracle port of cpu_paged_attention_forward<float> (cpu_attention_kernel.cpp:36-129,
    OpenMP over (b,h)) on `rows` rows x 32 heads x ctx tokens of ONE layer.  Returns
    (seconds per layer-pass [list], threads)."""
    import numpy as np

    os.environ["OMP_NUM_THREADS"] = str(host_threads())  # torchrun exports OMP_NUM_THREADS=1
    import oracle
    oracle.cpu.build()
    c = oracle.cpu
    c.set_threads(host_threads())
    rng = np.random.default_rng(1236)
    nt = ctx // TILE
    P = rows * HEADS * nt
    k = rng.standard_normal((P, TILE, HDIM), dtype=np.float32).astype(np.float16).astype(np.float32)
    v = rng.standard_normal((P, TILE, HDIM), dtype=np.float32).astype(np.float16).astype(np.float32)
    q = rng.standard_normal((rows, HEADS, HDIM), dtype=np.float32)
    table = rng.permutation(P).astype(np.int32).reshape(rows, HEADS, nt)
    times = []
    for _ in range(repeats):
        t0 = time.perf_counter()
        c.paged_attention(q, k, v, table, num_beams=rows, num_tiles=nt, tile_size=TILE, T=ctx,
                          temperature=float(np.sqrt(HDIM)))
        times.append(time.perf_counter() - t0)
    return times, c.num_threads()


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    rows = args.cpu_rows
    times, threads = cpu_reference_sample(rows, args.ctx, args.warmup + args.steps)
    timed = times[args.warmup:] if len(times) > args.warmup else times
    t_layer = sum(timed) / len(timed)
    step_s = t_layer * args.layers
    value = rows / step_s
    sample = (f"{rows} rows x {HEADS} heads x {args.ctx} ctx of ONE layer per step (float K/V, as "
              f"CPUAttention<float>), scaled x{args.layers} layers; oracle port of cpu_paged_attention_forward")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": step_s * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": workload_config(args),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------- clocks
class ClockSampler:
    """Polls SM clock, power and throttle reasons through NVML every ~10 ms on a thread while the
    timed region runs (the profiling recipe's clocks line, without nvidia-smi's 100 ms floor)."""

    def __init__(self, index):
        self.samples = []
        self.ok = False
        self._stop = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_sm = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
            self.th = threading.Thread(target=self._poll, daemon=True)
            self.th.start()
        except Exception as e:  # pragma: no cover
            self.err = repr(e)

    def _poll(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                sm = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                rs = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                pw = nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0
                self.samples.append((time.perf_counter(), sm, rs, pw))
            except Exception:
                pass
            time.sleep(0.01)

    def stop(self, t0, t1):
        if not self.ok:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable: " + getattr(self, "err", "")]}
        self._stop.set()
        self.th.join(timeout=1.0)
        nv = self.nv
        flags = {"hw_slowdown": nv.nvmlClocksEventReasonHwSlowdown,
                 "hw_thermal_slowdown": nv.nvmlClocksEventReasonHwThermalSlowdown,
                 "sw_thermal_slowdown": nv.nvmlClocksEventReasonSwThermalSlowdown,
                 "sw_power_cap": nv.nvmlClocksEventReasonSwPowerCap}
        win = [s for s in self.samples if t0 <= s[0] <= t1]
        reasons = sorted({n for s in win for n, bit in flags.items() if s[2] & bit})
        return {"sm_mhz": statistics.median([s[1] for s in win]) if win else None, "sm_max_mhz": self.max_sm,
                "reasons": reasons, "samples": len(win), "power_w_max": max([s[3] for s in win]) if win else None}


# --------------------------------------------------------------------------- ours
def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    import llm_decoder as ld
    from llm_decoder import _cabi

    os.environ.setdefault("NCCL_DEBUG", "WARN")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _cabi.lib()  # raises if the CUDA library is missing: no fallback

    B, L, H, D, T = args.batch, args.layers, HEADS, HDIM, args.ctx
    nt = T // TILE
    P = B * H * nt
    pool_bytes = 2 * P * TILE * D * 2
    free, _ = torch.cuda.mem_get_info()
    n_pools = int(max(1, min(L, (free - (10 << 30)) // pool_bytes)))
    g = torch.Generator(device=dev).manual_seed(1236 + rank)
    caches = []
    table = torch.randperm(P, generator=g, device=dev).to(torch.int32).reshape(B, H, nt)
    host_table = table.cpu().numpy()
    for i in range(n_pools):
        k = torch.empty((P, TILE, D), dtype=torch.float16, device=dev)
        v = torch.empty((P, TILE, D), dtype=torch.float16, device=dev)
        k.normal_(generator=g)
        v.normal_(generator=g)
        kvc = ld.KVTileCache("f16", device=dev)
        kvc.adopt_buffers(k, v)
        kvc.configure_table(B, H, nt)
        kvc.page_table_.load_host_table(host_table)
        caches.append(kvc)
    q_dev = torch.randn((L, B, H, D), generator=g, device=dev)
    nk_dev = torch.randn((L, B, H, D), generator=g, device=dev)
    nv_dev = torch.randn((L, B, H, D), generator=g, device=dev)
    out_dev = torch.empty((L, B, H, D), device=dev)
    pos = torch.full((B,), T - 1, dtype=torch.int32, device=dev)
    temp = float(np.sqrt(D))
    use_overlap = args.kernel == "overlap"
    stream = torch.cuda.current_stream()

    def layer_device(l, ev=None):
        kvc = caches[l % n_pools]
        kvc.append(nk_dev[l], nv_dev[l], pos)
        if ev is not None:
            ev[0].record(stream)
        ld.AttentionCUDA.forward(q_dev[l], out_dev[l], B, H, D, T, None, kvc, None, False, True, use_overlap, temp)
        if ev is not None:
            ev[1].record(stream)

    def step_device(evs=None):
        for l in range(L):
            layer_device(l, None if evs is None else evs[l])

    def sync_all():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    # ---- device-resident timing -------------------------------------------------------
    for _ in range(max(args.warmup, 3)):
        step_device()
    evs = [[(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(L)]
           for _ in range(args.steps)]
    sync_all()
    sampler = ClockSampler(local) if rank == 0 else None
    t_wall0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for s in range(args.steps):
        step_device(evs[s])
    e1.record(stream)
    sync_all()
    t_wall1 = time.perf_counter()
    clocks = sampler.stop(t_wall0, t_wall1) if sampler else None
    ms_total = e0.elapsed_time(e1)
    kern_ms = [a.elapsed_time(b) for st in evs for (a, b) in st]
    if world > 1:
        t = torch.tensor([ms_total], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_step = ms_total / args.steps
    value = world * B / (ms_step * 1e-3)

    # ---- end to end: host buffers through AttentionCUDA.forward -------------------------------
    # host-side I/O buffers live in page-locked memory (the contract's "pinned host memory")
    q_host = torch.empty((L, B, H, D)).pin_memory().copy_(q_dev.cpu())
    nk_host = torch.empty((L, B, H, D)).pin_memory().copy_(nk_dev.cpu())
    nv_host = torch.empty((L, B, H, D)).pin_memory().copy_(nv_dev.cpu())
    out_host = torch.empty((L, B, H, D)).pin_memory()

    def step_e2e():
        # Host q / new K / new V in, host out back, every layer.  sync=False: the calls only enqueue; the copies
        # ride the two copy engines (llm_decoder.attention.HostPipe) under the neighbouring layers' kernels and
        # the step ends with ONE synchronisation, after which all 32 host outputs are valid.
        for l in range(L):
            kvc = caches[l % n_pools]
            kvc.append(nk_host[l], nv_host[l], pos)
            ld.AttentionCUDA.forward(q_host[l], out_host[l], B, H, D, T, None, kvc, None, False, True,
                                     use_overlap, temp, sync=False)
        ld.AttentionCUDA.synchronize(dev)

    e2e_steps = max(2, min(args.steps, 5))
    for _ in range(2):
        step_e2e()
    sync_all()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        step_e2e()
    torch.cuda.synchronize(dev)
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    if world > 1:
        t = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = world * B / e2e_s
    row_bytes = B * H * D * 4

    # The same host-buffer step when every layer NEEDS the previous layer's host result (the layers of a real decoder
    # are dependent; above they are not, so their copies hide under neighbouring kernels): one synchronisation per
    # layer, copies in line with the kernels.  Reported beside e2e, not instead of it.
    def step_e2e_dependent():
        for l in range(L):
            kvc = caches[l % n_pools]
            kvc.append(nk_host[l], nv_host[l], pos)
            ld.AttentionCUDA.forward(q_host[l], out_host[l], B, H, D, T, None, kvc, None, False, True, use_overlap, temp,
                                     sync=True)
    step_e2e_dependent()
    sync_all()
    t0 = time.perf_counter()
    for _ in range(2):
        step_e2e_dependent()
    torch.cuda.synchronize(dev)
    e2e_dep_s = (time.perf_counter() - t0) / 2
    if world > 1:
        t = torch.tensor([e2e_dep_s], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_dep_s = float(t.item())
    # parity guard inside the bench: e2e output of the last layer == device-resident output
    agree = float((out_host[L - 1] - out_dev[L - 1].cpu()).abs().max())

    # ---- roofline of the dominant kernel (paged decode) ------------------------------------------
    kv_bytes = B * H * T * D * 2 * 2
    alg_bytes = kv_bytes + 2 * row_bytes + B * H * nt * 4
    avg_ms = sum(kern_ms) / len(kern_ms)
    achieved = alg_bytes / (avg_ms * 1e-3) / 1e9
    peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peak = float(json.load(f)["hbm_gbs"])
            peak_src = "MEASURED_PEAKS.json hbm_gbs"
    except Exception:
        pass
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "decode_traffic.json")) as f:
            traffic = json.load(f).get("dram_bytes_per_launch")
    except Exception:
        pass

    # ---- parity of the timed configuration itself: a bounded sample of the last layer vs the CPU oracle -------
    parity = None
    if rank == 0:
        try:
            sys.path.insert(0, os.path.join(ROOT, "benchmarks"))
            import extras
            pairs = [(0, 0), (B // 2, H // 2), (B - 1, H - 1)]
            exp = extras.oracle_rows(caches[(L - 1) % n_pools], q_dev[L - 1], pairs, T, temp)
            got = np.stack([out_dev[L - 1][b, h].cpu().numpy() for b, h in pairs])
            parity = {"max_abs_err_vs_oracle": float(np.abs(got - exp).max()),
                      "ok": bool(np.allclose(got, exp, rtol=2e-3, atol=1e-3)), "tolerance": "2e-3 rel + 1e-3 abs",
                      "sample": "3 (row, head) pairs of the last layer at the full 4096-token context"}
        except Exception as e:  # the bench number stands; the failure is reported, not hidden
            parity = {"error": repr(e)}

    # ---- the other BASELINE.json configurations under the same clock ----------------------------------------
    extra = None
    if not args.no_extra:
        del caches, q_dev, nk_dev, nv_dev, out_dev, q_host, nk_host, nv_host, out_host
        torch.cuda.empty_cache()
        sys.path.insert(0, os.path.join(ROOT, "benchmarks"))
        import extras
        extra = {}

        def run_extra(name, fn, *a, **kw):
            try:
                extra[name] = fn(*a, **kw)
            except Exception as e:
                extra[name] = {"error": repr(e)}
                torch.cuda.empty_cache()
        if world == 1:
            run_extra("c4_int8_decode", extras.c4_int8_decode, dev, peak)
            run_extra("c4_gemm_pair", extras.c4_gemm_pair, dev)
            run_extra("c3_group", extras.c3_group, dev, peak)
            run_extra("c2_fp32_mlp", extras.c2_fp32_mlp, dev, peak)
        run_extra("c5_splitkv", extras.c5_splitkv, dev, world, rank, peak)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    cpu_base = None
    if world == 1 and not args.no_cpu_baseline:
        rows = args.cpu_rows
        times, threads = cpu_reference_sample(rows, T, 3)
        t_layer = min(times[1:]) if len(times) > 1 else times[0]
        cpu_base = {"value": rows / (t_layer * L), "unit": UNIT, "cores": threads, "kind": "port",
                    "sample": f"{rows} rows x {H} heads x {T} ctx of one layer (float K/V), best of {len(times) - 1}, "
                              f"scaled x{L} layers; oracle port of cpu_paged_attention_forward (reference TU does not compile)"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32 (fp16 K/V storage, fp32 accumulate)", "data": "synthetic",
        "config": workload_config(args),
        "run_info": {"kv_pools": n_pools, "kv_bytes_per_gpu": n_pools * pool_bytes, "kernel": args.kernel,
                     "l2": "each launch streams 4 GiB >> 126 MB L2; no flush needed",
                     "decode_launches_per_layer": 1 if os.environ.get("PA_DECODE_MERGE_KERNEL") == "0" else 2},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "peak_source": peak_src, "kernel": "paged_decode_%s_kernel<128,f16>" % args.kernel,
                     "alg_bytes_per_launch": alg_bytes, "avg_launch_ms": avg_ms, "min_launch_ms": min(kern_ms),
                     "frac_of_8TBps": achieved / 8000.0},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 3 * row_bytes * L,
                "d2h_bytes_per_step": row_bytes * L, "ms_per_step": e2e_s * 1e3,
                "max_abs_diff_vs_device_path": agree,
                "dependent_layers_value": world * B / e2e_dep_s, "dependent_layers_ms_per_step": e2e_dep_s * 1e3,
                "note": "value: the 32 layer calls of a step are independent, copies overlap neighbouring kernels; "
                        "dependent_layers_value: one host synchronisation per layer (each layer waits for the previous "
                        "layer's host result)"},
        "gpu_launches": args.steps * L * (2 if os.environ.get("PA_DECODE_MERGE_KERNEL") == "0" else 3),
        "clocks": clocks,
        "parity": parity,
    }
    if extra is not None:
        line["extra"] = extra
    if cpu_base:
        line["cpu_baseline"] = cpu_base
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
