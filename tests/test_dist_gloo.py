"""Multi-GPU host logic on CPU: world_size-2 gloo process group.  The device kernels
are replaced by the oracle here only as the CHECKER of the exchange plumbing (partials computed by
the oracle over each rank's page range, packed, all-gathered, combined)."""
import os
import subprocess
import sys
import textwrap

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_range_properties():
    sys.path.insert(0, os.path.join(ROOT, "pagedattention-based-transformer-decoder-inference-framework_b200"))
    from llm_decoder.dist import page_range, shard_range
    for n, w, g in [(64, 8, 1), (128, 8, 4), (10, 4, 1), (12, 8, 4), (256, 3, 1), (0, 2, 1)]:
        spans = [shard_range(n, w, r, g) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        for (a0, a1), (b0, b1) in zip(spans, spans[1:]):
            assert a1 == b0 and a0 <= a1
        assert all((b - a) % g == 0 for a, b in spans)           # beam groups never split
        sizes = [b - a for a, b in spans]
        assert max(sizes) - min(sizes) <= g
    with pytest.raises(ValueError):
        shard_range(10, 2, 0, group=4)
    assert [page_range(8192, 8, r) for r in (0, 7)] == [(0, 1024), (7168, 8192)]  # C5: 1024 pages per GPU


WORKER = textwrap.dedent("""
    import os, sys
    import numpy as np, torch, torch.distributed as dist
    ROOT = sys.argv[1]
    for p in (ROOT, os.path.join(ROOT, "pagedattention-based-transformer-decoder-inference-framework_b200"),
              os.path.join(ROOT, "tests")):
        sys.path.insert(0, p)
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:" + sys.argv[2], rank=int(sys.argv[3]), world_size=2)
    import oracle
    from llm_decoder import dist as pd
    from synth import make_case, oracle_attention
    rank, world = dist.get_rank(), dist.get_world_size()
    case = make_case(B=2, H=3, D=64, T=256, seed=9)
    nt = case["num_tiles"]
    t0, t1 = pd.page_range(nt, world, rank)
    # this rank's pages only: oracle attention over them, un-normalised via return_logits
    tb = np.full_like(case["table"], -1); tb[:, :, t0:t1] = case["table"][:, :, t0:t1]
    sub = dict(case); sub["table"] = tb
    out_loc, probs, logits = oracle_attention(sub, return_probs=True, return_logits=True)
    B, H, D = case["q"].shape
    pm = np.full((B, H), -np.inf, np.float32); pl = np.zeros((B, H), np.float32); po = np.zeros((B, H, D), np.float32)
    for b in range(B):
        for h in range(H):
            s = logits[b, h, t0 * 16:t1 * 16]
            pm[b, h] = s.max(); pl[b, h] = np.exp(s - s.max()).sum()
            po[b, h] = out_loc[b, h] * (pl[b, h] + 1e-6)
    msg = pd.pack_partials(torch.from_numpy(pm), torch.from_numpy(pl), torch.from_numpy(po).reshape(B * H, D))
    gathered = pd.gather_partials(msg)
    assert gathered.shape == (world, B * H, D + 2)
    comb = lambda m, l, o: torch.from_numpy(oracle.cpu.lse_combine(m.numpy(), l.numpy(), o.numpy()))
    out = pd.combine_gathered(gathered, combine_fn=comb).reshape(B, H, D).numpy()
    np.testing.assert_allclose(out, oracle_attention(case), rtol=1e-4, atol=1e-5)
    # product path refuses to combine on the CPU (no fallback)
    try:
        pd.combine_gathered(gathered)
        raise SystemExit("expected RuntimeError")
    except RuntimeError:
        pass
    # batch sharding: disjoint cover, every rank's rows equal the same rows of the full result
    r0, r1 = pd.shard_range(B, world, rank)
    rows = torch.zeros(B); rows[r0:r1] = 1
    dist.all_reduce(rows)
    assert torch.equal(rows, torch.ones(B))
    dist.barrier()
    dist.destroy_process_group()
    print("RANK_OK", rank)
""")


def test_split_kv_exchange_world2_gloo(tmp_path, oracle):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    port = str(29600 + os.getpid() % 300)
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, port, str(r)], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=240)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0 and f"RANK_OK {r}" in o, o[-3000:]
