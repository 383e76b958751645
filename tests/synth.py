"""Seeded synthetic inputs for the paged-decode hot path.

Generated on the CPU so the oracle and the kernels see identical bits: q, K, V ~ N(0,1)
with K/V rounded to fp16 (or quantised with the int8_quant oracle), page ids a random
permutation of the pool so every gather is genuinely indirect.
"""
import numpy as np

import oracle


def make_case(B, H, D, T, tile_size=16, seed=0, kv="f16", unmapped_frac=0.0, ragged=False,
              beam_width=1, shared_prefix=0, spare_pages=3, temperature=None):
    """Returns a dict of numpy arrays.

    beam_width > 1: B rows = B/beam_width groups; rows of a group share the page ids of their
    first `shared_prefix` tokens (copy-on-write prefix pages), table rows are permuted and
    reached through beam_ids.
    """
    rng = np.random.default_rng(1234 + seed)
    num_tiles = (T + tile_size - 1) // tile_size
    num_beams = B
    n_entries = num_beams * H * num_tiles
    prefix_tiles = shared_prefix // tile_size if beam_width > 1 else 0
    groups = B // beam_width
    unique = groups * H * prefix_tiles + B * H * (num_tiles - prefix_tiles)
    P = unique + spare_pages
    perm = rng.permutation(P).astype(np.int32)
    table = np.full((num_beams, H, num_tiles), -1, dtype=np.int32)
    beam_ids = None
    row_of_beam = np.arange(B)
    if beam_width > 1:
        beam_ids = rng.permutation(B).astype(np.int32)  # row b reads table row beam_ids[b]
        row_of_beam = beam_ids
    nxt = 0
    for gidx in range(groups):
        rows = [row_of_beam[gidx * beam_width + w] for w in range(beam_width)]
        for h in range(H):
            shared = perm[nxt:nxt + prefix_tiles]
            nxt += prefix_tiles
            for r in rows:
                table[r, h, :prefix_tiles] = shared
                table[r, h, prefix_tiles:] = perm[nxt:nxt + num_tiles - prefix_tiles]
                nxt += num_tiles - prefix_tiles
    assert nxt == unique
    if unmapped_frac > 0:
        mask = rng.random(table.shape) < unmapped_frac
        table[mask] = -1
        table[0, 0, :] = -1          # one row with no keys at all
        if n_entries > 3:
            table.reshape(-1)[3] = P + 7  # out-of-range page id == unmapped (kv_tile_cache.hpp:23)
    q = rng.standard_normal((B, H, D)).astype(np.float32)
    kf = rng.standard_normal((P, tile_size, D)).astype(np.float32)
    vf = rng.standard_normal((P, tile_size, D)).astype(np.float32)
    case = dict(B=B, H=H, D=D, T=T, tile_size=tile_size, num_tiles=num_tiles, num_beams=num_beams,
                total_pages=P, table=table, q=q, beam_ids=beam_ids, kv=kv,
                temperature=float(np.sqrt(D)) if temperature is None else float(temperature))
    if ragged:
        ctx = rng.integers(0, T + 1, size=B).astype(np.int32)
        ctx[0] = T
        if B > 1:
            ctx[1] = 0
        if B > 2:
            ctx[2] = max(1, T - tile_size - 3)
        case["ctx_lens"] = ctx
    else:
        case["ctx_lens"] = None
    if kv == "f16":
        case["k_pool"] = kf.astype(np.float16)
        case["v_pool"] = vf.astype(np.float16)
    elif kv == "f32":  # KVTileCache<float>: the values as they are
        case["k_pool"], case["v_pool"] = kf, vf
    else:
        c = oracle.cpu
        ks = c.batch_minmax_scale(kf, D)
        vs = c.batch_minmax_scale(vf, D)
        case["k_pool"] = c.batch_quantize(kf, ks, D).reshape(P, tile_size, D)
        case["v_pool"] = c.batch_quantize(vf, vs, D).reshape(P, tile_size, D)
        case["k_scales"] = ks.reshape(P, tile_size)
        case["v_scales"] = vs.reshape(P, tile_size)
    return case


def oracle_attention(case, **kw):
    c = oracle.cpu
    args = dict(num_beams=case["num_beams"], num_tiles=case["num_tiles"], tile_size=case["tile_size"],
                T=case["T"], ctx_lens=case["ctx_lens"], beam_ids=case["beam_ids"],
                temperature=case["temperature"])
    args.update(kw)
    if case["kv"] in ("f16", "f32"):
        return c.paged_attention(case["q"], case["k_pool"].astype(np.float32),
                                 case["v_pool"].astype(np.float32), case["table"], **args)
    return c.paged_attention(case["q"], case["k_pool"], case["v_pool"], case["table"],
                             k_scales=case["k_scales"], v_scales=case["v_scales"], **args)


def to_device_cache(case, device="cuda"):
    """Build a llm_decoder.KVTileCache holding the case's pools and table."""
    import torch
    from llm_decoder import KVTileCache
    kvc = KVTileCache(case["kv"], device=device)
    k = torch.from_numpy(case["k_pool"]).to(device)
    v = torch.from_numpy(case["v_pool"]).to(device)
    if case["kv"] == "i8":
        kvc.adopt_buffers(k, v, torch.from_numpy(case["k_scales"]).to(device),
                          torch.from_numpy(case["v_scales"]).to(device))
    else:
        kvc.adopt_buffers(k, v)
    kvc.configure_table(case["num_beams"], case["H"], case["num_tiles"])
    kvc.page_table_.load_host_table(case["table"])
    return kvc
