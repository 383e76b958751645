"""Page lifecycle: the host logic either side of gather/append.

kv_cache/kv_tile_cache.cpp:27-37 (resize), :65-98 (register_tile, LRU, evict_if_needed), :105-125
(save_to_file / load_from_file: raw K pool then raw V pool, no header) -- with the decisions of SURVEY App. A
D14 (free-list page ids, eviction clears the device entry) and the copy-on-write refcounts of the beam path:
a page shared by several beams is returned to the free list only when its LAST reference is released.
"""
import os

import numpy as np
import pytest
import torch

from synth import make_case, oracle_attention, to_device_cache

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ld(oracle):
    import llm_decoder
    return llm_decoder


def _cache(ld, pages, beams=4, heads=2, tiles=4, kv="f16", tile=16, D=64):
    kvc = ld.KVTileCache(kv)
    kvc.init(pages, tile, D)
    kvc.configure_table(beams, heads, tiles)
    return kvc


def _device_table(kvc):
    return kvc.page_table_.device_data().cpu().numpy().reshape(kvc.page_table_.num_beams_, kvc.page_table_.num_heads_,
                                                               kvc.page_table_.num_tiles_)


def test_register_tile_lru_order_and_eviction(ld):
    """Pool of 4 pages: the 5th registration evicts the least recently USED tile (a re-registration of a live
    tile refreshes it, update_lru kv_tile_cache.cpp:93-98), the evicted tile's device entry becomes -1 and
    its page id is the one re-issued (free list, never a live id: App. A D14)."""
    kvc = _cache(ld, pages=4)
    p = [kvc.register_tile(0, 0, t) for t in range(4)]
    assert sorted(p) == [0, 1, 2, 3] and len(set(p)) == 4
    assert kvc.register_tile(0, 0, 0) == p[0]           # already mapped: same page, now most recently used
    p4 = kvc.register_tile(1, 0, 0)                      # evicts (0,0,1), the oldest untouched tile
    assert p4 == p[1]
    tb = _device_table(kvc)
    assert tb[0, 0, 1] == -1 and tb[1, 0, 0] == p[1] and tb[0, 0, 0] == p[0]
    assert (0, 0, 1) not in kvc.tile_to_page_map_ and kvc.page_table_.lookup(0, 0, 1) == -1
    p5 = kvc.register_tile(1, 0, 1)                      # next victim: (0,0,2)
    assert p5 == p[2] and _device_table(kvc)[0, 0, 2] == -1
    live = [pg for pg in _device_table(kvc).reshape(-1) if pg >= 0]
    assert len(live) == len(set(live)) == 4              # never two tiles on one page


def test_eviction_respects_shared_pages(ld):
    """fork_beam shares pages (refcount 2).  Evicting ONE of the two tiles that map a shared page must not free
    it: the page returns to the free list only when the second reference goes too."""
    kvc = _cache(ld, pages=3, beams=3, heads=1, tiles=3)
    a = kvc.register_tile(0, 0, 0)
    b = kvc.register_tile(0, 0, 1)
    kvc.fork_beam(0, 1, num_tiles=2)                     # beam 1 shares pages a, b
    assert kvc.page_refcount(a) == 2 and kvc.page_refcount(b) == 2
    c = kvc.register_tile(2, 0, 0)                       # the third and last free page
    assert sorted([a, b, c]) == [0, 1, 2]
    # pool full.  LRU order: (0,0,0), (0,0,1), (1,0,0), (1,0,1), (2,0,0).  A new tile needs a page: evicting
    # (0,0,0) only drops a reference (page `a` still mapped by beam 1), so the walk continues until a page is free.
    d = kvc.register_tile(2, 0, 1)
    tb = _device_table(kvc)
    live = {(bm, t): tb[bm, 0, t] for bm in range(3) for t in range(3) if tb[bm, 0, t] >= 0}
    assert d in (a, b, c)
    owners = [k for k, pg in live.items() if pg == d]
    assert owners == [(2, 1)], (live, d)                 # the re-issued page has exactly one owner
    for pg in set(live.values()):                        # refcounts equal the number of table entries
        assert kvc.page_refcount(pg) == sum(1 for v in live.values() if v == pg)
    # while beam 1 still maps a shared page, that page is not in the free list
    for (bm, t), pg in live.items():
        assert pg not in kvc._free


def test_pool_exhausted_is_an_explicit_error(ld):
    kvc = _cache(ld, pages=2, beams=2, heads=1, tiles=4)
    kvc.register_tile(0, 0, 0)
    kvc.register_tile(0, 0, 1)
    kvc.tile_to_page_map_.clear()                        # nothing left to evict, nothing free
    with pytest.raises(RuntimeError, match="page pool exhausted"):
        kvc.register_tile(1, 0, 0)


def test_resize_clears_table_free_list_and_refcounts(ld):
    """kv_tile_cache.cpp:27-37: free, reallocate with the new geometry, clear the map and the page table."""
    kvc = _cache(ld, pages=6, beams=2, heads=2, tiles=3)
    for t in range(3):
        kvc.register_tile(0, 1, t)
    kvc.fork_beam(0, 1)
    assert any(kvc.page_refcount(p) == 2 for p in range(6))
    kvc.resize(10, 32)
    assert kvc.total_pages_ == 10 and kvc.tile_size_ == 32
    assert tuple(kvc.key_buffer_.shape) == (10, 32, 64) == tuple(kvc.value_buffer_.shape)
    assert (_device_table(kvc) == -1).all() and (kvc.page_table_.host_table_ == -1).all()
    assert sorted(kvc._free) == list(range(10)) and not kvc.tile_to_page_map_
    assert all(kvc.page_refcount(p) == 1 for p in range(10))       # stale shared counts are gone
    got = sorted(kvc.register_tile(1, 0, t) for t in range(3))
    assert got == [0, 1, 2]                                         # fresh pool hands out fresh ids, no COW state
    # init() on a live cache resets the same state (and clears a previously configured table)
    kvc.init(4, 16, 64)
    assert sorted(kvc._free) == [0, 1, 2, 3] and (_device_table(kvc) == -1).all()


@pytest.mark.parametrize("kv", ["f16", "i8"])
def test_save_load_file_layout_and_attention_round_trip(ld, oracle, kv, tmp_path):
    """save_to_file writes the raw K pool then the raw V pool, nothing else in front (kv_tile_cache.cpp:105-113);
    load_from_file into a second cache of the same geometry attends bit-identically."""
    case = make_case(B=3, H=2, D=128, T=112, seed=92, kv=kv, unmapped_frac=0.05)
    kvc = to_device_cache(case)
    path = str(tmp_path / "pools.bin")
    kvc.save_to_file(path)
    raw = open(path, "rb").read()
    kb, vb = case["k_pool"].tobytes(), case["v_pool"].tobytes()
    assert raw[:len(kb)] == kb and raw[len(kb):len(kb) + len(vb)] == vb
    if kv == "f16":
        assert len(raw) == len(kb) + len(vb)             # exactly the reference's payload
    else:                                                # int8 extension: the two scale arrays follow
        assert raw[len(kb) + len(vb):] == case["k_scales"].tobytes() + case["v_scales"].tobytes()
    kvc2 = ld.KVTileCache(kv)
    kvc2.init(case["total_pages"], 16, 128)
    kvc2.configure_table(case["num_beams"], case["H"], case["num_tiles"])
    kvc2.page_table_.load_host_table(case["table"])
    kvc2.load_from_file(path)
    assert torch.equal(kvc2.key_buffer_, kvc.key_buffer_) and torch.equal(kvc2.value_buffer_, kvc.value_buffer_)
    B, H, D = case["q"].shape
    q = torch.from_numpy(case["q"]).cuda()
    outs = []
    for c in (kvc, kvc2):
        out = torch.empty((B, H, D), device="cuda")
        ld.AttentionCUDA.forward(q, out, B, H, D, case["T"], None, c, None, False, kv == "f16", True, case["temperature"])
        outs.append(out.cpu().numpy())
    np.testing.assert_array_equal(outs[0], outs[1])
    np.testing.assert_allclose(outs[0], oracle_attention(case), rtol=2e-3, atol=1e-3)
    # a short file is an error, as in the reference ("Failed to read ..." kv_tile_cache.cpp:120-123)
    with open(path, "wb") as f:
        f.write(raw[:len(kb) // 2])
    with pytest.raises(RuntimeError):
        kvc2.load_from_file(path)


def test_cpu_format_load_never_aliases_tiles(ld, tmp_path):
    """load_tiles_cpu_format into a pool that cannot hold the file's tiles must fail instead of evicting tiles
    registered earlier in the same load (two tiles on one page)."""
    case = make_case(B=2, H=2, D=64, T=64, seed=93)
    kvc = to_device_cache(case)
    path = str(tmp_path / "tiles.bin")
    n = kvc.save_tiles_cpu_format(path)
    small = ld.KVTileCache("f16")
    small.init(n - 3, 16, 64)
    small.configure_table(2, 2, case["num_tiles"])
    with pytest.raises(RuntimeError, match="do not fit"):
        small.load_tiles_cpu_format(path)
    assert not small.tile_to_page_map_                    # nothing half-loaded
    ok = ld.KVTileCache("f16")
    ok.init(n, 16, 64)
    ok.configure_table(2, 2, case["num_tiles"])
    assert ok.load_tiles_cpu_format(path) == n
    pages = [p for p in _device_table(ok).reshape(-1) if p >= 0]
    assert len(pages) == len(set(pages)) == n
