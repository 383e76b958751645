"""KVTileCache -- GPU page pool + page table + LRU registration.

Mirrors kv_cache/kv_tile_cache.hpp:9-80 / kv_tile_cache.cpp:7-128: two device pools
(`key_buffer_`, `value_buffer_`) of `total_pages * tile_size * head_dim` elements, page p
of either pool starting at element p*tile_size*head_dim (cpp:52-62), K and V of a tile
sharing one page id.  Storage dtype: 'f16' (fp16 pages, KVTileCache<half>), 'f32' (fp32 pages,
KVTileCache<float> -- the two instantiations of kv_tile_cache.cpp:127-128) or 'i8' (int8 pages plus
one f32 scale per (page, row) for K and for V).

Decisions taken where the reference contradicts itself: page ids come
from a free list (the reference's `map.size()` re-issues live ids after an eviction),
eviction clears the device entry, and the page-table dimensions are given explicitly with
`configure_table(num_beams, num_heads, num_tiles)` (the reference passes
(total_pages, head_dim, tile_size) into PageTable::init by mistake, cpp:23).  64-bit page
offsets (D17).
"""
import threading
from collections import OrderedDict

import numpy as np
import torch

from . import _cabi
from .page_table import PageTable

_DTYPES = {"f16": torch.float16, "i8": torch.int8, "f32": torch.float32}


class KVTileCache:
    def __init__(self, dtype="f16", device=None):
        if dtype not in _DTYPES:
            raise ValueError("KVTileCache dtype must be 'f16', 'i8' or 'f32'")
        self.dtype = dtype
        self._device = torch.device(device) if device is not None else None
        self.key_buffer_ = self.value_buffer_ = None
        self.k_scales_ = self.v_scales_ = None
        self.tile_size_ = self.head_dim_ = self.total_pages_ = 0
        self.page_table_ = PageTable(device)
        self.tile_to_page_map_ = OrderedDict()  # (beam, head, tile) -> page, LRU order (oldest first)
        self._free = []
        self._page_refs = {}  # page -> number of table entries sharing it (absent = 1); copy-on-write state
        self.mutex_ = threading.Lock()
        self._ws = {}         # decode scratch, one buffer per stream (the chunk counter and partials live in it)

    # ---- kv_tile_cache.cpp:18-24, 39-50 ------------------------------------------------
    def init(self, num_pages, tile_size, head_dim):
        if min(num_pages, tile_size, head_dim) <= 0:
            raise ValueError("KVTileCache.init: sizes must be positive")
        self.tile_size_, self.head_dim_, self.total_pages_ = int(tile_size), int(head_dim), int(num_pages)
        self._allocate_buffers()
        self._reset_allocator()
        self.page_table_.clear()  # a table configured before a re-init must not keep ids of the old pool

    def _reset_allocator(self):
        """Fresh pool: every page free, no tile mapped, no page shared, scratch re-sized on next use."""
        self.tile_to_page_map_.clear()
        self._free = list(range(self.total_pages_ - 1, -1, -1))
        self._page_refs.clear()
        self._ws = {}

    def configure_table(self, num_beams, num_heads, num_tiles):
        self.page_table_.init(num_beams, num_heads, num_tiles)

    def _dev(self):
        return self._device or torch.device("cuda", torch.cuda.current_device())

    def _allocate_buffers(self):
        shape = (self.total_pages_, self.tile_size_, self.head_dim_)
        dev = self._dev()
        self.key_buffer_ = torch.zeros(shape, dtype=_DTYPES[self.dtype], device=dev)
        self.value_buffer_ = torch.zeros(shape, dtype=_DTYPES[self.dtype], device=dev)
        if self.dtype == "i8":
            self.k_scales_ = torch.ones(shape[:2], dtype=torch.float32, device=dev)
            self.v_scales_ = torch.ones(shape[:2], dtype=torch.float32, device=dev)

    def adopt_buffers(self, key_buffer, value_buffer, k_scales=None, v_scales=None):
        """Use caller-provided pools [total_pages, tile_size, head_dim] (no copy)."""
        assert key_buffer.shape == value_buffer.shape and key_buffer.is_contiguous() and value_buffer.is_contiguous()
        assert key_buffer.dtype == _DTYPES[self.dtype]
        self.total_pages_, self.tile_size_, self.head_dim_ = key_buffer.shape
        self.key_buffer_, self.value_buffer_ = key_buffer, value_buffer
        self.k_scales_, self.v_scales_ = k_scales, v_scales
        self._reset_allocator()

    # ---- kv_tile_cache.cpp:27-37 -------------------------------------------------------
    def resize(self, new_num_pages, new_tile_size):
        with self.mutex_:
            self.tile_size_, self.total_pages_ = int(new_tile_size), int(new_num_pages)
            self._allocate_buffers()
            self._reset_allocator()
            self.page_table_.clear()

    # ---- kv_tile_cache.cpp:52-62 (device addresses) ------------------------------------
    def _page_offset_bytes(self, page_id):
        assert 0 <= page_id < self.total_pages_, "page_id out of range"
        return page_id * self.tile_size_ * self.head_dim_ * self.key_buffer_.element_size()

    def get_key_ptr(self, page_id):
        return self.key_buffer_.data_ptr() + self._page_offset_bytes(page_id)

    def get_value_ptr(self, page_id):
        return self.value_buffer_.data_ptr() + self._page_offset_bytes(page_id)

    # ---- kv_tile_cache.cpp:65-98 -------------------------------------------------------
    def register_tile(self, beam_id, head_id, tile_id):
        with self.mutex_:
            key = (int(beam_id), int(head_id), int(tile_id))
            page = self.tile_to_page_map_.get(key)
            if page is None:
                self._evict_if_needed()
                if not self._free:  # every remaining page is shared by a live beam, or the pool is empty
                    raise RuntimeError("KVTileCache.register_tile: page pool exhausted")
                page = self._free.pop()
                self.tile_to_page_map_[key] = page
                self.page_table_.assign(*key, page)
            self.tile_to_page_map_.move_to_end(key)  # update_lru
            return page

    def _evict_if_needed(self):
        while not self._free and self.tile_to_page_map_:
            last, page = self.tile_to_page_map_.popitem(last=False)  # least recently used
            self.page_table_.remove(*last)
            self._release_page(page)  # a page still shared by another beam is not freed yet

    def sync_page_table_to_gpu(self):
        self.page_table_.sync_to_gpu()

    # ---- kv_tile_cache.cpp:105-125: raw K pool then raw V pool, no header ----------------
    def save_to_file(self, path):
        with open(path, "wb") as f:
            f.write(self.key_buffer_.cpu().numpy().tobytes())
            f.write(self.value_buffer_.cpu().numpy().tobytes())
            if self.dtype == "i8":  # extension: scales follow the reference payload
                f.write(self.k_scales_.cpu().numpy().tobytes())
                f.write(self.v_scales_.cpu().numpy().tobytes())

    def load_from_file(self, path):
        n = self.key_buffer_.numel()
        npdt = {"f16": np.float16, "i8": np.int8, "f32": np.float32}[self.dtype]
        with open(path, "rb") as f:
            k = np.frombuffer(f.read(n * np.dtype(npdt).itemsize), dtype=npdt)
            v = np.frombuffer(f.read(n * np.dtype(npdt).itemsize), dtype=npdt)
            if k.size != n or v.size != n:
                raise RuntimeError(f"Failed to read KV pools from file: {path}")
            self.key_buffer_.copy_(torch.from_numpy(k.copy()).view(self.key_buffer_.shape))
            self.value_buffer_.copy_(torch.from_numpy(v.copy()).view(self.value_buffer_.shape))
            if self.dtype == "i8":
                ns = self.k_scales_.numel()
                ks = np.frombuffer(f.read(ns * 4), dtype=np.float32)
                vs = np.frombuffer(f.read(ns * 4), dtype=np.float32)
                if ks.size == ns and vs.size == ns:
                    self.k_scales_.copy_(torch.from_numpy(ks.copy()).view(self.k_scales_.shape))
                    self.v_scales_.copy_(torch.from_numpy(vs.copy()).view(self.v_scales_.shape))

    # ---- GPU <-> CPU tile offload in the reference's CPU tile-store format --------------------
    # KVTileCacheCPU<T>::save / load (kv_cache/kv_tile_cache_cpu.cpp:89-123): int32 count, then per tile
    # {TileIndex = 3 x int32 (batch, head, tile), tile_size x T payload}.  The CPU store has no K/V
    # distinction, so a tile's payload here is its K page followed by its V page (int8
    # caches append the two f32 scale rows); the file can be loaded by the reference's own
    # KVTileCacheCPU with tile_size = tile_payload_bytes() / sizeof(T).
    def tile_payload_bytes(self):
        page = self.tile_size_ * self.head_dim_ * self.key_buffer_.element_size()
        return 2 * page + (2 * self.tile_size_ * 4 if self.dtype == "i8" else 0)

    def _payload_of_pages(self, pages):
        idx = torch.as_tensor(pages, dtype=torch.int64, device=self.key_buffer_.device)
        n = idx.numel()
        parts = [self.key_buffer_.index_select(0, idx).view(torch.uint8).reshape(n, -1),
                 self.value_buffer_.index_select(0, idx).view(torch.uint8).reshape(n, -1)]
        if self.dtype == "i8":
            parts += [self.k_scales_.index_select(0, idx).view(torch.uint8).reshape(n, -1),
                      self.v_scales_.index_select(0, idx).view(torch.uint8).reshape(n, -1)]
        return torch.cat(parts, dim=1).cpu().numpy()

    def save_tiles_cpu_format(self, path):
        pt = self.page_table_
        pt.flush()
        flat = np.nonzero((pt.host_table_ >= 0) & (pt.host_table_ < self.total_pages_))[0]
        pages = pt.host_table_[flat]
        payload = self._payload_of_pages(pages) if flat.size else np.zeros((0, self.tile_payload_bytes()), np.uint8)
        hn = pt.num_heads_ * pt.num_tiles_
        rec = np.zeros(flat.size, dtype=np.dtype([("idx", "<i4", 3), ("data", "u1", self.tile_payload_bytes())]))
        rec["idx"][:, 0] = flat // hn
        rec["idx"][:, 1] = (flat % hn) // pt.num_tiles_
        rec["idx"][:, 2] = flat % pt.num_tiles_
        rec["data"] = payload
        try:
            with open(path, "wb") as f:
                f.write(np.int32(flat.size).tobytes())
                f.write(rec.tobytes())
        except OSError:
            raise RuntimeError(f"Failed to open file for saving: {path}")
        return int(flat.size)

    def load_tiles_cpu_format(self, path):
        """Tiles of the file are (re)mapped to pages (fresh pages from the free list for unmapped tiles)
        and their K/V bytes uploaded in one batched scatter."""
        nb = self.tile_payload_bytes()
        try:
            with open(path, "rb") as f:
                raw = f.read()
        except OSError:
            raise RuntimeError(f"Failed to open file for loading: {path}")
        count = int(np.frombuffer(raw[:4], dtype="<i4")[0])
        rec = np.frombuffer(raw, dtype=np.dtype([("idx", "<i4", 3), ("data", "u1", nb)]), count=count, offset=4)
        keys = [(int(b), int(h), int(t)) for b, h, t in rec["idx"]]
        if len(set(keys)) != len(keys):
            raise RuntimeError(f"duplicate tile index in {path}")
        new = sum(1 for k in keys if k not in self.tile_to_page_map_)
        evictable = sum(1 for k, p in self.tile_to_page_map_.items()
                        if k not in set(keys) and self._page_refs.get(p, 1) == 1)
        if new > len(self._free) + evictable:
            # registering them would evict tiles of this same file and alias two tiles to one page
            raise RuntimeError(f"KVTileCache.load_tiles_cpu_format: {count} tiles do not fit the page pool "
                               f"({len(self._free)} free + {evictable} evictable pages)")
        for k in keys:  # tiles of the file that are already mapped become most-recently-used first
            if k in self.tile_to_page_map_:
                self.tile_to_page_map_.move_to_end(k)
        pages = [self.register_tile(*k) for k in keys]
        if not pages:
            return 0
        dev = self.key_buffer_.device
        idx = torch.tensor(pages, dtype=torch.int64, device=dev)
        data = torch.from_numpy(np.ascontiguousarray(rec["data"])).to(dev)
        page = self.tile_size_ * self.head_dim_ * self.key_buffer_.element_size()
        self.key_buffer_.view(torch.uint8).reshape(self.total_pages_, page).index_copy_(0, idx, data[:, :page].contiguous())
        self.value_buffer_.view(torch.uint8).reshape(self.total_pages_, page).index_copy_(
            0, idx, data[:, page:2 * page].contiguous())
        if self.dtype == "i8":
            sb = self.tile_size_ * 4
            self.k_scales_.view(torch.uint8).reshape(self.total_pages_, sb).index_copy_(
                0, idx, data[:, 2 * page:2 * page + sb].contiguous())
            self.v_scales_.view(torch.uint8).reshape(self.total_pages_, sb).index_copy_(
                0, idx, data[:, 2 * page + sb:].contiguous())
        self.page_table_.flush()
        return count

    # ---- hot path: append (get_write_ptr + row write, hpp:29-34) -------------------------
    def append(self, new_k, new_v, positions, beam_ids=None):
        """new_k/new_v: [R, H, D] device tensors (f16 or f32), or page-locked host tensors (uploaded on the copy
        stream, attention.HostPipe); positions: [R] int32 device."""
        pt = self.page_table_
        table = pt.device_data()
        releases = []
        if not new_k.is_cuda:
            from .attention import HostPipe
            assert new_k.is_pinned() and new_v.is_pinned(), "host new_k/new_v must be page-locked"
            pipe = HostPipe.get(table.device)
            with torch.cuda.device(table.device):
                new_k, rk = pipe.upload(new_k, "new_k")
                new_v, rv = pipe.upload(new_v, "new_v")
            releases = [rk, rv]
        R = new_k.shape[0]
        assert new_k.is_contiguous() and new_v.is_contiguous() and new_k.shape == new_v.shape
        assert new_k.shape[1] == pt.num_heads_ and new_k.shape[2] == self.head_dim_
        lib = _cabi.lib()
        common = (table.data_ptr(), pt.num_beams_, pt.num_heads_, pt.num_tiles_, self.total_pages_,
                  self.tile_size_, self.head_dim_, new_k.data_ptr(), new_v.data_ptr(),
                  _cabi.ptr(beam_ids), positions.data_ptr(), R, _cabi.stream())
        with torch.cuda.device(table.device):
            if self.dtype == "f32":
                assert new_k.dtype == torch.float32
                st = lib.pa_kv_append_f32(self.key_buffer_.data_ptr(), self.value_buffer_.data_ptr(), *common)
            elif self.dtype == "i8":
                assert new_k.dtype == torch.float32
                st = lib.pa_kv_append_f32_i8(self.key_buffer_.data_ptr(), self.value_buffer_.data_ptr(),
                                             self.k_scales_.data_ptr(), self.v_scales_.data_ptr(), *common)
            elif new_k.dtype == torch.float16:
                st = lib.pa_kv_append_f16(self.key_buffer_.data_ptr(), self.value_buffer_.data_ptr(), *common)
            else:
                assert new_k.dtype == torch.float32
                st = lib.pa_kv_append_f32_f16(self.key_buffer_.data_ptr(), self.value_buffer_.data_ptr(), *common)
        _cabi.check(st, "pa_kv_append")
        with torch.cuda.device(table.device):
            for rel in releases:
                rel()

    # ---- beam search: shared-prefix pages with copy-on-write (north star; no reference code) ---
    def _refs(self):
        return self._page_refs

    def fork_beam(self, src_beam, dst_beam, num_tiles=None):
        """Make table row `dst_beam` share the pages of `src_beam` (first `num_tiles` tiles of every head):
        the beam-search fork.  No K/V bytes move; the shared pages become copy-on-write."""
        pt = self.page_table_
        nt = pt.num_tiles_ if num_tiles is None else int(num_tiles)
        refs = self._refs()
        with self.mutex_:
            for h in range(pt.num_heads_):
                for t in range(nt):
                    old = pt.lookup(dst_beam, h, t)
                    page = pt.lookup(src_beam, h, t)
                    if old == page:
                        continue
                    if old >= 0:
                        self._release_page(old)
                    if page >= 0:
                        refs[page] = refs.get(page, 1) + 1
                    pt.assign(dst_beam, h, t, page)
                    key = (int(dst_beam), h, t)
                    if page >= 0:
                        self.tile_to_page_map_[key] = page
                    else:
                        self.tile_to_page_map_.pop(key, None)

    def _release_page(self, page):
        refs = self._refs()
        n = refs.get(page, 1) - 1
        if n <= 0:
            refs.pop(page, None)
            self._free.append(page)
        else:
            refs[page] = n

    def page_refcount(self, page):
        return self._refs().get(int(page), 1)

    def append_cow(self, new_k, new_v, positions, beam_ids=None):
        """append() for beams that may share pages: before the row write, every (beam, head) whose
        target page is shared (refcount > 1) gets a private copy -- a fresh page from the free list, one
        batched device copy (pa_kv_copy_pages), one batched page-table update.  `positions` (and
        `beam_ids`) are HOST integer sequences: page management is host logic, as in the reference
        (kv_tile_cache.cpp:65-98)."""
        pt = self.page_table_
        pos_h = [int(p) for p in positions]
        rows = list(range(len(pos_h))) if beam_ids is None else [int(b) for b in beam_ids]
        refs = self._refs()
        src, dst = [], []
        with self.mutex_:
            for r, pos in zip(rows, pos_h):
                if pos < 0:
                    continue
                t = pos // self.tile_size_
                for h in range(pt.num_heads_):
                    page = pt.lookup(r, h, t)
                    if page >= 0 and refs.get(page, 1) > 1:
                        if not self._free:
                            raise RuntimeError("KVTileCache.append_cow: page pool exhausted")
                        new_page = self._free.pop()
                        refs[page] -= 1
                        if refs[page] <= 1:
                            refs.pop(page)
                        src.append(page)
                        dst.append(new_page)
                        pt.assign(r, h, t, new_page)
                        self.tile_to_page_map_[(r, h, t)] = new_page
        dev = self.key_buffer_.device
        if src:
            d_src = torch.tensor(src, dtype=torch.int32).to(dev)
            d_dst = torch.tensor(dst, dtype=torch.int32).to(dev)
            with torch.cuda.device(dev):
                _cabi.check(_cabi.lib().pa_kv_copy_pages(
                    self.key_buffer_.data_ptr(), self.value_buffer_.data_ptr(), _cabi.ptr(self.k_scales_),
                    _cabi.ptr(self.v_scales_), d_src.data_ptr(), d_dst.data_ptr(), len(src), self.total_pages_,
                    self.tile_size_, self.head_dim_, self.key_buffer_.element_size(), _cabi.stream()),
                    "pa_kv_copy_pages")
        d_pos = torch.tensor(pos_h, dtype=torch.int32).to(dev)
        d_beam = None if beam_ids is None else torch.tensor(rows, dtype=torch.int32).to(dev)
        self.append(new_k, new_v, d_pos, d_beam)
        return len(src)

    # ---- hot path: gather (KVTileCache::get materialised, hpp:21-26) ----------------------
    def gather(self, which="k", beam_ids=None, rows=None, fill_byte=0):
        pt = self.page_table_
        table = pt.device_data()
        pool = self.key_buffer_ if which == "k" else self.value_buffer_
        R = rows if rows is not None else (beam_ids.numel() if beam_ids is not None else pt.num_beams_)
        dense = torch.empty((R, pt.num_heads_, pt.num_tiles_ * self.tile_size_, self.head_dim_),
                            dtype=pool.dtype, device=pool.device)
        with torch.cuda.device(pool.device):
            _cabi.check(_cabi.lib().pa_kv_gather(pool.data_ptr(), dense.data_ptr(), table.data_ptr(),
                                                 pt.num_beams_, pt.num_heads_, pt.num_tiles_,
                                                 self.total_pages_, self.tile_size_, self.head_dim_,
                                                 pool.element_size(), _cabi.ptr(beam_ids), R, fill_byte,
                                                 _cabi.stream()), "pa_kv_gather")
        return dense

    def workspace(self, B):
        """Scratch for the decode kernels (chunk counter + partials): one buffer per (cache, stream), so decode
        calls on different streams never share it; sized on the cache's own device."""
        dev = self.key_buffer_.device
        with torch.cuda.device(dev):
            need = _cabi.lib().pa_decode_workspace_bytes(B, self.page_table_.num_heads_, self.head_dim_,
                                                         self.page_table_.num_tiles_, self.tile_size_)
            key = torch.cuda.current_stream(dev).cuda_stream
        ws = self._ws.get(key)
        if ws is None or ws.numel() < need:
            ws = self._ws[key] = torch.empty(need, dtype=torch.uint8, device=dev)
        return ws
