"""PageTable -- dense device-resident int32 map [num_beams][num_heads][num_tiles] -> page id.

Mirrors kv_cache/page_table.hpp:5-49 / page_table.cpp:6-67 (same method names and
argument meaning).  Differences that are decisions, not accidents:
`assign`/`remove` update a host mirror and are flushed to the device in ONE batched
kernel (pa_page_table_update) instead of one blocking 4-byte cudaMemcpy per entry, and
`sync_to_gpu` uploads the mirror WITH the assignments (the reference's mirror never sees
them, so its sync erases the table).
"""
import numpy as np
import torch

from . import _cabi


class PageTable:
    def __init__(self, device=None):
        self.num_beams_ = self.num_heads_ = self.num_tiles_ = 0
        self.total_entries_ = 0
        self.d_table_ = None
        self.host_table_ = np.empty(0, dtype=np.int32)
        self._pending = {}
        self._device = torch.device(device) if device is not None else None

    # page_table.cpp:14-26
    def init(self, num_beams, num_heads, num_tiles):
        if min(num_beams, num_heads, num_tiles) <= 0:
            raise ValueError("PageTable.init: dimensions must be positive")
        self.num_beams_, self.num_heads_, self.num_tiles_ = int(num_beams), int(num_heads), int(num_tiles)
        self.total_entries_ = self.num_beams_ * self.num_heads_ * self.num_tiles_
        dev = self._device or torch.device("cuda", torch.cuda.current_device())
        self.d_table_ = torch.empty(self.total_entries_, dtype=torch.int32, device=dev)
        self.host_table_ = np.full(self.total_entries_, -1, dtype=np.int32)
        self._pending.clear()
        self._clear_device()

    def _clear_device(self):
        with torch.cuda.device(self.d_table_.device):
            _cabi.check(_cabi.lib().pa_page_table_clear(self.d_table_.data_ptr(), self.total_entries_,
                                                        _cabi.stream()), "pa_page_table_clear")

    # page_table.cpp:41-47
    def clear(self):
        if self.d_table_ is not None:
            self.host_table_.fill(-1)
            self._pending.clear()
            self._clear_device()

    # page_table.hpp:39-42
    def index(self, beam_id, head_id, tile_id):
        return beam_id * (self.num_heads_ * self.num_tiles_) + head_id * self.num_tiles_ + tile_id

    # page_table.cpp:49-53 (assert on range, as the reference)
    def assign(self, beam_id, head_id, tile_id, page_id):
        idx = self.index(beam_id, head_id, tile_id)
        assert 0 <= idx < self.total_entries_, "PageTable.assign: index out of range"
        self.host_table_[idx] = page_id
        self._pending[idx] = int(page_id)

    def assign_many(self, flat_idx, pages):
        """Vectorised assign: flat_idx / pages int arrays (host)."""
        flat_idx = np.asarray(flat_idx, dtype=np.int64)
        pages = np.asarray(pages, dtype=np.int32)
        assert flat_idx.min(initial=0) >= 0 and flat_idx.max(initial=-1) < self.total_entries_
        self.host_table_[flat_idx] = pages
        self.flush()
        dev = self.d_table_.device
        di = torch.from_numpy(flat_idx.astype(np.int32)).to(dev)
        dp = torch.from_numpy(pages).to(dev)
        with torch.cuda.device(dev):
            _cabi.check(_cabi.lib().pa_page_table_update(self.d_table_.data_ptr(), self.total_entries_,
                                                         di.data_ptr(), dp.data_ptr(), di.numel(),
                                                         _cabi.stream()), "pa_page_table_update")

    # page_table.cpp:64-67 (the reference edits only the host mirror; here the device
    # entry is cleared at the next flush as well)
    def remove(self, beam_id, head_id, tile_id):
        idx = self.index(beam_id, head_id, tile_id)
        self.host_table_[idx] = -1
        self._pending[idx] = -1

    def flush(self):
        """Push queued assign/remove calls to the device in one kernel."""
        if not self._pending:
            return
        dev = self.d_table_.device
        idx = torch.tensor(list(self._pending.keys()), dtype=torch.int32).to(dev)
        pg = torch.tensor(list(self._pending.values()), dtype=torch.int32).to(dev)
        self._pending.clear()
        with torch.cuda.device(dev):
            _cabi.check(_cabi.lib().pa_page_table_update(self.d_table_.data_ptr(), self.total_entries_,
                                                         idx.data_ptr(), pg.data_ptr(), idx.numel(),
                                                         _cabi.stream()), "pa_page_table_update")

    # page_table.hpp:44-49 -- host-visible form of the device lookup
    def lookup(self, beam_id, head_id, tile_id):
        idx = self.index(beam_id, head_id, tile_id)
        if idx < 0 or idx >= self.total_entries_:
            return -1
        return int(self.host_table_[idx])

    def lookup_device(self, beams, heads, tiles):
        """Batched lookup executed ON the device (pa_page_table_lookup)."""
        self.flush()
        dev = self.d_table_.device
        b = torch.as_tensor(beams, dtype=torch.int32).to(dev)
        h = torch.as_tensor(heads, dtype=torch.int32).to(dev)
        t = torch.as_tensor(tiles, dtype=torch.int32).to(dev)
        out = torch.empty_like(b)
        with torch.cuda.device(dev):
            _cabi.check(_cabi.lib().pa_page_table_lookup(self.d_table_.data_ptr(), self.num_beams_,
                                                         self.num_heads_, self.num_tiles_, b.data_ptr(),
                                                         h.data_ptr(), t.data_ptr(), out.data_ptr(),
                                                         b.numel(), _cabi.stream()), "pa_page_table_lookup")
        return out

    # page_table.cpp:55-57
    def device_data(self):
        self.flush()
        return self.d_table_

    # page_table.cpp:59-62
    def sync_to_gpu(self):
        self._pending.clear()
        self.d_table_.copy_(torch.from_numpy(self.host_table_), non_blocking=False)

    def load_host_table(self, table):
        """Replace the whole map from a host array [num_beams, num_heads, num_tiles]."""
        table = np.ascontiguousarray(table, dtype=np.int32).reshape(-1)
        assert table.size == self.total_entries_
        self.host_table_[:] = table
        self.sync_to_gpu()
