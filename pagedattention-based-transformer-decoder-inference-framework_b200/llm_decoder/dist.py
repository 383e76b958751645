"""Multi-GPU modes of the decode hot path (one process per GPU, torch.distributed / NCCL).

The reference has no multi-device code at all; BASELINE.json's north star adds two
modes, both implemented here:

* batch sharding -- rows (sequences) are independent (grid(B,H), attention_tile_launcher.hpp:53),
  so each rank owns a contiguous block of rows with its pages and page-table rows.  No
  collective on the data path.  `shard_range` keeps beam groups on one rank.
* split-KV of ONE long sequence -- rank r holds a contiguous range of the sequence's pages,
  runs the same decode kernel over its range emitting un-normalised partials (m, l, O)
  (pa_paged_decode_{f16,i8}_partial); the partials ([H, D+2] floats = 16.6 KB per rank at the
  Llama-7B shape) are exchanged over NVLink and LSE-combined on every rank.  Three exchange forms
  (split_kv_decode): NCCL all-gather + combine (`NcclCombine`, pa_nccl_allgather_combine: the north
  star's baseline), a stand-alone peer-memory exchange kernel, and decode + merge + exchange fused
  into ONE launch (`PeerExchange.decode`, pa_paged_decode_*_splitkv).
"""
import torch
import torch.distributed as dist

from . import attention as _att


def shard_range(n_items, world_size, rank, group=1):
    """Contiguous [begin, end) of `n_items` rows for `rank`, in whole groups of `group` rows
    (beam groups are never split across GPUs).  Remainder groups go to the lowest ranks."""
    if n_items % group:
        raise ValueError("n_items must be a multiple of the beam group size")
    n_groups = n_items // group
    base, rem = divmod(n_groups, world_size)
    g0 = rank * base + min(rank, rem)
    g1 = g0 + base + (1 if rank < rem else 0)
    return g0 * group, g1 * group


def page_range(num_tiles, world_size, rank):
    """Contiguous tile (page index) range of one sequence owned by `rank` in split-KV mode."""
    return shard_range(num_tiles, world_size, rank)


def pack_partials(part_m, part_l, part_o):
    """(m [rows], l [rows], O [rows, D]) -> one [rows, D+2] message: O | m | l."""
    rows, D = part_o.shape[-2], part_o.shape[-1]
    msg = torch.empty((rows, D + 2), dtype=torch.float32, device=part_o.device)
    msg[:, :D] = part_o.reshape(rows, D)
    msg[:, D] = part_m.reshape(rows)
    msg[:, D + 1] = part_l.reshape(rows)
    return msg


def unpack_partials(gathered):
    """[world, rows, D+2] -> (m [world, rows], l [world, rows], O [world, rows, D]) contiguous."""
    D = gathered.shape[-1] - 2
    return (gathered[..., D].contiguous(), gathered[..., D + 1].contiguous(), gathered[..., :D].contiguous())


def gather_partials(msg, group=None):
    """all_gather of every rank's packed partials -> [world, rows, D+2]."""
    world = dist.get_world_size(group)
    msg = msg.contiguous()
    out = torch.empty((world * msg.shape[0],) + tuple(msg.shape[1:]), dtype=msg.dtype, device=msg.device)
    dist.all_gather_into_tensor(out, msg, group=group)  # concatenation along dim 0 (gloo and nccl)
    return out.view((world,) + tuple(msg.shape))


def combine_gathered(gathered, combine_fn=None):
    """LSE-combine [world, rows, D+2] partials into [rows, D].  On CUDA tensors this is
    pa_lse_combine; there is no CPU implementation in the product (tests inject the oracle)."""
    pm, pl, po = unpack_partials(gathered)
    if combine_fn is not None:
        return combine_fn(pm, pl, po)
    if not gathered.is_cuda:
        raise RuntimeError("combine_gathered: the LSE combine runs on the GPU only (libpa_b200.so)")
    return _att.lse_combine(pm, pl, po)


class PeerExchange:
    """Exchange buffers of all ranks mapped into this process over CUDA IPC (NVLink P2P), for the
    fused exchange+combine kernel pa_splitkv_exchange_combine.  One instance per (rows, D) shape."""

    def __init__(self, rows, head_dim, group=None, device=None):
        import ctypes as C

        from . import _cabi
        self._cabi, self._C = _cabi, C
        self.rows, self.D, self.group = rows, head_dim, group
        self._dist = dist.is_available() and dist.is_initialized()   # a single process is a world of one rank
        self.world, self.rank = (dist.get_world_size(group), dist.get_rank(group)) if self._dist else (1, 0)
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        lib = _cabi.lib()
        nbytes = lib.pa_splitkv_exchange_bytes(self.world, rows, head_dim)
        with torch.cuda.device(self.device):
            self._own = C.c_void_p()
            handle = C.create_string_buffer(64)
            _cabi.check(lib.pa_p2p_alloc(nbytes, C.byref(self._own), handle), "pa_p2p_alloc")
            handles = [bytes(handle.raw)]
            if self._dist:
                handles = [None] * self.world
                dist.all_gather_object(handles, bytes(handle.raw), group=group)
            self._peers, ptrs = [], []
            for r, h in enumerate(handles):
                if r == self.rank:
                    ptrs.append(self._own.value)
                    continue
                p = C.c_void_p()
                _cabi.check(lib.pa_p2p_open(C.create_string_buffer(h, 64), C.byref(p)), "pa_p2p_open")
                self._peers.append(p)
                ptrs.append(p.value)
            self.peer_ptrs = torch.tensor(ptrs, dtype=torch.int64).to(self.device)
            self.status = torch.zeros(1, dtype=torch.int32, device=self.device)
            # [rows] step counters advanced by the kernels + [rows] self-resetting row-completion counters
            self.epochs = torch.zeros(2 * rows, dtype=torch.int32, device=self.device)
        if self._dist:
            dist.barrier(group=group)  # every buffer is mapped (and zeroed) before the first kernel runs

    def combine(self, part_m, part_l, part_o, out=None, lse_out=None):
        """part_m/l [rows], part_o [rows, D] (this rank's partials) -> out [rows, D] on every rank."""
        _cabi = self._cabi
        if out is None:
            out = torch.empty((self.rows, self.D), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _cabi.check(_cabi.lib().pa_splitkv_exchange_combine(
                part_m.data_ptr(), part_l.data_ptr(), part_o.data_ptr(), self.peer_ptrs.data_ptr(), self.rank,
                self.world, self.rows, self.D, self.epochs.data_ptr(), out.data_ptr(), _cabi.ptr(lse_out),
                self.status.data_ptr(), _cabi.stream()), "pa_splitkv_exchange_combine")
        return out

    def decode(self, q, kv_cache, B, T_local, temperature=1.0, beam_ids=None, ctx_lens=None, out=None):
        """pa_paged_decode_{f16,i8}_splitkv: ONE launch -- the streaming kernel over this rank's pages; the warp that
        finishes a row merges it and sends it to every peer (flag-in-data packets, never blocks); rows are received
        and LSE-combined at the end of the same kernel."""
        _cabi = self._cabi
        pt = kv_cache.page_table_
        H, D = pt.num_heads_, kv_cache.head_dim_
        assert B * H == self.rows and D == self.D
        if out is None:
            out = torch.empty((B, H, D), dtype=torch.float32, device=self.device)
        ws = kv_cache.workspace(B)
        lib = _cabi.lib()
        pools = (kv_cache.key_buffer_.data_ptr(), kv_cache.value_buffer_.data_ptr())
        if kv_cache.dtype == "i8":
            fn, name = lib.pa_paged_decode_i8_splitkv, "pa_paged_decode_i8_splitkv"
            pools += (kv_cache.k_scales_.data_ptr(), kv_cache.v_scales_.data_ptr())
        else:
            fn, name = lib.pa_paged_decode_f16_splitkv, "pa_paged_decode_f16_splitkv"
        with torch.cuda.device(self.device):
            st = fn(q.data_ptr(), out.data_ptr(), *pools, pt.device_data().data_ptr(), pt.num_beams_, H, pt.num_tiles_,
                    kv_cache.total_pages_, _cabi.ptr(beam_ids), _cabi.ptr(ctx_lens), B, T_local, D, kv_cache.tile_size_,
                    float(temperature), None, None, ws.data_ptr(), ws.numel(), self.peer_ptrs.data_ptr(), self.rank,
                    self.world, self.epochs.data_ptr(), self.status.data_ptr(), _cabi.stream())
        _cabi.check(st, name)
        return out

    def check(self):
        """Raises if a peer ever failed to arrive (synchronises the device: call it at a host sync point, e.g. once
        per generated token when the ids are read back).  A timed-out row was written as NaN and its epoch was not
        advanced, so a missed step can never be mistaken for a result."""
        if int(self.status.item()) != 0:
            raise RuntimeError("split-KV exchange: a peer rank did not arrive within the timeout")

    def close(self):
        lib = self._cabi.lib()
        if self._dist:
            dist.barrier(group=self.group)
        with torch.cuda.device(self.device):
            torch.cuda.synchronize()
            self.check()
            for p in self._peers:
                lib.pa_p2p_close(p)
            self._peers = []
            if self._own is not None:
                lib.pa_p2p_free(self._own)
                self._own = None


class NcclCombine:
    """pa_nccl_*: the library's own NCCL communicator (ncclCommInitRank from an id broadcast through
    torch.distributed) + all-gather of (m, l, O) straight into pa_lse_combine's layout + combine kernel."""

    def __init__(self, rows, head_dim, group=None, device=None):
        import ctypes as C

        from . import _cabi
        self._cabi = _cabi
        self.rows, self.D, self.group = rows, head_dim, group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        lib = _cabi.lib()
        ident = C.create_string_buffer(128)
        if self.rank == 0:
            _cabi.check(lib.pa_nccl_unique_id(ident), "pa_nccl_unique_id")
        box = [bytes(ident.raw)]
        dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        self._comm = C.c_void_p()
        with torch.cuda.device(self.device):
            _cabi.check(lib.pa_nccl_init(C.create_string_buffer(box[0], 128), self.rank, self.world, C.byref(self._comm)),
                        "pa_nccl_init")
            self._gather = torch.empty(lib.pa_nccl_gather_bytes(self.world, rows, head_dim), dtype=torch.uint8,
                                       device=self.device)

    def combine(self, part_m, part_l, part_o, out=None, lse_out=None):
        _cabi = self._cabi
        if out is None:
            out = torch.empty((self.rows, self.D), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _cabi.check(_cabi.lib().pa_nccl_allgather_combine(
                self._comm, self.world, part_m.data_ptr(), part_l.data_ptr(), part_o.data_ptr(), self.rows, self.D,
                self._gather.data_ptr(), self._gather.numel(), out.data_ptr(), _cabi.ptr(lse_out), _cabi.stream()),
                "pa_nccl_allgather_combine")
        return out

    def close(self):
        if self._comm:
            with torch.cuda.device(self.device):
                torch.cuda.synchronize()
                self._cabi.lib().pa_nccl_destroy(self._comm)
            self._comm = None


def split_kv_decode(q, kv_cache, B, T_local, temperature=1.0, beam_ids=None, ctx_lens=None, group=None,
                    exchange=None, fused=False):
    """Decode attention of `B` rows whose KV pages are split across the ranks of `group`.
    `kv_cache` holds THIS rank's pages (table row b = the rank's tile range of sequence b);
    T_local / ctx_lens are the token counts held locally.  Returns out [B, H, D] on every rank.
    exchange=None: all-gather of the partials through torch.distributed followed by pa_lse_combine.
    exchange=NcclCombine: pa_nccl_allgather_combine (the library's own NCCL path, no pack / unpack passes).
    exchange=PeerExchange: partial kernel + ONE peer-memory exchange+combine kernel; with fused=True decode, row
    merge, exchange and combine are ONE launch (pa_paged_decode_*_splitkv; the two peer-memory forms use the same
    buffers but must not be mixed within one PeerExchange because they advance the same step counters)."""
    if isinstance(exchange, PeerExchange) and fused:
        return exchange.decode(q, kv_cache, B, T_local, temperature, beam_ids, ctx_lens)
    pm, pl, po = _att.paged_decode_partial(q, kv_cache, B, T_local, temperature, beam_ids, ctx_lens)
    H, D = po.shape[1], po.shape[2]
    if exchange is not None:
        return exchange.combine(pm.reshape(-1), pl.reshape(-1), po.reshape(B * H, D)).reshape(B, H, D)
    msg = pack_partials(pm, pl, po.reshape(B * H, D))
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        gathered = gather_partials(msg, group)
    else:
        gathered = msg.unsqueeze(0)
    return combine_gathered(gathered).reshape(B, H, D)
