"""Multi-GPU modes of the decode hot path (one process per GPU, torch.distributed / NCCL).

The reference has no multi-device code at all; BASELINE.json's north star adds two
modes, both implemented here:

* batch sharding -- rows (sequences) are independent (grid(B,H), attention_tile_launcher.hpp:53),
  so each rank owns a contiguous block of rows with its pages and page-table rows.  No
  collective on the data path.  `shard_range` keeps beam groups on one rank.
* split-KV of ONE long sequence -- rank r holds a contiguous range of the sequence's pages,
  runs the same decode kernel over its range emitting un-normalised partials (m, l, O)
  (pa_paged_decode_f16_partial), the partials ([H, D+2] floats = 16.6 KB per rank at the
  Llama-7B shape) are all-gathered over NVLink and merged by pa_lse_combine on every rank.
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
    out = torch.empty((world,) + tuple(msg.shape), dtype=msg.dtype, device=msg.device)
    dist.all_gather_into_tensor(out, msg.contiguous(), group=group)
    return out


def combine_gathered(gathered, combine_fn=None):
    """LSE-combine [world, rows, D+2] partials into [rows, D].  On CUDA tensors this is
    pa_lse_combine; there is no CPU implementation in the product (tests inject the oracle)."""
    pm, pl, po = unpack_partials(gathered)
    if combine_fn is not None:
        return combine_fn(pm, pl, po)
    if not gathered.is_cuda:
        raise RuntimeError("combine_gathered: the LSE combine runs on the GPU only (libpa_b200.so)")
    return _att.lse_combine(pm, pl, po)


def split_kv_decode(q, kv_cache, B, T_local, temperature=1.0, beam_ids=None, ctx_lens=None, group=None):
    """Decode attention of `B` rows whose KV pages are split across the ranks of `group`.
    `kv_cache` holds THIS rank's pages (table row b = the rank's tile range of sequence b);
    T_local / ctx_lens are the token counts held locally.  Returns out [B, H, D] on every rank."""
    pm, pl, po = _att.paged_decode_partial(q, kv_cache, B, T_local, temperature, beam_ids, ctx_lens)
    H, D = po.shape[1], po.shape[2]
    msg = pack_partials(pm, pl, po.reshape(B * H, D))
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        gathered = gather_partials(msg, group)
    else:
        gathered = msg.unsqueeze(0)
    return combine_gathered(gathered).reshape(B, H, D)
