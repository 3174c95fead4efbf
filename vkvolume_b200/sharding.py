"""Host-side logic of the multi-GPU decomposition: pure index arithmetic plus the two
torch.distributed exchanges.  Works on CPU tensors with the gloo backend (how tests/test_sharding_gloo.py runs it)
and on CUDA tensors with NCCL (how bench.py runs it) — there is no compute here.

* Ray casting shards by image tiles: tile t of the row-major tile list goes to rank t % world (round-robin, because
  ESS/ERT make per-pixel cost wildly non-uniform); every rank holds a full replica and stores its tiles straight into
  rank 0's framebuffer (peer mapping), so that path has no collective at all.
* Frame sequences (an orbit, an animation) shard by frames instead: step s of an N-rank job renders views s*N .. s*N + N - 1, view
  s*N + r on rank r, into the rank's own HBM (or delivers it to its own pinned host buffer; bench.py --gather-frames stores it into
  slot r of a ring of N frames in rank 0's HBM instead); no collective, no exchange.  reduce_step_times says when such a job is done.
* The TF-change rebuild shards the O(N) occupancy pass by z-slabs of blocks.  In the library's own group
  (vkv_update_transfer_function_sharded, csrc/group.cu) the isotropic distance map is sharded too: x and y passes on the rank's slab,
  exchange of the xy-intermediate slabs, z pass on the rank's share of the block rows (row_range), exchange of the result rows — peer
  copies over NVLink and barriers in peer memory.  The torch.distributed helpers below (all-gather of occupancy slabs, all-reduce
  of the count) are the same decomposition with a collective library doing the exchange; the gloo tests use them on CPU.
"""
from __future__ import annotations


def slab_size(depth_blocks: int, world: int) -> int:
    """Block slices per rank (the last ranks may get fewer, or none)."""
    return (depth_blocks + world - 1) // world


def slab_range(rank: int, world: int, depth_blocks: int) -> tuple[int, int]:
    """(first block slice, number of block slices) of `rank`; count is 0 for ranks beyond the map."""
    s = slab_size(depth_blocks, world)
    z0 = min(rank * s, depth_blocks)
    return z0, max(0, min(s, depth_blocks - z0))


def row_range(rank: int, world: int, height_blocks: int) -> tuple[int, int]:
    """(first block row, number of block rows) whose z lines `rank` transforms in the sharded distance-map build
    (vkv_update_transfer_function_sharded: x and y passes on the z-slab, exchange, z pass on these rows, exchange)."""
    return slab_range(rank, world, height_blocks)


def tiles_of_rank(rank: int, world: int, n_tiles: int) -> range:
    """Tiles rendered by `rank` — what vkv_render_tiles(tile_first=rank, tile_stride=world) covers."""
    return range(rank, n_tiles, world)


def view_of_rank(step: int, rank: int, world: int) -> int:
    """Frames decomposition: the view (index into the frame sequence) `rank` renders at `step`."""
    return step * world + rank


def frame_slot_offset(rank: int, width: int, height: int) -> int:
    """Byte offset of `rank`'s slot in the ring of RGBA8 frames on rank 0 (slot r holds view step*world + r)."""
    return rank * width * height * 4


def n_tiles(width: int, height: int, tile_w: int, tile_h: int) -> int:
    return ((width + tile_w - 1) // tile_w) * ((height + tile_h - 1) // tile_h)


def all_gather_occupancy(full, rank: int, world: int, map_extent, gather_buf=None, group=None):
    """In place: `full` (flat uint8 tensor of Wb*Hb*Db cells, this rank's slab rows already written) receives every other
    rank's slab rows.  Slabs are padded to the common slab size for all_gather_into_tensor."""
    import torch
    import torch.distributed as dist

    Wb, Hb, Db = map_extent
    plane = Wb * Hb
    s = slab_size(Db, world)
    z0, zc = slab_range(rank, world, Db)
    if gather_buf is None:
        gather_buf = torch.empty(world * s * plane, dtype=torch.uint8, device=full.device)
    mine = torch.zeros(s * plane, dtype=torch.uint8, device=full.device)
    mine[: zc * plane] = full[z0 * plane:(z0 + zc) * plane]
    dist.all_gather_into_tensor(gather_buf, mine, group=group)
    for r in range(world):
        rz0, rzc = slab_range(r, world, Db)
        if rzc and r != rank:
            full[rz0 * plane:(rz0 + rzc) * plane] = gather_buf[r * s * plane: r * s * plane + rzc * plane]
    return full


def all_reduce_count(count_tensor, group=None):
    """Sum of the per-slab occupied-voxel counts (int64 tensor of one element), in place."""
    import torch.distributed as dist

    dist.all_reduce(count_tensor, group=group)
    return count_tensor


def reduce_step_times(step_ms, independent: bool, group=None):
    """Job time of K timed steps from every rank's per-step device times (1-D float64 tensor, one entry per step).

    independent = False (tiles of one frame): a step is done when its slowest rank is -> per-step max over ranks, summed.
    independent = True (frames: one view per rank per step, no exchange between ranks): the job is done when the slowest rank has
    rendered its K views -> per-rank sum, max over ranks; the per-step-max figure ("lockstep": as if a barrier followed every frame)
    is returned beside it.  Returns (total_ms, lockstep_total_ms or None); every rank gets the same numbers."""
    import torch.distributed as dist

    if not independent:
        t = step_ms.clone()
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
        return float(t.sum().item()), None
    lock = step_ms.clone()
    dist.all_reduce(lock, op=dist.ReduceOp.MAX, group=group)
    tot = step_ms.sum().reshape(1)
    dist.all_reduce(tot, op=dist.ReduceOp.MAX, group=group)
    return float(tot.item()), float(lock.sum().item())
