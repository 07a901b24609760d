"""Row sharding of the inner domain over the GPUs of one box.

The reference parallelises over rows of the inner domain
(``tbb::blocked_range<size_t>(0, dim_in_0)``, ``horizon_comp.cpp:739-744``;
``shadow_comp.cpp:390-394``): no cell reads another cell's result, so contiguous
row blocks are independent and -- the output being C-ordered ``[y][x][azim]`` --
each shard is one contiguous byte range.  One process per GPU computes its
block against a replicated DEM/BVH; a single all-gather (NCCL over NVLink on
the GPU box, gloo in CPU tests) joins the blocks.
"""
import numpy as np


def row_shards(num_rows, world_size):
    """[(row_begin, row_end)] per rank: equal blocks of ceil(rows / world) rows,
    the trailing ranks getting the (possibly empty) remainder."""
    per = -(-int(num_rows) // int(world_size)) if num_rows > 0 else 0
    out = []
    for r in range(world_size):
        b = min(r * per, num_rows)
        e = min(b + per, num_rows)
        out.append((b, e))
    return out


def padded_rows(num_rows, world_size):
    """Rows per rank after padding to an equal count (all-gather needs it)."""
    return -(-int(num_rows) // int(world_size)) if num_rows > 0 else 0


def allgather_rows(local_block, num_rows, dist, group=None):
    """All-gather equal-sized row blocks and trim the padding.

    ``local_block``: tensor ``[padded_rows, ...]`` holding this rank's rows
    first (rows past the shard end are padding).  Returns ``[num_rows, ...]``.
    ``dist`` is ``torch.distributed`` (injected so CPU tests can use gloo)."""
    import torch
    world = dist.get_world_size(group)
    per = local_block.shape[0]
    full = torch.empty((world * per,) + tuple(local_block.shape[1:]), dtype=local_block.dtype,
                       device=local_block.device)
    dist.all_gather_into_tensor(full, local_block.contiguous(), group=group)
    return full[:num_rows]
