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


# ---- block-interleaved sharding (cost-balanced) ---------------------------------------------
# Contiguous row blocks cost different amounts (rim rows see half the terrain of centre rows), and the
# slowest GPU sets the step time.  Dealing the inner domain out in 4-row blocks -- block b to shard
# b % world -- gives every GPU the same mix; each shard stores its blocks back to back (the kernels'
# packed mode), ONE all-gather joins the shards, and block j of shard r is block j * world + r of the
# result: a reshape + permute, no arithmetic.

BLOCK_ROWS = 4


def shard_block_rows(num_rows, rank, world_size):
    """Rows (whole 4-row blocks, the last one possibly padding) shard `rank` stores."""
    blocks = -(-int(num_rows) // BLOCK_ROWS)
    return BLOCK_ROWS * ((blocks - rank + world_size - 1) // world_size) if blocks > rank else 0


def padded_block_rows(num_rows, world_size):
    """Rows per shard after padding to an equal count (shard 0 has the most blocks)."""
    return shard_block_rows(num_rows, 0, world_size)


def shard_row_indices(num_rows, rank, world_size):
    """Inner-domain row of every row of shard `rank`'s packed buffer (-1: padding)."""
    n = shard_block_rows(num_rows, rank, world_size)
    j = np.arange(n)
    rows = ((j // BLOCK_ROWS) * world_size + rank) * BLOCK_ROWS + j % BLOCK_ROWS
    return np.where(rows < num_rows, rows, -1)


def unpack_blocks(gathered, num_rows, world_size):
    """[world * per, ...] all-gather result of packed shards -> [num_rows, ...] in domain order."""
    per = gathered.shape[0] // world_size
    nb = per // BLOCK_ROWS
    rest = tuple(gathered.shape[1:])
    g = gathered.reshape((world_size, nb, BLOCK_ROWS) + rest)
    order = (1, 0, 2) + tuple(range(3, 3 + len(rest)))
    g = g.permute(*order) if hasattr(g, "permute") else g.transpose(order)
    return g.reshape((nb * world_size * BLOCK_ROWS,) + rest)[:num_rows]
