"""Row partitioning of the fake-quant path across the GPUs of one box.

Every statistic of the path is local to one reduction row (per token for
activations and K/V, per output channel for weights — reference
utils_quant.py:56,118-124), so shards are contiguous row ranges and no
data-path collective exists.  ``layerwise=True`` is the
one exception: it would need an all-reduce(max/min) of one scalar and is not
sharded here.
"""
from __future__ import annotations


def row_partition(rows: int, world_size: int, rank: int) -> tuple[int, int]:
    """(first_row, n_rows) of ``rank``'s contiguous shard; shards differ by at
    most one row and cover [0, rows) exactly."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError(f"bad rank {rank} for world size {world_size}")
    if rows < 0:
        raise ValueError("rows must be non-negative")
    base, extra = divmod(rows, world_size)
    start = rank * base + min(rank, extra)
    return start, base + (1 if rank < extra else 0)


def shard_rows(t, world_size: int, rank: int, layerwise: bool = False):
    """The rank's row shard of a tensor whose reduction rows are its leading
    dims (ndim <= 3: all but the last dim flattened).  Returns a view."""
    if layerwise:
        raise ValueError("layerwise fake-quant reduces over the whole tensor and does not shard by rows")
    if t.dim() < 2:
        raise ValueError("need at least [rows, cols]")
    flat = t.reshape(-1, t.shape[-1])
    start, n = row_partition(flat.shape[0], world_size, rank)
    return flat[start:start + n]
