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


# ----------------------------------------------------------------------------------------------------
# Weights quantized by output-channel shard (BASELINE configs[4])
# ----------------------------------------------------------------------------------------------------
# In a data-parallel step every rank holds the same weights.  With weight sharding on, a QuantizeLinear
# forward fake-quantizes only this rank's out/world output channels (each channel is one reduction row:
# no cross-rank statistic exists) and all-gathers the RESULT — int8 codes (1 B/elem), row divisors
# (4 B/row) and the packed STE mask (1/8 B/elem): 1.125 B per weight element instead of the 3.125 B of
# HBM traffic a full local quantization costs — into the blob the tcgen05 GEMM reads.  Bit-identical to
# the unsharded codes (tests/test_gpu_parity.py, tests/test_sharding.py).
_STATE = {"group": None, "world": 1, "rank": 0}


def enable_weight_sharding(group=None):
    """Quantize QuantizeLinear weights by output-channel shard across ``group`` (default: the world
    group) and all-gather codes / divisors / masks.  Call after torch.distributed is initialised."""
    import torch.distributed as dist

    if not dist.is_initialized():
        raise RuntimeError("enable_weight_sharding: torch.distributed is not initialised")
    _STATE["group"] = group if group is not None else dist.group.WORLD
    _STATE["world"] = dist.get_world_size(group)
    _STATE["rank"] = dist.get_rank(group)


def disable_weight_sharding():
    _STATE.update(group=None, world=1, rank=0)


def weight_sharding():
    """(group, world, rank) when active, else None."""
    return (_STATE["group"], _STATE["world"], _STATE["rank"]) if _STATE["world"] > 1 else None


def shardable(rows: int, cols: int, world: int) -> bool:
    """Equal shards whose code, divisor and mask slices keep the alignment the kernels and NCCL's in-place
    all-gather need (16-byte code rows are given by cols % 16 == 0)."""
    if rows % world:
        return False
    n = rows // world
    return (n * cols) % 128 == 0 and n >= 1


def feed_slices(blob, rows: int, cols: int, world: int, rank: int):
    """Views of ``rank``'s slice and of the whole region for each of the three parts of a feed blob
    ([codes int8 rows x cols | divisors f32 rows | packed mask]): [(whole, mine), ...]."""
    from .utils_quant import _feed_layout

    off_e, off_m, total = _feed_layout(rows, cols)
    n = rows // world
    regions = ((0, rows * cols, n * cols), (off_e, rows * 4, n * 4), (off_m, rows * cols // 8, n * cols // 8))
    out = []
    for off, size, mine in regions:
        whole = blob[off: off + size]
        out.append((whole, whole[rank * mine: (rank + 1) * mine]))
    return out


def all_gather_feed(blob, rows: int, cols: int, group, world: int, rank: int):
    """In-place all-gather of the three regions (each rank has written its own slices)."""
    import torch.distributed as dist

    parts = feed_slices(blob, rows, cols, world, rank)
    mgr = None
    if blob.is_cuda:   # NCCL: one group launch for the three regions
        try:
            from torch.distributed.distributed_c10d import _coalescing_manager as mgr
        except ImportError:
            mgr = None
    if mgr is not None:
        with mgr(group=group, device=blob.device, async_ops=False):
            for whole, mine in parts:
                dist.all_gather_into_tensor(whole, mine, group=group)
    else:
        for whole, mine in parts:
            dist.all_gather_into_tensor(whole, mine.clone() if not blob.is_cuda else mine, group=group)
