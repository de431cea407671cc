"""CPU: the multi-GPU row partition (no data-path collective), exercised with a
2-rank gloo group; the oracle stands in for the kernels so only the host logic
is under test."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from llm_qat_b200.sharding import row_partition, shard_rows


@pytest.mark.parametrize("rows,world", [(8192, 8), (11008, 8), (7, 4), (3, 8), (0, 2), (2048, 1)])
def test_partition_covers_rows_exactly(rows, world):
    spans = [row_partition(rows, world, r) for r in range(world)]
    assert spans[0][0] == 0
    for (s0, n0), (s1, _) in zip(spans, spans[1:]):
        assert s0 + n0 == s1
    assert spans[-1][0] + spans[-1][1] == rows
    counts = [n for _, n in spans]
    assert max(counts) - min(counts) <= 1


def test_partition_rejects_bad_rank():
    with pytest.raises(ValueError):
        row_partition(10, 2, 2)
    with pytest.raises(ValueError):
        shard_rows(torch.zeros(4, 4), 2, 0, layerwise=True)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    import sys

    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from oracle import quant_oracle as qo

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(1234)
    x = torch.randn(37, 64, generator=g)                      # identical on every rank
    mine = shard_rows(x, world, rank)
    y_shard = torch.from_numpy(qo.sym_forward(mine.numpy(), 4)["y"])
    # rows are independent: concatenating the shards must equal the unsharded result
    sizes = [row_partition(37, world, r)[1] for r in range(world)]
    parts = [torch.empty(n, 64) for n in sizes]
    dist.all_gather(parts, y_shard) if len(set(sizes)) == 1 else None
    if len(set(sizes)) != 1:
        # ragged shards: gather through padded buffers
        pad = torch.zeros(max(sizes), 64)
        pad[: y_shard.shape[0]] = y_shard
        bufs = [torch.empty(max(sizes), 64) for _ in range(world)]
        dist.all_gather(bufs, pad)
        parts = [b[:n] for b, n in zip(bufs, sizes)]
    whole = torch.from_numpy(qo.sym_forward(x.numpy(), 4)["y"])
    ok = torch.equal(torch.cat(parts, 0), whole)
    # the timing contract: max over ranks of a per-rank scalar
    t = torch.tensor([float(rank + 1)])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ok = ok and t.item() == float(world)
    np.save(os.path.join(out_dir, f"ok{rank}.npy"), np.array([ok]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_shards_reassemble(tmp_path):
    world = 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert bool(np.load(tmp_path / f"ok{r}.npy")[0])
