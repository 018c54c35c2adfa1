"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: shard partitioning and the
unequal-shard gather used after epoch-sharded synthesis / row-block-sharded covariance
propagation."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def test_shard_range_partitions_exactly():
    from grates_b200.distributed import shard_range, shard_counts
    for count in (0, 1, 7, 240, 360, 500):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(count, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == count
            for (a0, a1), (b0, b1) in zip(spans[:-1], spans[1:]):
                assert a1 == b0 and a1 >= a0
            sizes = shard_counts(count, world)
            assert sum(sizes) == count and max(sizes) - min(sizes) <= 1
    assert shard_range(240, 8, 3) == (90, 120)         # config 2: 30 epochs per GPU
    assert shard_range(500, 8, 7) == (438, 500)        # config 5: 62/63 epochs
    assert shard_range(360, 8, 0) == (0, 45)           # config 4: 45 parallels per GPU
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from grates_b200.distributed import gather_shards, shard_counts, shard_range
    total = 7                                            # uneven: 4 + 3
    full = torch.arange(total * 6, dtype=torch.float64).reshape(total, 2, 3)
    a, b = shard_range(total, world, rank)
    local = full[a:b].clone() * 1.0
    gathered = gather_shards(local, shard_counts(total, world))
    ok = torch.equal(gathered, full)
    # row-block variance gather: every rank contributes its block of "parallels"
    rows = 5
    a, b = shard_range(rows, world, rank)
    block = torch.full((b - a, 4), float(rank), dtype=torch.float64)
    std = gather_shards(block, shard_counts(rows, world)).reshape(-1)
    ok = ok and std.shape[0] == rows * 4 and float(std[0]) == 0.0 and float(std[-1]) == float(world - 1)
    # covariance row blocks: northern shard + mirror image per rank (+ the equator on the last rank), gathered rank-major
    # and put back into the order of the parallels
    from grates_b200.distributed import covariance_row_blocks, covariance_row_counts, reorder_row_blocks
    for nlat in (9, 12):
        mine = []
        for kind, start, count in covariance_row_blocks(nlat, world, rank):
            mine += list(range(start, start + count))
            if kind == "mirrored":
                mine += list(range(nlat - start - count, nlat - start))
        block = torch.tensor(mine, dtype=torch.float64).reshape(-1, 1).repeat(1, 3)
        rows_back = reorder_row_blocks(gather_shards(block, covariance_row_counts(nlat, world)), nlat, world)
        ok = ok and torch.equal(rows_back[:, 0], torch.arange(nlat, dtype=torch.float64))
    # max-over-ranks timing reduction as used by bench.py
    t = torch.tensor([1.0 + rank], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ok = ok and float(t) == float(world)
    np.save(os.path.join(out_dir, "rank%d.npy" % rank), np.array([int(ok)]))
    dist.destroy_process_group()


def test_gather_shards_gloo_world2(tmp_path):
    world = 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert np.load(tmp_path / ("rank%d.npy" % r))[0] == 1
