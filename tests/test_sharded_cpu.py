"""N>1 host logic on CPU: world_size-2 (and 3) `gloo` processes shard an item set, compute per-item rows and
gather them; the result must equal the single-process result exactly, for even and ragged splits."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from spegnet_b200 import sharded


def _row_fn(indices):
    idx = torch.tensor(indices, dtype=torch.float64)
    return torch.stack([torch.sin(idx) * 0.5 + 0.5, idx * idx, (idx % 7) / 7.0], dim=1) if len(indices) else torch.zeros(0, 3, dtype=torch.float64)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_items, batch, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rows = sharded.sharded_map(n_items, batch, _row_fn)
        torch.save({"rows": rows, "mean": sharded.mean_in_index_order(rows)}, os.path.join(out_dir, f"r{rank}.pt"))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n_items,batch", [(2, 2026, 64), (2, 7, 3), (3, 10, 4), (2, 1, 8)])
def test_sharded_gather_matches_single_process(tmp_path, world, n_items, batch):
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n_items, batch, str(tmp_path)), nprocs=world, join=True)
    expect = _row_fn(list(range(n_items)))
    for r in range(world):
        got = torch.load(os.path.join(tmp_path, f"r{r}.pt"))
        assert torch.equal(got["rows"], expect)  # identical rows, in index order, on every rank
        assert torch.equal(got["mean"], sharded.mean_in_index_order(expect))


def test_shard_indices_partition():
    for world in (1, 2, 4, 8):
        seen = sorted(i for r in range(world) for i in sharded.shard_indices(2026, r, world))
        assert seen == list(range(2026))
        sizes = [len(sharded.shard_indices(2026, r, world)) for r in range(world)]
        assert max(sizes) - min(sizes) <= 1 and max(sizes) == sharded.padded_shard_size(2026, world)
    with pytest.raises(ValueError):
        sharded.shard_indices(10, 2, 2)


def test_single_process_path():
    rows = sharded.sharded_map(5, 2, _row_fn)
    assert torch.equal(rows, _row_fn([0, 1, 2, 3, 4]))
