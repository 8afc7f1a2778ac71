"""Multi-process path on CPU: world_size-2 gloo run of the image-index sharding and the one all-gather of detections."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def test_shard_indices_partition():
    from fsd_b200.shard import owner_of, shard_indices

    for world in (1, 2, 4, 8):
        parts = [shard_indices(37, r, world) for r in range(world)]
        assert sorted(i for p in parts for i in p) == list(range(37))
        assert all(owner_of(i, world) == r for r, p in enumerate(parts) for i in p)
        assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
    with pytest.raises(ValueError):
        shard_indices(4, 2, 2)


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from fsd_b200.shard import gather_detections, shard_indices

    mine = shard_indices(11, rank, world)
    rng = np.random.default_rng(0)
    per_image = {i: rng.random((int(rng.integers(0, 5)), 6)).astype(np.float32) + i for i in range(11)}  # same on all ranks
    ids = torch.tensor([i for i in mine for _ in range(len(per_image[i]))], dtype=torch.int64)
    rows = torch.from_numpy(np.concatenate([per_image[i] for i in mine] + [np.zeros((0, 6), np.float32)], 0))
    gi, gr = gather_detections(ids, rows)
    torch.save((gi, gr), os.path.join(out_dir, f"r{rank}.pt"))
    dist.destroy_process_group()


def test_gather_detections_world2_gloo(tmp_path):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    g0, r0 = torch.load(tmp_path / "r0.pt")
    g1, r1 = torch.load(tmp_path / "r1.pt")
    assert torch.equal(g0, g1) and torch.equal(r0, r1)           # identical on every rank
    rng = np.random.default_rng(0)
    per_image = {i: rng.random((int(rng.integers(0, 5)), 6)).astype(np.float32) + i for i in range(11)}
    want = np.concatenate([per_image[i] for i in range(11)], 0)   # == the single-process result, image order
    assert np.array_equal(r0.numpy(), want)
    assert g0.tolist() == [i for i in range(11) for _ in range(len(per_image[i]))]


def test_gather_detections_single_process():
    from fsd_b200.shard import gather_detections

    ids = torch.tensor([3, 1, 3, 0])
    rows = torch.arange(8, dtype=torch.float32).view(4, 2)
    gi, gr = gather_detections(ids, rows)
    assert gi.tolist() == [0, 1, 3, 3] and gr[:, 0].tolist() == [6.0, 2.0, 0.0, 4.0]
