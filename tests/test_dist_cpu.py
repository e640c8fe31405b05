"""world_size-2 gloo test of the N>1 path's host logic: contiguous sharding of a batch of independent
frame pairs and the pose gather (the GPUs run the same code over NCCL)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_items, out_dir):
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    import plsvo_b200  # noqa: F401
    from plsvo_b200 import dist as pd

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    b, e = pd.shard_range(n_items, rank, world)
    idx = np.arange(b, e, dtype=np.float64)
    local = np.stack([idx * 10 + k for k in range(7)], axis=1)  # stand-in for [n_local, 7] poses
    full = pd.gather_rows(local, n_items)
    np.save(os.path.join(out_dir, f"rank{rank}.npy"), full)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_items", [8, 5, 1])
def test_shard_and_gather_world2(tmp_path, n_items):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), n_items, str(tmp_path)), nprocs=world, join=True)
    idx = np.arange(n_items, dtype=np.float64)
    expect = np.stack([idx * 10 + k for k in range(7)], axis=1)
    for r in range(world):
        np.testing.assert_array_equal(np.load(tmp_path / f"rank{r}.npy"), expect)


def test_shard_range_partitions_exactly():
    import plsvo_b200  # noqa: F401
    from plsvo_b200 import dist as pd

    for n in (0, 1, 7, 8, 1024, 1025):
        for world in (1, 2, 3, 4, 8):
            ranges = [pd.shard_range(n, r, world) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == n
            assert all(ranges[i][1] == ranges[i + 1][0] for i in range(world - 1))
            sizes = [e - b for b, e in ranges]
            assert max(sizes) - min(sizes) <= 1
