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


def _sharded_worker(rank, world, port, out_dir):
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path[:0] = [root, os.path.join(root, "oracle")]
    import plsvo_b200  # noqa: F401
    from plsvo_b200 import abi, dist as pd, synth
    import oracle_lib  # the CPU checker stands in for the CUDA path: this test covers the host logic only

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    data = None
    if rank == 0:
        data = synth.make_align_batch(cam=synth.QVGA, batch=5, n_pts=48, n_segs=10, max_level=3, min_level=1, seed=77, margin=32,
                                      motion_t=0.02, motion_r=0.006)
        data.pt_valid = (np.arange(5 * 48).reshape(5, 48) % 7 != 0).astype(np.uint8)
    full = pd.align_sharded(data, 3, 1, 30, src=0, run_fn=lambda shard: oracle_lib.align(abi, shard, abi.align_params(3, 1, 30)))
    np.savez(os.path.join(out_dir, f"sharded{rank}.npz"), **full)
    if rank == 0:
        whole = oracle_lib.align(abi, data, abi.align_params(3, 1, 30))
        np.savez(os.path.join(out_dir, "whole.npz"), **{f: getattr(whole, f) for f in ("T_cur_w", "n_tracked", "H", "seg_killed", "iters", "status")})
    dist.barrier()
    dist.destroy_process_group()


def test_align_sharded_scatter_run_gather_world2(tmp_path):
    """One host batch on rank 0 -> packed shards scattered -> every rank aligns its shard -> results gathered in batch
    order on every rank; equal to aligning the whole batch in one place (uneven shards: 5 pairs on 2 ranks)."""
    world = 2
    mp.spawn(_sharded_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    whole = np.load(tmp_path / "whole.npz")
    for r in range(world):
        got = np.load(tmp_path / f"sharded{r}.npz")
        for f in whole.files:
            np.testing.assert_array_equal(got[f], whole[f], err_msg=f"rank {r} {f}")
