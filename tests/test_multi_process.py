"""CPU, world_size 2 over gloo: the multi-GPU host logic (batch sharding + the single all-gather of
the pose tensor, incl. uneven shards) reproduces the unsharded batch."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, B, ok):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from talkshow_b200.pipeline import allgather_poses, shard_range

    g = torch.Generator().manual_seed(5)
    full = torch.rand(B, 6, 265, generator=g)             # what one GPU would produce for the whole batch
    lo, hi = shard_range(B, rank, world)
    out = allgather_poses(full[lo:hi].clone(), B, world)
    ok[rank] = int(torch.equal(out, full))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("B", [8, 5, 3])
def test_sharded_allgather_matches_unsharded(B):
    world = 2
    port = _free_port()
    ok = mp.get_context("spawn").Array("i", [0] * world)
    mp.spawn(_worker, args=(world, port, B, ok), nprocs=world, join=True)
    assert list(ok) == [1] * world
