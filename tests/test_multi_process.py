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


class _StubEngine:
    """host-only stand-in for talkshow_b200.engine.Engine: enough surface for WholeBody.generate_sharded."""
    device = torch.device("cpu")

    @staticmethod
    def latent_rows(M):
        m = (M + 2 - 4) // 2 + 1
        return (m + 2 - 4) // 2 + 1


def _stub_wholebody():
    from talkshow_b200.pipeline import WholeBody

    class Stub(WholeBody):
        def generate(self, mfcc, wave, label, noise=None, stand=False, per_step_noise=True):
            # a per-sample function of every per-sample input (incl. the sample's own noise column)
            B = mfcc.shape[0]
            frame = wave.shape[1] * 30 // 16000
            key = mfcc.sum((1, 2)) + wave.sum(1) + label.float() + noise.sum((0, 2))
            return key.view(B, 1, 1).expand(B, frame, 265).contiguous()

    return Stub(_StubEngine())


def _worker_sharded(rank, world, port, B, ok):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    wb = _stub_wholebody()
    g = torch.Generator().manual_seed(11)
    mfcc, wave, label = torch.rand(B, 64, 16, generator=g), torch.rand(B, 1600, generator=g), torch.arange(B) % 4
    T = wb.e.latent_rows(16)
    noise = torch.rand(2 * T, B, 2048, generator=g) + 0.1
    full = wb.generate(mfcc, wave, label, noise=noise)                 # what ONE device computes for the whole batch
    out = wb.generate_sharded(mfcc, wave, label, rank, world, noise_full=noise)
    # seed path: every rank draws the same full-batch stream and keeps its slice
    a = wb.generate_sharded(mfcc, wave, label, rank, world, seed=77)
    b = wb.generate_sharded(mfcc, wave, label, 0, 1, seed=77)
    ok[rank] = int(torch.equal(out, full) and torch.equal(a, b))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("B", [12, 5, 1])
def test_generate_sharded_is_world_independent(B):
    """full-batch noise drawn in reference order, sliced per rank -> the gathered result equals the
    one-device result for even, uneven and empty shards (BASELINE config 4's 12 samples; 5; 1 sample on 2 ranks)."""
    world = 2
    port = _free_port()
    ok = mp.get_context("spawn").Array("i", [0] * world)
    mp.spawn(_worker_sharded, args=(world, port, B, ok), nprocs=world, join=True)
    assert list(ok) == [1] * world
