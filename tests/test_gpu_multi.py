"""GPU (-m gpu), needs >= 2 devices (skipped otherwise; run with `gpurun --gpus 2`): the sharded whole-body path over
real NCCL — the library's own ts_allgather — reproduces the one-GPU result bit for bit."""
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ok):
    import torch.distributed as dist

    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    from talkshow_b200 import synth
    from talkshow_b200.engine import Engine
    from talkshow_b200.pipeline import WholeBody

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.set_grad_enabled(False)
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)          # bootstrap only: the data path is ts_allgather
    e = Engine(rank).nccl_init(rank, world)
    wb = WholeBody(e)
    wb.load(synth.body_pixel_checkpoint(0), synth.body_vq_checkpoint(0), synth.face_checkpoint(0))
    B, M = 5, 60                                                           # uneven shards: 3 + 2
    mfcc, wave = synth.synth_mfcc(B, M, seed=701), synth.synth_wave(B, 16000 * 2, seed=702)
    label = torch.arange(B) % 4
    full_noise = torch.empty(2 * e.latent_rows(M), B, 2048).exponential_(1, generator=torch.Generator().manual_seed(703))
    got = wb.generate_sharded(mfcc, wave, label, rank, world, noise_full=full_noise.cuda(rank))
    ref = wb.generate(mfcc.cuda(rank), wave.cuda(rank), label.cuda(rank), noise=full_noise.cuda(rank))   # whole batch on this GPU
    body = list(range(3, 165))
    same = torch.equal(got[:, :, body], ref[:, :, body]) and (got - ref).abs().max().item() <= 1e-5
    ok[rank] = int(same and got.shape == (B, 60, 265))
    torch.cuda.synchronize()
    e.close()
    dist.barrier()
    dist.destroy_process_group()


def test_two_gpu_sharded_equals_single_gpu():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    world = 2
    ok = mp.get_context("spawn").Array("i", [0] * world)
    mp.spawn(_worker, args=(world, _free_port(), ok), nprocs=world, join=True)
    assert list(ok) == [1] * world
