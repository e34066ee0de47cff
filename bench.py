#!/usr/bin/env python
"""bench.py — SMPL-X motion frames/sec of the TalkSHOW generation path on B200.

    python bench.py --gpus N --steps K --warmup W            (N>1: launched by torch.distributed.run)
    python bench.py --impl reference ...                      (CPU arm: oracle port of the reference)

A step = one pass of the whole hot path over one batch of synthetic clips on every GPU:
face regressor (wav2vec2-based) + body path (audio encoder -> gated-PixelCNN sampler incl. its
Exp(1) noise draw -> two VQ-VAE decoders) + SMPL-X pose assembly, and for N>1 the single NCCL
all-gather of the [b,F,265] pose tensor.  Workload at every N: BASELINE config 5 per GPU (64 clips x
10 s x 4 speaker ids) — weak scaling, the batch shards with no data-path collective but the final
gather.  `value` has inputs resident in HBM; `e2e` goes through the public host-buffer call
(talkshow_b200.pipeline.WholeBody.generate_host) with H2D/D2H copies inside the timed region.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

FPS = 30
METRIC = "smplx_motion_frames_per_sec"


def env_int(k, d):
    return int(os.environ.get(k, d))


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,power.draw")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.stop = threading.Event()
        self.t = None

    def _run(self):
        while not self.stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                p = [x.strip() for x in out.strip().split(",")]
                if len(p) >= 6:
                    self.rows.append(p)
            except Exception:
                pass
            self.stop.wait(0.2)

    def __enter__(self):
        self.t = threading.Thread(target=self._run, daemon=True)
        self.t.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.t.join(timeout=6)

    def summary(self):
        sm = sorted(float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit())
        mx = [float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows for n, v in zip(names, r[2:6]) if v.lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(self.rows)}


def make_inputs(B, seconds, seed):
    """Synthetic 16 kHz clips + their MFCC features through the host front-end (feature extraction is
    outside the timed region on both arms)."""
    from talkshow_b200 import synth
    from talkshow_b200.data_utils.utils import mfcc_from_wave

    N = 16000 * seconds
    wave = synth.synth_wave(B, N, seed=seed)
    mf = [torch.from_numpy(mfcc_from_wave(wave[b:b + 1], 16000, sr=22000, fps=30).T.copy()) for b in range(B)]
    mfcc = torch.stack(mf, 0).contiguous()                     # [B,64,M]
    label = (torch.arange(B) % 4).to(torch.int64)
    return wave.contiguous(), mfcc, label


def cpu_reference_step(ck, mfcc, wave, label):
    """The reference's own CPU path, restated (oracle/): face Generator.forward + AudioEncoder +
    GatedPixelCNN.generate (literal O(T^2) loop, torch multinomial) + 2 VQ decoders + assembly."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import talkshow_oracle as O

    frame = wave.shape[1] * FPS // 16000
    face = O.face_forward(ck["face"]["generator"], wave, torch.zeros(wave.shape[0], 4), frame)
    _, body = O.body_generate(ck["pixel"], ck["vq"], mfcc, label, noise=None, window=None)
    out = [O.assemble_pose(face[b], body[b]) for b in range(wave.shape[0])]
    return torch.stack(out, 0)


def pick_threads(ck, B, seconds):
    """The reference's small convs do not scale to every core of a big host (oversubscription makes
    them slower): time one PixelCNN forward of the sample's shape at a few thread counts and keep
    the fastest — 'all the host threads it can use'."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import talkshow_oracle as O

    cores = os.cpu_count() or 1
    T = seconds * FPS // 4
    x = torch.zeros(B, T, 2, dtype=torch.int64)
    aud = torch.zeros(B, 256, T, 2)
    lab = torch.zeros(B, dtype=torch.int64)
    best = (None, 1e30)
    for n in sorted({cores, 64, 32, 16, 8, 4}, reverse=True):
        if n > cores:
            continue
        torch.set_num_threads(n)
        O.pixelcnn_forward(ck["pixel"]["generator"], x, lab, aud)
        t0 = time.perf_counter()
        O.pixelcnn_forward(ck["pixel"]["generator"], x, lab, aud)
        dt = time.perf_counter() - t0
        sys.stderr.write("[bench] cpu arm: %d threads -> %.3f s per sampler forward\n" % (n, dt))
        if dt < best[1]:
            best = (n, dt)
    torch.set_num_threads(best[0])
    return best[0]


def time_cpu(ck, B, seconds, steps, warmup):
    pick_threads(ck, B, seconds)
    wave, mfcc, label = make_inputs(B, seconds, 4321)
    torch.manual_seed(2024)
    ts = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        out = cpu_reference_step(ck, mfcc, wave, label)
        t1 = time.perf_counter()
        if i >= warmup:
            ts.append(t1 - t0)
    frames = B * out.shape[1]
    return frames, ts


def synthetic_ckpts():
    from talkshow_b200 import synth

    return {"pixel": synth.body_pixel_checkpoint(0), "vq": synth.body_vq_checkpoint(0), "face": synth.face_checkpoint(0)}


def run_reference(args, rank, world):
    if rank != 0:
        return
    torch.set_grad_enabled(False)
    ck = synthetic_ckpts()
    B = args.cpu_clips
    frames, ts = time_cpu(ck, B, args.seconds, args.steps, args.warmup)
    total = sum(ts)
    val = frames * len(ts) / total
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / len(ts), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, world, note="CPU arm runs a bounded sample of the same workload"),
        "cpu_baseline": {"value": val, "unit": "frames/s", "cores": torch.get_num_threads(), "kind": "port",
                         "sample": "%d clips x %d s (of the %d-clip batch), face+body+assembly, literal O(T^2) reference "
                                   "sampler loop, torch CPU fp32" % (B, args.seconds, args.batch)},
        "e2e": {"value": val, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def workload_config(args, world, note=None):
    c = {"workload": "BASELINE config 5 per GPU: %d clips x %d s x 4 speaker ids, face+body fused -> [B,%d,265]"
                     % (args.batch, args.seconds, args.seconds * FPS),
         "batch_per_gpu": args.batch, "global_batch": args.batch * world, "seconds": args.seconds,
         "frames_per_clip": args.seconds * FPS, "weights": "synthetic seed 0 (talkshow_b200/synth.py)",
         "parallelism": "dp%d: batch shard, one NCCL all-gather of the pose tensor" % world,
         "l2": "no explicit flush: per-step working set (0.68 GB weights + >8 GB activations) exceeds the 126 MB L2"}
    if note:
        c["note"] = note
    return c


def run_ours(args, rank, world, local_rank):
    import torch.distributed as dist

    from talkshow_b200 import build as tsbuild
    from talkshow_b200.engine import Engine
    from talkshow_b200.pipeline import WholeBody, allgather_poses

    torch.set_grad_enabled(False)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"          # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=dev)
    if rank == 0:
        tsbuild.build()
    if world > 1:
        dist.barrier()
    eng = Engine(local_rank)
    wb = WholeBody(eng)
    ck = synthetic_ckpts()
    wb.load(ck["pixel"], ck["vq"], ck["face"])
    B = args.batch
    wave_h, mfcc_h, label_h = make_inputs(B, args.seconds, 1234 + rank)
    wave_p, mfcc_p, label_p = wave_h.pin_memory(), mfcc_h.pin_memory(), label_h.pin_memory()
    wave_d, mfcc_d, label_d = wave_p.to(dev), mfcc_p.to(dev), label_p.to(dev)
    F = args.seconds * FPS
    T = eng.latent_rows(mfcc_h.shape[2])
    out_p = torch.empty(B, F, 265, pin_memory=True)
    torch.manual_seed(2024 + rank)

    def step_device():
        poses = wb.generate(mfcc_d, wave_d, label_d)
        return allgather_poses(poses, B * world, world)

    def step_host():
        poses = wb.generate(mfcc_p.to(dev, non_blocking=True), wave_p.to(dev, non_blocking=True),
                            label_p.to(dev, non_blocking=True))
        allp = allgather_poses(poses, B * world, world)
        out_p.copy_(allp[rank * B:(rank + 1) * B], non_blocking=True)     # each rank reads its own shard's result
        torch.cuda.current_stream().synchronize()

    def bracket():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        bracket()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        w0 = time.perf_counter()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        bracket()
        w1 = time.perf_counter()
        t = torch.tensor([e0.elapsed_time(e1), (w1 - w0) * 1e3], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t[0].item(), t[1].item()

    for _ in range(args.warmup):
        step_device()
    eng.pixelcnn_timing(True)
    l0 = eng.launches
    with ClockSampler(local_rank) as clk:
        dev_ms, _ = timed(step_device, args.steps)
        l1 = eng.launches
        pix_ms = [eng.pixelcnn_last_ms()]
        step_host()                                   # e2e warm-up (pinned staging buffers, allocator)
        _, e2e_ms = timed(step_host, args.steps)
    # a few more timed persistent-kernel launches for the roofline average (events on its launch stream)
    for _ in range(3):
        step_device()
        pix_ms.append(eng.pixelcnn_last_ms())
    eng.pixelcnn_timing(False)
    pix_ms = [x for x in pix_ms if x > 0]
    pix_avg = sum(pix_ms) / len(pix_ms)
    # dense-contraction side of the step: the face regressor alone (wav2vec2 CNN + transformer; tcgen05 3xTF32)
    idz = torch.zeros(B, 4, device=dev)
    face_ms = []
    for _ in range(3):
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        eng.face_forward(wave_d, idz, F)
        f1.record()
        torch.cuda.synchronize()
        face_ms.append(f0.elapsed_time(f1))
    face_avg = sum(face_ms[1:]) / len(face_ms[1:])

    frames_step = B * world * F
    value = frames_step * args.steps / (dev_ms / 1e3)
    e2e_val = frames_step * args.steps / (e2e_ms / 1e3)

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    alg_bytes = eng.pixelcnn_row_bytes * T                      # algorithmic weight bytes per launch
    achieved = alg_bytes / (pix_avg * 1e-3) / 1e9
    roofline = {"kernel": "pixelcnn_kernel<true,5> (persistent gated-PixelCNN sampler, fused 52-stage plan, %d rows/launch)" % T,
                "bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured)" if peaks else "fallback 6650 GB/s",
                "traffic": eng.pixelcnn_staged_row_bytes * T, "traffic_source": "bytes the kernel stages per launch (packed "
                "blob); ncu dram__bytes_read+write of the same kernel: 163.8 MB/row (profiles/r01h_pixelcnn_ncu_summary.md)", "launch_ms": pix_avg,
                "algorithmic_bytes": alg_bytes, "share_of_step": pix_avg / (dev_ms / args.steps)}

    # 106 GFLOP per 10 s clip (53 GMAC), scaled with the clip length
    face_flop = 106.0e9 * B * args.seconds / 10.0
    tf32x3_peak = float(peaks.get("bf16_tflops", 1590.0)) / 2.0 / 3.0     # tf32 rate = bf16/2, three products per MAC
    roofline_dense = {"kernel": "face path (tc2_gemm_kernel x56 tcgen05 3xTF32 + HMMA attention + FFMA2 layers per forward)", "bound": "tensor",
                      "achieved": face_flop / (face_avg * 1e-3) / 1e12, "peak": tf32x3_peak, "unit": "TFLOP/s",
                      "frac": face_flop / (face_avg * 1e-3) / 1e12 / tf32x3_peak,
                      "peak_source": "MEASURED_PEAKS.json bf16_tflops / 2 (tf32) / 3 (3xTF32 split)" if peaks else "fallback 1590/6",
                      "traffic": None, "launch_ms": face_avg, "share_of_step": face_avg / (dev_ms / args.steps)}
    line = {
        "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": workload_config(args, world),
        "e2e": {"value": e2e_val, "unit": "frames/s",
                "h2d_bytes_per_step": (wave_p.numel() * 4 + mfcc_p.numel() * 4 + label_p.numel() * 8) * world,
                "d2h_bytes_per_step": out_p.numel() * 4 * world, "ms_per_step": e2e_ms / args.steps},
        "gpu_launches": int(l1 - l0), "clocks": clk.summary(), "roofline": roofline, "roofline_dense": roofline_dense,
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sys.stderr.write("[bench] device legs done: value %.0f frames/s, e2e %.0f frames/s; timing the CPU arm sample\n" % (value, e2e_val))
        frames, ts = time_cpu(ck, args.cpu_clips, args.seconds, 1, 0)
        line["cpu_baseline"] = {"value": frames / ts[0], "unit": "frames/s", "cores": torch.get_num_threads(), "kind": "port",
                                "sample": "%d clips x %d s of the same workload, 1 run, oracle port of the reference "
                                          "(literal O(T^2) sampler), torch CPU fp32" % (args.cpu_clips, args.seconds)}
    if rank == 0:
        print(json.dumps(line))
    eng.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=64, help="clips per GPU (BASELINE config 5: 64)")
    ap.add_argument("--seconds", type=int, default=10)
    ap.add_argument("--cpu-clips", type=int, default=2, help="bounded sample size of the CPU arm")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank, world, local_rank = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
