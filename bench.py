#!/usr/bin/env python
"""bench.py — SMPL-X motion frames/sec of the TalkSHOW generation path on B200.

    python bench.py --gpus N --steps K --warmup W            (N>1: launched by torch.distributed.run)
    python bench.py --impl reference ...                      (CPU arm: oracle port of the reference)

A step = one pass of the whole hot path over one batch of synthetic clips: face regressor
(wav2vec2-based) + body path (audio encoder -> gated-PixelCNN sampler incl. its Exp(1) noise draw ->
two VQ-VAE decoders) + SMPL-X pose assembly, and for N>1 the single NCCL all-gather of the
[b,F,265] pose tensor.

Workload (default, `--scaling strong`): BASELINE.json config 5 — ONE global batch of 64 clips x 10 s
x 4 speaker ids, sharded 64 / 32 / 16 / 8 clips per GPU at N = 1 / 2 / 4 / 8: the sampler
noise is drawn for the full batch on every rank (same seed, same stream) and sliced, so the gathered
result does not depend on N.  `value` has inputs resident in HBM; `e2e` goes through host buffers with
the H2D / D2H copies inside the timed region.  The same JSON line also carries `weak` (64 clips per
GPU), `config4` (12 diversity samples x 10 s sharded over the ranks) and `config3` (1 clip x 4 s).
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
CPU_THREADS = 32     # thread count of the CPU arm (see cpu_threads())
# DRAM bytes per latent row of the persistent PixelCNN kernel by batch tile, from `ncu --set full` captures of the kernel alone
# (profiles/r02_pixelcnn_ncu_summary.md); the algorithmic figure is 89.8 MB per row
NCU_DRAM_BYTES_PER_ROW = {64: 163.6e6, 32: 158.0e6, 16: 153.0e6}


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


def make_inputs(B, seconds, seed, same_clip=False):
    """Synthetic 16 kHz clips + their MFCC features through the host front-end (feature extraction is
    outside the timed region on both arms).  same_clip: one clip repeated (diversity samples)."""
    from talkshow_b200 import synth
    from talkshow_b200.data_utils.utils import mfcc_from_wave

    N = 16000 * seconds
    nb = 1 if same_clip else B
    wave = synth.synth_wave(nb, N, seed=seed)
    mf = [torch.from_numpy(mfcc_from_wave(wave[b:b + 1], 16000, sr=22000, fps=30).T.copy()) for b in range(nb)]
    mfcc = torch.stack(mf, 0).contiguous()                     # [B,64,M]
    if same_clip:
        wave, mfcc = wave.repeat(B, 1), mfcc.repeat(B, 1, 1)
        label = torch.zeros(B, dtype=torch.int64)
    else:
        label = (torch.arange(B) % 4).to(torch.int64)
    return wave.contiguous(), mfcc.contiguous(), label


# ---------------------------------------------------------------------------------------------------------------
# CPU arm: the reference's own CPU path, restated (oracle/)
# ---------------------------------------------------------------------------------------------------------------
def cpu_reference_step(ck, mfcc, wave, label):
    """face Generator.forward + AudioEncoder + GatedPixelCNN.generate (literal O(T^2) loop, torch multinomial) +
    2 VQ decoders + assembly."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import talkshow_oracle as O

    frame = wave.shape[1] * FPS // 16000
    face = O.face_forward(ck["face"]["generator"], wave, torch.zeros(wave.shape[0], 4), frame)
    _, body = O.body_generate(ck["pixel"], ck["vq"], mfcc, label, noise=None, window=None)
    out = [O.assemble_pose(face[b], body[b]) for b in range(wave.shape[0])]
    return torch.stack(out, 0)


def cpu_threads():
    """Fixed thread count of the CPU arm: min(32, host cores).  The reference's small convolutions do not scale past
    one NUMA node of the pool's hosts — 128 threads measured 45x slower than 32 in round 1, and a per-run calibration
    made the arm's numbers jump between 16 and 128 threads — so the count is pinned and stated."""
    n = min(CPU_THREADS, os.cpu_count() or 1)
    torch.set_num_threads(n)
    # denormal operands make the host's conv kernels 60x slower on these synthetic checkpoints (measured: 2.98 s vs 0.047 s
    # for one sampler forward) and the arm's numbers erratic; flush-to-zero is the configuration a CPU user would run
    torch.set_flush_denormal(True)
    return n


def time_cpu(ck, B, seconds, steps, warmup):
    cpu_threads()
    wave, mfcc, label = make_inputs(B, seconds, 4321)
    torch.manual_seed(2024)
    ts = []
    out = None
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        out = cpu_reference_step(ck, mfcc, wave, label)
        t1 = time.perf_counter()
        if i >= warmup:
            ts.append(t1 - t0)
    return B * out.shape[1], ts


def synthetic_ckpts():
    from talkshow_b200 import synth

    return {"pixel": synth.body_pixel_checkpoint(0), "vq": synth.body_vq_checkpoint(0), "face": synth.face_checkpoint(0)}


def cpu_sample_text(B, seconds, nsteps):
    return ("%d clips x %d s of the same workload (face+body+assembly), literal O(T^2) reference sampler loop with "
            "torch.multinomial, torch CPU fp32 with flush-denormal on, %d threads, %d timed step(s), median" % (B, seconds, torch.get_num_threads(), nsteps))


def run_reference(args, rank, world):
    if rank != 0:
        return
    torch.set_grad_enabled(False)
    ck = synthetic_ckpts()
    B = args.cpu_clips
    steps, warmup = max(1, args.steps), min(args.warmup, 1)      # one warm-up pass is enough for the CPU allocator
    frames, ts = time_cpu(ck, B, args.seconds, steps, warmup)
    med = sorted(ts)[len(ts) // 2]
    val = frames / med
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "frames/s", "n_gpus": args.gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": 1e3 * med, "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, world, note="CPU arm: every step is a bounded sample of the workload (%d of the %d clips); "
                                  "frames/s does not depend on how many clips are timed (clips are independent, the arm is "
                                  "compute-bound per clip)" % (B, args.batch)),
        "cpu_baseline": {"value": val, "unit": "frames/s", "cores": torch.get_num_threads(), "kind": "port",
                         "sample": cpu_sample_text(B, args.seconds, steps)},
        "e2e": {"value": val, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "step_times_s": [round(t, 3) for t in ts],
    }
    print(json.dumps(line))


def workload_config(args, world, note=None):
    F = args.seconds * FPS
    if args.scaling == "strong":
        per = [(args.batch // world) + (1 if r < args.batch % world else 0) for r in range(world)]
        c = {"workload": "BASELINE config 5: global batch of %d clips x %d s x 4 speaker ids, face+body fused -> [%d,%d,265], "
                         "sharded over %d GPU(s)" % (args.batch, args.seconds, args.batch, F, world),
             "global_batch": args.batch, "batch_per_gpu": per, "parallelism": "dp%d: contiguous batch shards, full-batch sampler "
             "noise sliced per rank (result independent of N), one NCCL all-gather of the pose tensor" % world}
    else:
        c = {"workload": "BASELINE config 5 per GPU: %d clips x %d s x 4 speaker ids, face+body fused" % (args.batch, args.seconds),
             "global_batch": args.batch * world, "batch_per_gpu": args.batch,
             "parallelism": "dp%d: batch shard, one NCCL all-gather of the pose tensor" % world}
    c.update({"seconds": args.seconds, "frames_per_clip": F, "weights": "synthetic seed 0 (talkshow_b200/synth.py)",
              "sampler_noise": "one batched exponential_ draw of [2T,B,2048] per step inside the timed region",
              "l2": "no explicit flush: per-step working set (0.68 GB weights + GBs of activations) exceeds the 126 MB L2"})
    if note:
        c["note"] = note
    return c


# ---------------------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------------------
def run_ours(args, rank, world, local_rank):
    import torch.distributed as dist

    from talkshow_b200 import build as tsbuild
    from talkshow_b200.engine import Engine
    from talkshow_b200.pipeline import WholeBody, allgather_poses, shard_range

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
    if world > 1:
        eng.nccl_init(rank, world)                    # the pose all-gather runs inside the library (ts_allgather)
    wb = WholeBody(eng, overlap_batch=int(os.environ.get("TS_OVERLAP_BATCH", "64")), overlap_ctas=int(os.environ.get("TS_OVERLAP_CTAS", "96")))
    ck = synthetic_ckpts()
    wb.load(ck["pixel"], ck["vq"], ck["face"])

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

    class Workload:
        """One sharded batch: rank-local device inputs, pinned host copies, the step functions."""

        def __init__(self, Bg, seconds, seed, sliced_noise, same_clip=False):
            self.Bg, self.seconds = Bg, seconds
            self.lo, self.hi = shard_range(Bg, rank, world)
            self.b = self.hi - self.lo
            wave_h, mfcc_h, label_h = make_inputs(Bg, seconds, seed, same_clip)
            sl = slice(self.lo, self.hi)
            self.host = [t[sl].contiguous().pin_memory() for t in (mfcc_h, wave_h, label_h)]
            self.devt = [t.to(dev) for t in self.host]
            self.F = seconds * FPS
            self.T = eng.latent_rows(mfcc_h.shape[2])
            self.sliced = sliced_noise
            self.out_p = torch.empty(max(self.b, 1), self.F, 265).pin_memory()
            self.h2d = sum(t.numel() * t.element_size() for t in self.host)

        def noise(self):
            if self.sliced:      # full-batch draw, every rank the same stream (same seed): result independent of N
                full = torch.empty(2 * self.T, self.Bg, 2048, device=dev).exponential_(1)
                return full[:, self.lo:self.hi].contiguous()
            return torch.empty(2 * self.T, self.b, 2048, device=dev).exponential_(1)

        def run(self, inputs):
            if self.b == 0:
                local = torch.empty(0, self.F, 265, device=dev)
                self.noise()                         # keep the generator streams of the ranks in step
            else:
                local = wb.generate(inputs[0], inputs[1], inputs[2], noise=self.noise())
            return allgather_poses(local, self.Bg, world, engine=eng)

        def step_device(self):
            return self.run(self.devt)

        def step_host(self):
            allp = self.run([t.to(dev, non_blocking=True) for t in self.host])
            if self.b:
                self.out_p.copy_(allp[self.lo:self.hi], non_blocking=True)     # each rank reads its own shard's result
            torch.cuda.current_stream().synchronize()

        def measure(self, steps, warmup, e2e=True):
            torch.manual_seed(2024)               # same generator state on every rank
            for _ in range(warmup):
                self.step_device()
            dev_ms, _ = timed(self.step_device, steps)
            res = {"value": self.Bg * self.F * steps / (dev_ms / 1e3), "ms_per_step": dev_ms / steps}
            if e2e:
                self.step_host()
                _, e2e_ms = timed(self.step_host, steps)
                res["e2e"] = {"value": self.Bg * self.F * steps / (e2e_ms / 1e3), "unit": "frames/s",
                              "h2d_bytes_per_step": self.h2d_total(), "d2h_bytes_per_step": self.Bg * self.F * 265 * 4,
                              "ms_per_step": e2e_ms / steps}
            return res

        def h2d_total(self):
            t = torch.tensor([float(self.h2d)], device=dev, dtype=torch.float64)
            if world > 1:
                dist.all_reduce(t)
            return int(t.item())

    strong = args.scaling == "strong"
    Bg = args.batch if strong else args.batch * world
    main = Workload(Bg, args.seconds, 1234, sliced_noise=strong)
    wb.pixelcnn_timing(True)
    with ClockSampler(local_rank) as clk:
        for _ in range(args.warmup):
            main.step_device()
        l0 = wb.launches
        torch.manual_seed(2024)
        dev_ms, _ = timed(main.step_device, args.steps)
        l1 = wb.launches
        pix_in_step = [wb.pixelcnn_last_ms()]
        main.step_host()                               # e2e warm-up (pinned staging buffers, allocator)
        _, e2e_ms = timed(main.step_host, args.steps)
    # a few more timed sampler launches (events on its launch stream, inside the library): in the step as it runs (side by
    # side with the face path when the batch is overlapped) ...
    for _ in range(3):
        main.step_device()
        pix_in_step.append(wb.pixelcnn_last_ms())
    pix_in_step = [x for x in pix_in_step if x > 0]
    pix_step_avg = sum(pix_in_step) / max(1, len(pix_in_step))
    # ... and ALONE on the whole GPU (sequential order, every SM): the roofline figure of the kernel itself
    ob, wb.overlap_batch = wb.overlap_batch, 0
    pix_ms = []
    for i in range(4):
        main.step_device()
        if i:
            pix_ms.append(wb.pixelcnn_last_ms())
    wb.overlap_batch = ob
    pix_ms = [x for x in pix_ms if x > 0]
    pix_avg = sum(pix_ms) / max(1, len(pix_ms))
    # dense-contraction side of the step: the face regressor alone (wav2vec2 CNN + transformer; tcgen05 3xTF32)
    face_avg = 0.0
    if main.b:
        idz = torch.zeros(main.b, 4, device=dev)
        face_ms = []
        for _ in range(3):
            f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            f0.record()
            eng.face_forward(main.devt[1], idz, main.F)
            f1.record()
            torch.cuda.synchronize()
            face_ms.append(f0.elapsed_time(f1))
        face_avg = sum(face_ms[1:]) / len(face_ms[1:])

    frames_step = Bg * main.F
    value = frames_step * args.steps / (dev_ms / 1e3)
    e2e_val = frames_step * args.steps / (e2e_ms / 1e3)

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    T = main.T
    alg_bytes = eng.pixelcnn_row_bytes * T                      # algorithmic weight bytes per launch
    achieved = alg_bytes / (pix_avg * 1e-3) / 1e9 if pix_avg > 0 else 0.0
    tile = 16 if main.b <= 16 else 32 if main.b <= 32 else 64
    overlapped = wb.e2 is not None and 0 < main.b <= wb.overlap_batch
    roofline = {"kernel": "pixelcnn_kernel<persistent, batch tile %d> (gated-PixelCNN sampler, fused 52-stage plan, %d rows/launch, "
                          "%d samples on this GPU%s)" % (tile, T, main.b, "; launch_ms = alone on %d SMs, launch_ms_in_step = on %d SMs next to the "
                                                         "face path (two streams)" % (eng.sm_count, wb.overlap_ctas) if overlapped else ""),
                "bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured)" if peaks else "fallback 6650 GB/s",
                "traffic": NCU_DRAM_BYTES_PER_ROW[tile] * T, "traffic_source": "ncu --set full dram__bytes_read.sum + dram__bytes_write.sum of this "
                "kernel per latent row (profiles/r02_pixelcnn_ncu_summary.md: 163.6 MB at the 64-sample tile, 153.0 MB at the 16-sample tile; 32: "
                "between, not captured) x %d rows; staged bytes per launch: %d" % (T, eng.pixelcnn_staged_row_bytes * T), "launch_ms": pix_avg,
                "launch_ms_in_step": pix_step_avg, "algorithmic_bytes": alg_bytes, "share_of_step": pix_step_avg / (dev_ms / args.steps)}
    # 106 GFLOP per 10 s clip (53 GMAC), scaled with the clip length
    face_flop = 106.0e9 * main.b * args.seconds / 10.0
    # fp32-grade MACs as three kind::f16 products on fp16-split operands (the default since r02; 3xTF32 would be bf16/2/3)
    tf32x3_peak = float(peaks.get("bf16_tflops", 1590.0)) / 3.0
    roofline_dense = None
    if face_avg > 0:
        roofline_dense = {"kernel": "face path (tc2_gemm_kernel<fp16-split> tcgen05 kind::f16 x3 + HMMA attention + FFMA2 layers per forward)",
                          "bound": "tensor", "achieved": face_flop / (face_avg * 1e-3) / 1e12, "peak": tf32x3_peak, "unit": "TFLOP/s",
                          "frac": face_flop / (face_avg * 1e-3) / 1e12 / tf32x3_peak,
                          "peak_source": "MEASURED_PEAKS.json bf16_tflops / 3 (three f16 products per fp32-grade MAC)" if peaks else "fallback 1590/3",
                          "traffic": None, "launch_ms": face_avg, "share_of_step": face_avg / (dev_ms / args.steps)}
    line = {
        "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": workload_config(args, world),
        "e2e": {"value": e2e_val, "unit": "frames/s", "h2d_bytes_per_step": main.h2d_total(),
                "d2h_bytes_per_step": Bg * main.F * 265 * 4, "ms_per_step": e2e_ms / args.steps},
        "gpu_launches": int(l1 - l0), "clocks": clk.summary(), "roofline": roofline, "roofline_dense": roofline_dense,
    }
    if not args.no_extras:
        ex_steps, ex_warm = max(3, args.steps), 3
        # the other scaling mode of config 5, config 4 (12 diversity samples x 10 s, sharded) and config 3 (1 clip x 4 s, rank 0's GPU)
        other = Workload(args.batch * world if strong else args.batch, args.seconds, 1234, sliced_noise=not strong)
        if world > 1 or not strong:
            r = other.measure(ex_steps, ex_warm, e2e=False)
        else:
            r = {"value": value, "ms_per_step": dev_ms / args.steps}       # N = 1: both modes are the same workload
        line["weak" if strong else "strong"] = dict(r, unit="frames/s", global_batch=other.Bg,
                                                    note="config 5, %s scaling" % ("weak: %d clips per GPU" % args.batch if strong else "strong"))
        c4 = Workload(12, args.seconds, 77, sliced_noise=True, same_clip=True)
        r4 = c4.measure(ex_steps, ex_warm)
        pix4 = wb.pixelcnn_last_ms()
        line["config4"] = dict(r4, unit="frames/s", workload="BASELINE config 4: 12 diversity samples x %d s, id 0, shards %s"
                               % (args.seconds, [shard_range(12, k, world)[1] - shard_range(12, k, world)[0] for k in range(world)]),
                               roofline={"bound": "hbm", "launch_ms": pix4, "frac": (alg_bytes / (pix4 * 1e-3) / 1e9 / hbm_peak) if pix4 > 0 else None,
                                         "samples_on_rank0": c4.b})
        c3 = Workload(1, 4, 5, sliced_noise=True)
        r3 = c3.measure(ex_steps, ex_warm)
        pix3 = wb.pixelcnn_last_ms() if c3.b else -1.0
        alg3 = eng.pixelcnn_row_bytes * c3.T
        line["config3"] = dict(r3, unit="frames/s", workload="BASELINE config 3: 1 clip x 4 s, id 0 (runs on rank 0's GPU)",
                               roofline={"bound": "hbm", "launch_ms": pix3, "frac": (alg3 / (pix3 * 1e-3) / 1e9 / hbm_peak) if pix3 > 0 else None})
    wb.pixelcnn_timing(False)
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sys.stderr.write("[bench] device legs done: value %.0f frames/s, e2e %.0f frames/s; timing the CPU arm sample\n" % (value, e2e_val))
        frames, ts = time_cpu(ck, args.cpu_clips, args.seconds, 3, 0)          # ~15 s per step on 32 threads; median of 3
        line["cpu_baseline"] = {"value": frames / sorted(ts)[1], "unit": "frames/s", "cores": torch.get_num_threads(), "kind": "port",
                                "sample": cpu_sample_text(args.cpu_clips, args.seconds, 3), "step_times_s": [round(t, 3) for t in ts]}
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
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"],
                    help="strong: --batch is the GLOBAL batch sharded over the GPUs (BASELINE config 5); weak: --batch clips per GPU")
    ap.add_argument("--batch", type=int, default=64, help="BASELINE config 5: 64 clips")
    ap.add_argument("--seconds", type=int, default=10)
    ap.add_argument("--cpu-clips", type=int, default=8, help="bounded sample size of the CPU arm (clips per step)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the weak / config 4 / config 3 lines")
    args = ap.parse_args()
    rank, world, local_rank = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
