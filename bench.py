#!/usr/bin/env python
"""Benchmark of the fused colour + block-DCT + deadzone path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Main workload (BASELINE.json configs[1]): batches of synthetic 3840x2160 RGB frames,
YCoCg, B=8, deadzone q cycling through {8,16,32,64}.  One *step* = one pass of the
hot path over one batch: encode (uint8 RGB -> uint8 indices) followed by decode
(indices -> uint8 RGB).  Metric: Mpixel/s of encode+decode, whole job.

Keys of the JSON line:

* ``value``    -- inputs resident in HBM, CUDA-event timed on the launching stream, natural-like
  frames.  Mode of the headline: the north star's operating mode -- encoder bit-exact with the
  reference's float32 path, float32 decoder (pixels within +-1 LSB of the reference, PSNR
  within 0.01 dB).  ``exact_mode`` is the same step with the reference's float64 decode chain
  reproduced bit for bit; ``noise`` repeats both on i.i.d. uniform frames.
* ``e2e``      -- the same step through the public host API (numpy in pinned host memory in,
  numpy out): host->device and device->host copies inside the timed region;
  ``host_copy_ceiling`` is the bare copy loop over the same buffers, ``e2e_codestream`` the
  encode leg with the entropy front-end on the GPU (only the code-stream crosses PCIe).
* ``roofline`` -- dominant kernel against the measured HBM copy peak.
* ``workloads`` -- BASELINE configs[3], [4] and the RD-statistics variant of configs[1], each
  with its own timed region at this N; ``rde`` and ``c5`` contain the design's only collective
  (one NCCL all-reduce of int64[776] per batch) inside the timed region.
* ``cpu_baseline`` -- the CPU oracle (numpy/scipy restatement of the reference) timed on this
  box's host cores on a bounded sample of the same frames.

Frame-parallel at N GPUs (one process per GPU, torchrun): every rank transforms its own batch,
no data-path collective (weak scaling; ``workloads.c4`` is the one strong-scaling case: 1024
frames in total).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# workload -> (H, W, B, qs, colour, statistics + all-reduce, frames per GPU per step, description)
WORKLOADS = {
    "c2": (2160, 3840, 8, (8, 16, 32, 64), "YCoCg", False, 64,
           "configs[1]: 3840x2160 RGB, YCoCg, B=8, q in {8,16,32,64} cycled per step, encode+decode"),
    "rde": (2160, 3840, 8, (8, 16, 32, 64), "YCoCg", True, 64,
            "configs[1] frames + RD statistics (SSE, non-zero count, sum|k|) + NCCL all-reduce per batch"),
    "c4": (1080, 1920, 8, (32,), "YCoCg", False, 1024,
           "configs[3]: 1024 1920x1080 intra frames in total, YCoCg, B=8, q=32, sharded frame-parallel (strong scaling)"),
    "c5": (4320, 7680, 16, (32,), "YCrCb", True, 32,
           "configs[4]: 7680x4320 frames, 32 per GPU, YCrCb (float extension) + B=16, q=32, RD statistics all-reduced"),
}
ALG_BYTES_PER_PX = {"encode": 6.0, "decode": 6.0}    # 3 B read + 3 B written each
METRIC = "Mpixel/s encode+decode (color+DCT+deadzone)"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--frames", type=int, default=0, help="frames per GPU per step (0 = workload default)")
    ap.add_argument("--e2e-frames", type=int, default=32, help="4K frames per GPU per end-to-end step")
    ap.add_argument("--decode", default="fp32", choices=["fp32", "fp64"],
                    help="decoder of the headline: fp32 (+-1 LSB, the north star's tolerance) or fp64 "
                         "(the reference's chain, bit-exact); the other one is reported beside it")
    ap.add_argument("--contract", action="store_true", help="encoder: allow fused multiply-adds (VCFB_F_CONTRACT)")
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--only-main", action="store_true", help="skip the side measurements (noise, workloads, e2e extras)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=20.0)
    ap.add_argument("--sustained-seconds", type=float, default=2.0)
    return ap.parse_args()


# ----------------------------------------------------------------------------
# clocks
# ----------------------------------------------------------------------------
class ClockSampler:
    """SM clock and throttle reasons during the timed region, polled through NVML
    every few milliseconds (the timed region is tens of milliseconds long)."""

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.samples = []
        self.stop_flag = threading.Event()
        self.ok = False
        self.mx = None
        try:
            import pynvml
            self.nv = pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[gpu_index]) if vis and vis.split(",")[gpu_index].isdigit() else gpu_index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.mx = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.ok = True
        except Exception:
            self.ok = False

    def _loop(self):
        nv = self.nv
        while not self.stop_flag.is_set():
            try:
                clk = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                rs = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                self.samples.append((time.time(), float(clk), int(rs)))
            except Exception:
                pass
            time.sleep(0.003)

    def start(self):
        if self.ok:
            self.t = threading.Thread(target=self._loop, daemon=True)
            self.t.start()

    def window(self, t0, t1):
        """Summary of the samples taken in [t0, t1] (the sampler keeps running)."""
        if not self.ok:
            return {"sm_mhz": None, "sm_max_mhz": None, "samples": 0, "reasons": ["nvml unavailable"]}
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4)}
        inside = [x for x in list(self.samples) if t0 <= x[0] <= t1]
        sm = sorted(x[1] for x in inside)
        reasons = sorted(n for n, bit in names.items() if any(x[2] & bit for x in inside))
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_min_mhz": sm[0] if sm else None,
                "sm_max_mhz": self.mx, "samples": len(sm), "reasons": reasons}

    def stop(self):
        if self.ok and getattr(self, "t", None) is not None:
            self.stop_flag.set()
            self.t.join(timeout=1)


# ----------------------------------------------------------------------------
# synthetic frames: ONE formula for both arms (torch on the device, numpy on the host)
# ----------------------------------------------------------------------------
def make_frames(torch, n, H, W, device, seed, content="natural"):
    """Synthetic frames generated on the device.  "natural": smooth field + N(0,6) noise
    (SURVEY 8d C2(ii)); "noise": i.i.d. uniform[0,255] (C2(i))."""
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    out = torch.empty((n, H, W, 3), dtype=torch.uint8, device=device)
    chunk = max(1, min(16, (64 << 20) // (H * W)))
    if content == "noise":
        for f0 in range(0, n, chunk):
            m = min(chunk, n - f0)
            out[f0:f0 + m] = torch.randint(0, 256, (m, H, W, 3), generator=g, device=device, dtype=torch.uint8)
        return out
    yy = torch.arange(H, device=device, dtype=torch.float32)[None, :, None]
    xx = torch.arange(W, device=device, dtype=torch.float32)[None, None, :]
    for c in range(3):
        for f0 in range(0, n, chunk):
            m = min(chunk, n - f0)
            ph = torch.rand((3, m, 1, 1), generator=g, device=device) * 6.283
            f = 70 * torch.sin(6.283 * (1 + c) * xx / W + ph[0]) * torch.cos(6.283 * (2 - 0.5 * c) * yy / H + ph[1])
            f += 30 * torch.sin(6.283 * (xx + yy) / 97.0 + ph[2])
            f += torch.randn((m, H, W), generator=g, device=device) * 6 + 128
            out[f0:f0 + m, :, :, c] = f.round_().clamp_(0, 255).to(torch.uint8)
            del f
    return out


def make_frame_np(H, W, seed, content="natural"):
    """The same formula in numpy, for the CPU legs (same distribution; the random streams of
    numpy and torch differ, the statistics do not)."""
    import numpy as np
    rng = np.random.default_rng(seed)
    if content == "noise":
        return rng.integers(0, 256, size=(H, W, 3), dtype=np.uint8)
    yy = np.arange(H, dtype=np.float32)[:, None]
    xx = np.arange(W, dtype=np.float32)[None, :]
    out = np.empty((H, W, 3), dtype=np.uint8)
    for c in range(3):
        ph = rng.random(3, dtype=np.float32) * np.float32(6.283)
        f = 70 * np.sin(np.float32(6.283 * (1 + c)) * xx / W + ph[0]) * np.cos(np.float32(6.283 * (2 - 0.5 * c)) * yy / H + ph[1])
        f = f + 30 * np.sin(np.float32(6.283) * (xx + yy) / np.float32(97.0) + ph[2])
        f = f + rng.standard_normal((H, W), dtype=np.float32) * 6 + 128
        out[:, :, c] = np.clip(np.rint(f), 0, 255).astype(np.uint8)
    return out


# ----------------------------------------------------------------------------
# CPU oracle timing (cpu_baseline leg and --impl reference)
# ----------------------------------------------------------------------------
C2 = WORKLOADS["c2"]


def _cpu_worker(args):
    seed, q, reps, loop, content = args
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    from oracle import vcf_oracle as O
    H, W, B = C2[0], C2[1], C2[2]
    img = make_frame_np(H, W, seed, content)
    t0 = time.perf_counter()
    for _ in range(reps):
        idx = O.encode_array(img, B, q, loop=loop)
        dec = O.decode_array(idx, img.shape, B, q, loop=loop)
    dt = time.perf_counter() - t0
    return dt, int(dec[0, 0, 0])


def cpu_oracle_throughput(budget_s: float, q: int = 32, content: str = "natural"):
    """Encode+decode Mpixel/s of the vectorised oracle with every usable host core
    (one 4K frame per worker process), plus the faithful per-block-loop form on one
    core on a 1/16 frame.  Bounded to about ``budget_s`` seconds."""
    import multiprocessing as mp
    H, W, B = C2[0], C2[1], C2[2]
    try:
        import psutil
        avail = psutil.virtual_memory().available
    except Exception:
        avail = 64 << 30
    try:
        cores_avail = len(os.sched_getaffinity(0))
    except Exception:
        cores_avail = os.cpu_count() or 1
    workers = max(1, min(cores_avail, int(avail // (3 << 30))))
    ctx = mp.get_context("spawn")
    with ctx.Pool(workers) as pool:
        pool.map(_cpu_worker_warm, range(workers))               # import cost outside the timing
        t0 = time.perf_counter()
        res = pool.map(_cpu_worker, [(1000 + i, q, 1, False, content) for i in range(workers)])
        wall = time.perf_counter() - t0
        reps = max(1, min(8, int(budget_s * 0.6 / max(wall, 1e-3))))
        if reps > 1:
            t0 = time.perf_counter()
            res = pool.map(_cpu_worker, [(1000 + i, q, reps, False, content) for i in range(workers)])
            wall = time.perf_counter() - t0
        else:
            reps = 1
    mpx = workers * reps * H * W / 1e6 / wall
    # faithful loop form (what the reference executes), one core, 1/16 frame
    from oracle import vcf_oracle as O
    img = make_frame_np(H // 4, W // 4, 5, content)
    t0 = time.perf_counter()
    idx = O.encode_array(img, B, q, loop=True)
    O.decode_array(idx, img.shape, B, q, loop=True)
    loop_mpx = (H // 4) * (W // 4) / 1e6 / (time.perf_counter() - t0)
    single = H * W / 1e6 / (sum(r[0] for r in res) / len(res) / reps)
    return dict(value=mpx, unit="Mpixel/s", cores=workers, kind="port",
                sample=(f"{workers * reps} synthetic 3840x2160 {content} frames (the GPU arm's generator), encode+decode, "
                        f"q={q}, vectorised numpy/scipy oracle, {workers} processes x {reps} frame(s); wall {wall:.1f}s"),
                per_core_value=single, loop_form_1core_value=loop_mpx,
                loop_form_sample="per-block Python loop (the reference's form), 960x540, 1 core",
                host_cpus=os.cpu_count())


def _cpu_worker_warm(_):
    import numpy  # noqa: F401
    import scipy.fftpack  # noqa: F401
    from oracle import vcf_oracle  # noqa: F401
    return 0


def main_config(n, world):
    H, W = C2[0], C2[1]
    return {"workload": C2[7], "frames_per_gpu_per_step": n, "parallelism": f"frame-parallel x{world}",
            "content": "natural-like synthetic frames (smooth field + N(0,6) noise); `noise` repeats the step on i.i.d. uniform frames",
            "l2": f"inputs exceed L2: {3 * n * H * W * 3 / 1e9:.1f} GB touched per step vs 126 MB"}


def run_reference(a):
    """--impl reference: the reference's CPU implementation of the path.  The four
    arithmetic packages it imports are not installable offline, so
    this times the oracle port with every usable host core; rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    H, W, QS = C2[0], C2[1], C2[3]
    per_step = []
    base = None
    for s in range(a.warmup + a.steps):
        base = cpu_oracle_throughput(a.cpu_seconds / max(1, a.steps), q=QS[s % len(QS)])
        if s >= a.warmup:
            per_step.append(base["value"])
    v = sum(per_step) / len(per_step)
    base["value"] = v
    px = base["cores"] * H * W
    cfg = main_config(a.frames, max(1, a.gpus))      # identical to the GPU arm's: same workload, same content generator
    line = dict(impl="reference", metric=METRIC, value=v, unit="Mpixel/s",
                sample_note="each step of this arm is a bounded sample of the workload: one frame per host core",
                n_gpus=a.gpus, steps=a.steps, warmup=a.warmup, ms_per_step=px / 1e6 / v * 1e3,
                higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32 encode / f64 decode",
                data="synthetic", config=cfg, cpu_baseline=base,
                e2e={"value": v, "unit": "Mpixel/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0})
    print(json.dumps(line))


# ----------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------
class Ctx:
    """Per-process state of the GPU arm."""

    def __init__(self):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device; there is no CPU fallback")
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            self.peak = float(json.load(open(peaks_path))["hbm_gbs"])
            self.peak_src = "MEASURED_PEAKS.json hbm_gbs (of measured)"
        else:
            self.peak, self.peak_src = 6650.0, "B200_PROFILING.md fallback (of fallback)"

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, v: float) -> float:
        if self.world == 1:
            return float(v)
        t = self.torch.tensor([v], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(self, v: float) -> float:
        if self.world == 1:
            return float(v)
        t = self.torch.tensor([v], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return float(t.item())


def timed_steps(cx: Ctx, step, steps, warmup, sampler=None, finalize=None):
    """W untimed + K timed steps bracketed by barrier + synchronize; CUDA events on the launching
    stream; returns (total ms = max over ranks, mean encode ms, mean decode ms, launches, wall window)."""
    from vcf_b200 import _lib
    torch = cx.torch
    for s in range(warmup):
        step(s)
    if finalize:
        finalize()
    cx.barrier()
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(steps)]
    cx.barrier()
    launches0 = _lib.launch_count()
    t0w = time.time()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in range(steps):
        step(s, evs[s])
    if finalize:
        finalize()             # e.g. the last batch's all-reduce: inside the timed region
    e1.record()
    cx.barrier()
    t1w = time.time()
    launches = _lib.launch_count() - launches0
    ms = cx.max_over_ranks(e0.elapsed_time(e1))
    enc_ms = sum(e[0].elapsed_time(e[1]) for e in evs) / steps
    dec_ms = sum(e[1].elapsed_time(e[2]) for e in evs) / steps
    return ms, enc_ms, dec_ms, launches, (t0w, t1w)


def measure_rd_sweep(cx: Ctx, a, frames_per_gpu: int = 1):
    """BASELINE configs[2]: the rate/distortion sweep B in {4,8,16,32} x 8 steps over a 4K frame, through the
    fused kernel (one pass per block size, vcfb_rd_sweep_dev) -- every rank sweeps its own frame(s) and the
    4 x 8 statistics vectors are summed over the ranks by one NCCL all-reduce per sweep (asynchronous, waited
    for inside the timed region).  `value` counts pixels x (B, q) points per second, whole job."""
    from vcf_b200 import _lib
    from vcf_b200.rd import rd_stats_fused
    torch = cx.torch
    H, W = 2160, 3840
    BS, QS = (4, 8, 16, 32), (4, 8, 12, 16, 24, 32, 48, 64)
    x = make_frames(torch, frames_per_gpu, H, W, cx.dev, 4321 + cx.rank, "natural")
    pend = {}

    def step(s, ev=None):
        if ev:
            ev[0].record()
        tab = torch.stack([rd_stats_fused(x, B, QS) for B in BS])       # (4, 8, 776) int64
        if ev:
            ev[1].record()
        if pend.get("work") is not None:
            pend["work"].wait()
        pend["v"] = tab
        pend["work"] = cx.dist.all_reduce(tab, async_op=True) if cx.world > 1 else None
        if ev:
            ev[2].record()

    def finalize():
        if pend.get("work") is not None:
            pend["work"].wait()
            pend["work"] = None

    ms, sweep_ms, _, launches, _ = timed_steps(cx, step, a.steps, a.warmup, finalize=finalize)
    npts = len(BS) * len(QS)
    px = frames_per_gpu * H * W
    # one point per block size through the separate kernels, for the ratio (encode + float64 decode + statistics, q = 32)
    from vcf_b200.rd import rd_point
    for B in BS:                     # first use loads the kernels
        rd_point(x, B, 32)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for B in BS:
        rd_point(x, B, 32)
    torch.cuda.synchronize()
    per_point_ms = (time.perf_counter() - t0) * 1e3 / len(BS)
    st = pend["v"][1, 5].cpu().numpy()                                   # B = 8, q = 32
    return {"value": cx.world * px * npts * a.steps / 1e6 / (ms / 1e3), "unit": "Mpixel/s x (B, q) points",
            "ms_per_step": ms / a.steps, "ms_per_sweep_kernels": sweep_ms, "points": npts,
            "frames_per_gpu_per_step": frames_per_gpu, "kernel": _lib.last_kernel(), "gpu_launches": launches,
            "alg_bytes_per_step": 3.0 * px * len(BS), "per_point_path_ms_per_point": per_point_ms,
            "speedup_over_per_point_path": per_point_ms * npts / (ms / a.steps),
            "check_B8_q32": {"sse": int(st[0:3].sum()), "nonzero": int(st[4])},
            "bound": "issue slots / FP64 pipe (float64 inverse chain per step), not HBM: profiles/r2_rd_sweep_ncu_summary.json"}


def fracs(cx, px_step, enc_ms, dec_ms, dec_extra_b=0.0):
    p = cx.peak
    return {"encode_frac": 6.0 * px_step / (enc_ms / 1e3) / 1e9 / p,
            "decode_frac": (6.0 + dec_extra_b) * px_step / (dec_ms / 1e3) / 1e9 / p,
            "roundtrip_frac_of_12B_per_px": (12.0 + dec_extra_b) * px_step / ((enc_ms + dec_ms) / 1e3) / 1e9 / p}


def measure_transform(cx: Ctx, a, wl: str, n: int, content: str, fp64_dec: bool, frames=None, bufs=None, fast_enc=None):
    """One workload at this N: returns a dict with value (whole job), ms_per_step, per-kernel times
    and roofline fractions.  ``rde`` workloads run the statistics and the NCCL all-reduce of the
    int64[776] vector inside every timed step."""
    from vcf_b200 import Codec, _lib
    from vcf_b200.frames import allreduce_stats
    torch = cx.torch
    H, W, B, QS, color, rde, _, desc = WORKLOADS[wl]
    x = frames if frames is not None else make_frames(torch, n, H, W, cx.dev, 1234 + cx.rank, content)
    Hp, Wp = (H + B - 1) // B * B, (W + B - 1) // B * B
    if bufs is not None:
        idx, y = bufs
    else:
        idx = torch.empty((n, Hp, Wp, 3), dtype=torch.uint8, device=cx.dev)
        y = torch.empty((n, H, W, 3), dtype=torch.uint8, device=cx.dev)
    # fast mode (the north star's): tensor-core encoder + float32 decoder; exact mode: the bit-exact pair
    if fast_enc is None:
        fast_enc = not fp64_dec and not rde
    enc = {q: Codec(block_size=B, q=q, contract=a.contract, color=color, hist=False, fast=fast_enc) for q in QS}
    dec = {q: Codec(block_size=B, q=q, fp64=fp64_dec, color=color) for q in QS}
    NQ = len(QS)
    seen = {}
    last_stats = {}

    def step(s, ev=None):
        q = QS[s % NQ]
        if ev:
            ev[0].record()
        if rde:
            _, st_e = enc[q].encode(x, out=idx, stats=True)
        else:
            enc[q].encode(x, out=idx)
        seen["encode"] = _lib.last_kernel()
        if ev:
            ev[1].record()
        if rde:
            r = dec[q].decode(idx, (H, W), out=y, original=x, stats=True)
            # one NCCL all-reduce of int64[776] per batch, issued asynchronously: batch s+1 is transformed while
            # the statistics of batch s are being summed; every reduction is waited for inside the timed region
            if last_stats.get("work") is not None:
                last_stats["work"].wait()
            last_stats["v"], last_stats["work"] = allreduce_stats(r[-1] + st_e, async_op=True)
        else:
            dec[q].decode(idx, (H, W), out=y)
        seen["decode"] = _lib.last_kernel()
        if ev:
            ev[2].record()

    def finalize():
        if last_stats.get("work") is not None:
            last_stats["work"].wait()
            last_stats["work"] = None

    ms, enc_ms, dec_ms, launches, win = timed_steps(cx, step, a.steps, a.warmup, finalize=finalize)
    px_step = n * H * W
    res = {"value": cx.world * px_step * a.steps / 1e6 / (ms / 1e3), "unit": "Mpixel/s",
           "ms_per_step": ms / a.steps, "frames_per_gpu_per_step": n,
           "encode_ms_per_launch": enc_ms, "decode_ms_per_launch": dec_ms,
           "kernels": dict(seen), "gpu_launches": launches, "content": content,
           "encode": ("f32 fast mode (tensor cores): < 1e-6 of the indices differ from the reference's float32 path"
                      if fast_enc else "f32, bit-exact with the reference's float32 path"),
           "decode": "f64, the reference's chain, bit-exact" if fp64_dec else "f32, pixels within +-1 LSB / PSNR within 0.01 dB",
           "_window": win, "_px_step": px_step}
    res.update(fracs(cx, px_step, enc_ms, dec_ms, 3.0 if rde else 0.0))
    if rde:
        v = last_stats["v"]
        res["collective"] = "NCCL all-reduce(sum) of int64[776] per step (async, overlapped with the next batch), all inside the timed region" if cx.world > 1 \
            else "single rank: the all-reduce is the identity"
        res["allreduced_nsamples"] = int(v[3].item())
        res["expected_nsamples"] = cx.world * px_step * 3
    return res, (x, idx, y), (enc, dec)


def strip(d):
    return {k: v for k, v in d.items() if not k.startswith("_")}


def run_ours(a):
    import numpy as np
    from vcf_b200 import Codec, _lib
    cx = Ctx()
    torch = cx.torch
    world, rank, dev = cx.world, cx.rank, cx.dev
    assert _lib.lib().vcfb_device_count() > 0
    H, W, B, QS, COLOR, rde_main, _, desc = WORKLOADS[a.workload]
    n = a.frames
    NQ = len(QS)
    fp64_head = a.decode == "fp64"

    sampler = ClockSampler(cx.local)
    if rank == 0:
        sampler.start()
        time.sleep(0.05)

    # ---- headline --------------------------------------------------------------------------
    head, (x, idx, y), (enc, dec) = measure_transform(cx, a, a.workload, n, "natural", fp64_head)
    clocks = sampler.window(*head["_window"]) if rank == 0 else None
    px_step = head["_px_step"]

    # ---- the other decoder on the same frames -----------------------------------------------
    other = None
    if not rde_main and B == 8:
        other, _, (_, dec_o) = measure_transform(cx, a, a.workload, n, "natural", not fp64_head, frames=x, bufs=(idx, y))
    exact = head if fp64_head else other
    fast = other if fp64_head else head
    # what the fast mode costs in fidelity, measured on this very batch (q = 8, the finest step of the cycle)
    fidelity = None
    if other is not None:
        qf = QS[0]
        k_fast = Codec(block_size=B, q=qf, color=COLOR, fast=True).encode(x)
        k_exact = Codec(block_size=B, q=qf, color=COLOR).encode(x)
        nd = int((k_fast != k_exact).sum().item())
        y_fast = Codec(block_size=B, q=qf, color=COLOR).decode(k_exact, (H, W))
        y_exact = Codec(block_size=B, q=qf, color=COLOR, fp64=True).decode(k_exact, (H, W))
        dmax = int((y_fast.to(torch.int16) - y_exact.to(torch.int16)).abs().max().item())
        npx = int((y_fast != y_exact).sum().item())
        fidelity = {"q": qf, "indices_differing": nd, "indices": k_exact.numel(), "index_mismatch_rate": nd / k_exact.numel(),
                    "decoded_samples_differing": npx, "decoded_samples": y_exact.numel(), "max_abs_pixel_difference": dmax,
                    "what": "fast-mode encoder vs bit-exact encoder on the bench batch; fast-mode decoder vs bit-exact decoder on the exact indices"}
        del k_fast, k_exact, y_fast, y_exact

    # ---- i.i.d. uniform frames: dense indices at every q ------------------------
    noise = None
    if not a.only_main and a.workload == "c2":
        xn = make_frames(torch, n, H, W, dev, 4321 + rank, "noise")
        nf, _, _ = measure_transform(cx, a, "c2", n, "noise", False, frames=xn, bufs=(idx, y))
        ne, _, _ = measure_transform(cx, a, "c2", n, "noise", True, frames=xn, bufs=(idx, y))
        noise = {"value": nf["value"] if not fp64_head else ne["value"], "unit": "Mpixel/s",
                 "fast_mode": strip(nf), "exact_mode": strip(ne),
                 "note": "i.i.d. uniform[0,255] frames; same step, same q cycle; `value` is the headline's mode"}
        del xn

    # ---- sustained: the headline step back to back for a couple of seconds ------------------
    sustained = None
    if not a.only_main and a.sustained_seconds > 0:
        reps = max(a.steps, int(a.sustained_seconds * 1e3 / max(head["ms_per_step"], 1e-3)))
        dd = dec
        cx.barrier()
        t0w = time.time()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for s in range(reps):
            q = QS[s % NQ]
            enc[q].encode(x, out=idx)
            dd[q].decode(idx, (H, W), out=y)
        e1.record()
        cx.barrier()
        t1w = time.time()
        sms = cx.max_over_ranks(e0.elapsed_time(e1))
        sustained = {"value": world * px_step * reps / 1e6 / (sms / 1e3), "unit": "Mpixel/s", "steps": reps,
                     "seconds": sms / 1e3, "ms_per_step": sms / reps,
                     "clocks": sampler.window(t0w, t1w) if rank == 0 else None}

    # ---- single-frame latency, frame resident in HBM (SURVEY.md 8d "Method") ----------------
    latency = None
    if not rde_main:
        x1, i1_, y1 = x[:1], idx[:1], y[:1]
        reps = 20
        for q in QS:
            enc[q].encode(x1, out=i1_)
            dec[q].decode(i1_, (H, W), out=y1)
        torch.cuda.synchronize()
        le = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        le[0].record()
        for r in range(reps):
            enc[QS[r % NQ]].encode(x1, out=i1_)
        le[1].record()
        for r in range(reps):
            dec[QS[r % NQ]].decode(i1_, (H, W), out=y1)
        le[2].record()
        torch.cuda.synchronize()
        latency = {"encode_ms": le[0].elapsed_time(le[1]) / reps, "decode_ms": le[1].elapsed_time(le[2]) / reps,
                   "note": "one frame per call, back-to-back launches on one stream, q cycled"}

    # ---- end to end through the host API: pinned numpy in, pinned numpy out ----------
    ne_ = max(1, min(a.e2e_frames, n, (832 << 20) // (H * W * 3)))
    Hp, Wp = (H + B - 1) // B * B, (W + B - 1) // B * B
    hx = torch.empty((ne_, H, W, 3), dtype=torch.uint8, pin_memory=True)
    hx.copy_(x[:ne_])
    hidx = torch.empty((ne_, Hp, Wp, 3), dtype=torch.uint8, pin_memory=True)
    hy = torch.empty((ne_, H, W, 3), dtype=torch.uint8, pin_memory=True)
    hxn, hidxn, hyn = hx.numpy(), hidx.numpy(), hy.numpy()
    enc_h = {q: Codec(block_size=B, q=q, contract=a.contract, device=cx.local, color=COLOR, fast=not fp64_head) for q in QS}
    dec_h = {q: Codec(block_size=B, q=q, fp64=fp64_head, device=cx.local, color=COLOR) for q in QS}

    def e2e_step(s):
        q = QS[s % NQ]
        enc_h[q].encode(hxn, out=hidxn)
        dec_h[q].decode(hidxn, (H, W), out=hyn)

    for s in range(max(a.warmup, 4)):
        e2e_step(s)
    cx.barrier()
    t0 = time.perf_counter()
    for s in range(a.steps):
        e2e_step(s)
    torch.cuda.synchronize()
    e2e_s = cx.max_over_ranks(time.perf_counter() - t0)
    e2e_val = world * ne_ * H * W * a.steps / 1e6 / e2e_s
    # the last device-timed step and the last e2e step used the same q when steps % NQ == 0
    dec[QS[(a.steps - 1) % NQ]].decode(enc[QS[(a.steps - 1) % NQ]].encode(x[:1]), (H, W), out=y[:1])
    same = bool(torch.equal(hy[:1].to(dev), y[:1]))
    e2e = {"value": e2e_val, "unit": "Mpixel/s", "h2d_bytes_per_step": 2 * ne_ * H * W * 3,
           "d2h_bytes_per_step": 2 * ne_ * H * W * 3, "frames_per_step": ne_,
           "api": "vcf_b200.Codec.encode/decode on pinned numpy arrays (vcfb_encode_host/vcfb_decode_host)",
           "matches_device_path": same,
           "achieved_GB_s_each_way_per_gpu": 2 * ne_ * H * W * 3 * a.steps / 1e9 / e2e_s}

    # ---- the copy loop alone over the same pinned buffers (what PCIe / host memory allow at this N) ----
    copy_ceiling = None
    e2e_cs = None
    if not a.only_main:
        s_up, s_dn = torch.cuda.Stream(), torch.cuda.Stream()
        d_a = torch.empty_like(x[:ne_])
        d_b = torch.empty_like(x[:ne_])
        d_c = torch.empty_like(x[:ne_])

        def copy_step():
            # one e2e step moves two arrays up and two down; both directions at once
            with torch.cuda.stream(s_up):
                d_a.copy_(hx, non_blocking=True)
                d_b.copy_(hx, non_blocking=True)
            with torch.cuda.stream(s_dn):
                hy.copy_(d_c, non_blocking=True)
                hy.copy_(d_c, non_blocking=True)
        for _ in range(2):
            copy_step()
        cx.barrier()
        t0 = time.perf_counter()
        for _ in range(a.steps):
            copy_step()
        torch.cuda.synchronize()
        cs = cx.max_over_ranks(time.perf_counter() - t0)
        gbs = 2 * ne_ * H * W * 3 * a.steps / 1e9 / cs
        copy_ceiling = {"GB_s_each_way_per_gpu": gbs, "ranks_copying_at_once": world,
                        "implied_e2e_ceiling_mpixel_s": world * gbs * 1e9 / 6.0 / 1e6,
                        "e2e_fraction_of_ceiling": e2e_val / (world * gbs * 1e9 / 6.0 / 1e6),
                        "what": "cudaMemcpyAsync H2D and D2H at once on two streams, the e2e step's buffers and byte counts, no kernels"}
        torch.cuda.synchronize()
        del d_a, d_b, d_c

        # ---- encode leg with the entropy front-end on the GPU: only the code-stream comes back ----
        try:
            from vcf_b200.entropy import encode_to_codestream
            q_cs = 32 if 32 in QS else QS[0]
            enc_cs = Codec(block_size=B, q=q_cs, color=COLOR, device=cx.local)
            for _ in range(2):
                cs_list = encode_to_codestream(enc_cs, hxn)
            cx.barrier()
            t0 = time.perf_counter()
            for _ in range(a.steps):
                cs_list = encode_to_codestream(enc_cs, hxn)
            torch.cuda.synchronize()
            stream_bytes = b"".join(cs_list)
            cs_s = cx.max_over_ranks(time.perf_counter() - t0)
            # encode leg of the plain host API for comparison (RGB up, indices down)
            cx.barrier()
            t0 = time.perf_counter()
            for _ in range(a.steps):
                enc_h[q_cs].encode(hxn, out=hidxn)
            torch.cuda.synchronize()
            en_s = cx.max_over_ranks(time.perf_counter() - t0)
            e2e_cs = {"value": world * ne_ * H * W * a.steps / 1e6 / cs_s, "unit": "Mpixel/s (encode leg only)",
                      "q": q_cs, "h2d_bytes_per_step": ne_ * H * W * 3, "d2h_bytes_per_step": len(stream_bytes),
                      "bits_per_pixel": 8.0 * len(stream_bytes) / (ne_ * H * W),
                      "encode_leg_indices_over_pcie_mpixel_s": world * ne_ * H * W * a.steps / 1e6 / en_s,
                      "api": "vcf_b200.entropy.encode_to_codestream: pinned RGB -> encode -> vcfb_deflate_dev -> raw deflate stream on the host"}
        except Exception as exc:
            e2e_cs = {"error": repr(exc)}
    del hx, hidx, hy

    # ---- other BASELINE configs at this N, each with its own timed region ---------------------
    workloads = {}
    if not a.only_main and a.workload == "c2":
        rde_res, _, _ = measure_transform(cx, a, "rde", n, "natural", True, frames=x, bufs=(idx, y))
        rde_res["config"] = WORKLOADS["rde"][7]
        rde_res["scaling"] = "weak"
        workloads["rde"] = strip(rde_res)
        try:     # the same step in the fast mode: tensor-core kernels + the streaming statistics passes over their outputs
            r2, _, _ = measure_transform(cx, a, "rde", n, "natural", False, frames=x, bufs=(idx, y), fast_enc=True)
            workloads["rde"]["fast_mode"] = {k: r2[k] for k in ("value", "ms_per_step", "encode_ms_per_launch", "decode_ms_per_launch",
                                                                 "kernels", "gpu_launches")}
        except Exception as exc:
            workloads["rde"]["fast_mode"] = {"error": repr(exc)}
        x1 = i1_ = y1 = None
        del x, idx, y
        torch.cuda.empty_cache()
        for wl in ("c4", "c5"):
            try:
                total = WORKLOADS[wl][6]
                n_wl = total // world if wl == "c4" else total
                # fast mode first (c4: tensor cores; c5: bit-exact float32 encoder + float32 decoder), the bit-exact pair beside it
                r, bufs_, _ = measure_transform(cx, a, wl, n_wl, "natural", False)
                r2, _, _ = measure_transform(cx, a, wl, n_wl, "natural", True, frames=bufs_[0], bufs=bufs_[1:])
                r["exact_mode"] = {k: r2[k] for k in ("value", "ms_per_step", "encode_ms_per_launch", "decode_ms_per_launch",
                                                      "encode_frac", "decode_frac", "kernels")}
                r["config"] = WORKLOADS[wl][7]
                r["scaling"] = "strong" if wl == "c4" else "weak"
                workloads[wl] = strip(r)
                del bufs_, r
            except Exception as exc:
                workloads[wl] = {"error": repr(exc)}
            torch.cuda.empty_cache()
        try:
            r = measure_rd_sweep(cx, a)
            r["config"] = ("configs[2]: RD sweep on a 3840x2160 frame per GPU, B in {4,8,16,32} x q in {4,8,12,16,24,32,48,64}, "
                           "fused (forward once per B, every step on chip), statistics all-reduced")
            r["scaling"] = "weak"
            workloads["c3"] = r
        except Exception as exc:
            workloads["c3"] = {"error": repr(exc)}
    sampler.stop()

    if rank != 0:
        if world > 1:
            cx.dist.destroy_process_group()
        return

    enc_ms, dec_ms = head["encode_ms_per_launch"], head["decode_ms_per_launch"]
    dom = "encode" if enc_ms >= dec_ms else "decode"
    dom_ms = max(enc_ms, dec_ms)
    extra_b = 3.0 if (rde_main and dom == "decode") else 0.0
    achieved = (ALG_BYTES_PER_PX[dom] + extra_b) * px_step / (dom_ms / 1e3) / 1e9
    # DRAM traffic of the dominant kernel: from the committed `ncu --set full` capture of this command
    # (profiles/ncu_traffic.json names the capture); null when the capture is of another kernel
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tpath):
        try:
            tj = json.load(open(tpath))
            ent = tj.get("kernels", {}).get(head["kernels"].get(dom))
            if ent and ent.get("frames") == n:
                traffic, traffic_src = ent["bytes"], f"profiles/ncu_traffic.json <- {tj.get('source')} (ncu capture, not measured in this run)"
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "kernel": f"{head['kernels'].get(dom)} ({dom})", "kernels": head["kernels"],
                "achieved": achieved, "peak": cx.peak, "unit": "GB/s", "frac": achieved / cx.peak,
                "traffic": traffic, "traffic_source": traffic_src, "peak_source": cx.peak_src,
                "alg_bytes_per_launch": (ALG_BYTES_PER_PX[dom] + extra_b) * px_step,
                "encode_ms_per_launch": enc_ms, "decode_ms_per_launch": dec_ms}
    roofline.update(fracs(cx, px_step, enc_ms, dec_ms, 3.0 if rde_main else 0.0))

    cfg = main_config(n, world) if a.workload == "c2" else {"workload": desc, "frames_per_gpu_per_step": n,
                                                             "parallelism": f"frame-parallel x{world}"}
    line = {"metric": METRIC, "value": head["value"], "unit": "Mpixel/s",
            "n_gpus": world, "steps": a.steps, "warmup": a.warmup, "ms_per_step": head["ms_per_step"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": ("f32 encode (bit-exact with the reference float32 path) / f64 decode (reference chain, bit-exact)" if fp64_head else
                      "f32 fast mode of the north star: encode on tensor cores (fp16 x 2-limb operands, fp32 accumulation; < 1e-6 of "
                      "the indices differ from the reference's float32 path) / f32 decode on tensor cores (pixels within +-1 LSB, "
                      "PSNR within 0.01 dB); `exact_mode` = the bit-exact pair"),
            "data": "synthetic", "config": cfg, "clocks": clocks, "e2e": e2e,
            "gpu_launches": head["gpu_launches"], "roofline": roofline}
    if fidelity:
        line["fast_mode_fidelity"] = fidelity
    if exact is not None and exact is not head:
        line["exact_mode"] = strip(exact)
    if fast is not None and fast is not head:
        line["fast_mode"] = strip(fast)
    if noise:
        line["value_noise"] = noise["value"]
        line["noise"] = noise
    if sustained:
        line["sustained"] = sustained
    if latency:
        line["single_frame_latency"] = latency
    if copy_ceiling:
        line["host_copy_ceiling"] = copy_ceiling
    if e2e_cs:
        line["e2e_codestream"] = e2e_cs
    if workloads:
        line["workloads"] = workloads
    if world == 1 and not a.no_cpu_baseline:
        line["cpu_baseline"] = cpu_oracle_throughput(a.cpu_seconds)
        line["entropy_stage"] = entropy_stage(cx, a, WORKLOADS[a.workload])
    print(json.dumps(line))
    if world > 1:
        cx.dist.destroy_process_group()


def entropy_stage(cx, a, wl):
    """The stage on the other side of the path, outside every timed region above (BASELINE:
    "reported separately"): the host entropy coder on one frame's index planes, and the same
    stage on the GPU (row F4)."""
    torch = cx.torch
    H, W, B, QS, COLOR = wl[0], wl[1], wl[2], wl[3], wl[4]
    try:
        import io
        import zlib
        import numpy as np
        from vcf_b200 import Codec, _lib as _L
        from vcf_b200.entropy import deflate_raw_dev, savez_compressed
        q_ent = 32 if 32 in QS else QS[0]
        codec = Codec(block_size=B, q=q_ent, color=COLOR)
        nf = 16 if H * W <= 3840 * 2160 else 2
        x = make_frames(torch, nf, H, W, cx.dev, 99, "natural")
        k = codec.encode(x[:1])[0].cpu().numpy()
        t0 = time.perf_counter()
        comp = zlib.compress(k.tobytes(), 6)
        dt = time.perf_counter() - t0
        out = {"codec": "zlib level 6 on one frame of indices (q=%d), 1 host core" % q_ent,
               "mpixel_s": H * W / 1e6 / dt, "bits_per_pixel": 8.0 * len(comp) / (H * W),
               "note": "not part of `value` or `e2e`; the reference's containers (TIFF/PNG/npz) wrap the same deflate"}
        kx = codec.encode(x)
        geom = (kx.shape[2] * kx.shape[3], kx.shape[3])      # bytes per row and per pixel of the H x W x 3 index image
        kb = kx.reshape(-1)
        dst, nb = deflate_raw_dev(kb, geom)           # warm-up, and the stream that is checked
        one, nb1 = deflate_raw_dev(kb[: k.size], geom)
        torch.cuda.synchronize()
        ok = zlib.decompress(one[: int(nb1.item())].cpu().numpy().tobytes(), -15) == k.tobytes()
        L_ = _L.lib()
        ws = torch.empty(L_.vcfb_deflate_workspace(kb.numel()), dtype=torch.uint8, device=kb.device)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 10
        e0.record()
        for _ in range(reps):
            _L.check(L_.vcfb_deflate_rows_dev(kb.data_ptr(), kb.numel(), geom[0], geom[1], dst.data_ptr(), dst.numel(),
                                              nb.data_ptr(), ws.data_ptr(), ws.numel(), torch.cuda.current_stream().cuda_stream))
        e1.record()
        torch.cuda.synchronize()
        gms = e0.elapsed_time(e1) / reps
        out["gpu_deflate"] = {
            "api": "vcfb_deflate_rows_dev (runs + previous sample + samples above as match candidates), %d frames of indices per call, in HBM" % nf,
            "size_vs_zlib6_one_frame": int(nb1.item()) / len(comp),
            "ms_per_call": gms, "mpixel_s": nf * H * W / 1e6 / (gms / 1e3), "input_GB_s": kb.numel() / 1e9 / (gms / 1e3),
            "bits_per_pixel": 8.0 * int(nb.item()) / (nf * H * W),
            "bits_per_pixel_one_frame": 8.0 * int(nb1.item()) / (H * W),
            "zlib_reads_it_back": bool(ok)}
        # the stage as the chain calls it (src/z_lib.py:19-23): one frame of indices (host array) -> .npz bytes
        t0 = time.perf_counter()
        ref_buf = io.BytesIO()
        np.savez_compressed(ref_buf, a=k)
        t_ref = time.perf_counter() - t0
        savez_compressed(io.BytesIO(), a=k)        # warm-up
        t0 = time.perf_counter()
        reps = 5
        for _ in range(reps):
            our_buf = io.BytesIO()
            savez_compressed(our_buf, a=k)
        t_our = (time.perf_counter() - t0) / reps
        our_buf.seek(0)
        out["npz_end_to_end"] = {
            "what": "one frame of indices, numpy array in -> .npz bytes out (H2D of the array, GPU deflate + GPU CRC-32, D2H of the stream, zip layout on the host)",
            "ms_np_savez_compressed": 1e3 * t_ref, "ms_vcf_b200_savez_compressed": 1e3 * t_our,
            "bytes_np": ref_buf.getbuffer().nbytes, "bytes_vcf_b200": our_buf.getbuffer().nbytes,
            "np_load_reads_it_back": bool(np.array_equal(np.load(our_buf)["a"], k))}
        return out
    except Exception as exc:      # never let the side measurement break the bench line
        return {"error": repr(exc)}


def main():
    a = parse()
    if a.frames <= 0:
        a.frames = WORKLOADS[a.workload][6]
    a.e2e_frames = max(1, min(a.e2e_frames, a.frames))
    if a.gpus > 1 and "WORLD_SIZE" not in os.environ:
        import socket
        s = socket.socket()
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
        s.close()
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={a.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.abspath(__file__)] + sys.argv[1:]
        raise SystemExit(subprocess.call(cmd))
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)


if __name__ == "__main__":
    main()
