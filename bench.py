#!/usr/bin/env python
"""bench.py -- audio-hours/sec of the SED hot path (fused log-mel + Cnn_AvgPooling frame probabilities) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--clips C] [--impl ours|reference]

One "step" = one pass of the hot path over one batch of synthetic TAU-SED-2019-shaped clips (60 s, 48 kHz, mono):
waveform [C, 2 880 000] f32 -> fused framing+Hann+DFT+mel+dB (tcgen05) -> per-mel-bin normalisation -> Cnn_AvgPooling
(main.py:35 config 32-64-128-128) -> frame probabilities [C, 176, 1].  Clips are independent units: with N GPUs every
rank processes its own C clips (weak scaling, no data-path collective); `value` is the whole-job aggregate.

Prints ONE JSON line (rank 0).  Keys beyond the base contract:
  roofline      dominant kernel (logmel_fused_kernel): algorithmic HBM bytes / CUDA-event time vs MEASURED_PEAKS.json
  tensor        executed tensor-core FLOPs of the same kernel vs the measured bf16 peak (the MMAs run at their floor; the rest is operand production)
  cpu_baseline  the oracle port of the reference CPU path timed on this box's host cores on a bounded sample
  e2e           same metric through the C-ABI host-buffer entry point (pinned host memory, H2D + D2H inside the timing)
  config3       BASELINE config 3: Cnn_AvgPooling frame inference, 128 clips IN TOTAL split 128/N over the ranks (strong scaling)
  config4       BASELINE config 4: one training step per GPU on 64 ten-second crops (fused log-mel + native train-mode
                forward / WeightedBCE / backward + ONE NCCL all-reduce + fused Adam-amsgrad), whole step in a CUDA graph
  config5       BASELINE config 5: M5 on raw waveform frames, 128 frames in total split over the ranks, plus a 1024-frame point
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

CLIP_SAMPLES = 2_880_000          # 60 s x 48 kHz
CLIP_SECONDS = 60.0
FRAMES = 182                      # 1 + 2 880 000 // 15 840
ALGO_BYTES_PER_CLIP = 11_566_592  # 11 520 000 read + 46 592 written
CNN_FLOP_PER_CLIP = 965_768_704
# executed tensor FLOPs of the fused log-mel kernel per frame: 96 tcgen05.mma of 128x128x16 (stage 1: 8 K-chunks x
# 6, stage 2: 4 K-chunks x 12; the x3 split products are included), 2 FLOP per MAC
LOGMEL_MMA_FLOP_PER_FRAME = 96 * 128 * 128 * 16 * 2
METRIC = "audio-hours/sec (log-mel + CNN frame SED)"


LOGMEL_SOURCES = ("logmel.cuh", "umma.cuh", "host_tables.h")


def kernel_source_sha(names=LOGMEL_SOURCES):
    """sha256 over the sources of a kernel: ties a committed ncu capture to the code it was taken from."""
    import hashlib
    h = hashlib.sha256()
    for n in names:
        with open(os.path.join(ROOT, "soundeventdetection-pytorch_b200", "csrc", n), "rb") as f:
            h.update(f.read())
    return h.hexdigest()[:16]


def profiled_logmel(clips):
    """DRAM bytes per launch and tensor-pipe activity of the dominant kernel from the newest committed `ncu --set full`
    capture (profiles/r*_logmel_full.json, same clip count), and whether the kernel sources changed since it was taken."""
    import glob
    best = None
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_logmel_full.json"))):
        try:
            d = json.load(open(path))
            if int(d.get("clips", -1)) == int(clips):
                best = (path, d)
        except Exception:
            pass
    if best is None:
        return None
    path, d = best
    stale = d.get("source_sha") != kernel_source_sha()
    return {"traffic": float(d["traffic_bytes_per_launch"]), "pipe_active_pct": float(d["tensor_pipe_active_pct"]),
            "stale": bool(stale), "capture": os.path.relpath(path, ROOT)}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            d = json.load(open(path))
            return float(d["hbm_gbs"]), float(d.get("bf16_tflops_sustained", d["bf16_tflops"])), "measured"
        except Exception:
            pass
    return 6650.0, 1400.0, "fallback"


def burst_tflops():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops"])
    except Exception:
        return 1600.0


class ClockSampler:
    """Streams nvidia-smi clocks / throttle reasons for one GPU (one background process, 50 ms period) and keeps the
    samples taken between mark_start() and stop(): the timed region."""

    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.t0, self.proc = [], None, None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                          "-i", str(index), "-lms", "50"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.strip().split(",")]))

    def mark_start(self):
        self.t0 = time.perf_counter()

    def stop(self):
        t1 = time.perf_counter()
        time.sleep(0.12)
        if self.proc is not None:
            self.proc.terminate()
        rows = [r for (ts, r) in self.rows if self.t0 is not None and self.t0 - 0.06 <= ts <= t1 + 0.12]
        if not rows:
            rows = [r for (_, r) in self.rows[-3:]]

        def num(x):
            try:
                return float(x)
            except Exception:
                return None
        sm = sorted(v for v in (num(r[0]) for r in rows if r) if v is not None)
        mx = [v for v in (num(r[1]) for r in rows if len(r) > 1) if v is not None]
        pw = [v for v in (num(r[2]) for r in rows if len(r) > 2) if v is not None]
        reasons = set()
        for r in rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v == "Active":
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx[0] if mx else None,
                "power_w_max": max(pw) if pw else None, "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------- CPU baseline
def _cpu_logmel_clip(seed):
    """Reference formulation of dataset/spectogram/preprocess.py:21-45 on one synthetic clip (oracle port)."""
    import numpy as np
    from oracle import logmel_ref
    import signals
    y = signals.white(CLIP_SAMPLES, seed)
    spec = logmel_ref.multichannel_stft(y[:, None])                    # float64 rFFT per frame -> complex64
    power = np.abs(spec) ** 2                                           # float32
    mel = np.dot(power, logmel_ref.mel_filter_bank_matrix())            # the reference's 3-D x 2-D np.dot
    return logmel_ref.power_to_db(mel).astype(np.float32)[0]


def cpu_reference_pass(n_clips, workers):
    """Oracle port of the whole path on the host: log-mel per clip (process pool) + Cnn_AvgPooling on torch CPU."""
    import multiprocessing as mp
    import numpy as np
    import torch
    from oracle import cnn_ref
    import refmodels
    t0 = time.perf_counter()
    if workers > 1:
        with mp.get_context("fork").Pool(workers) as pool:
            lms = pool.map(_cpu_logmel_clip, range(n_clips))
    else:
        lms = [_cpu_logmel_clip(i) for i in range(n_clips)]
    lm = np.stack(lms)
    mean, std = lm.mean((0, 1)), lm.std((0, 1))
    _, sd = refmodels.seeded_cnn(refmodels.MAIN_CFG)
    x = torch.from_numpy(((lm - mean) / std)[:, None].astype(np.float32))
    with torch.no_grad():
        probs = torch.sigmoid(cnn_ref.cnn_avgpooling_forward(sd, x, [2, 2, 2, 1]))
    dt = time.perf_counter() - t0
    return dt, tuple(probs.shape)


def cpu_baseline(n_clips=None):
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    if n_clips is None:
        n_clips = max(8, min(128, 4 * cores))
    workers = min(cores, n_clips)
    cpu_reference_pass(min(2, n_clips), min(2, workers))               # warm caches / imports
    dt, _ = cpu_reference_pass(n_clips, workers)
    return {"value": n_clips * CLIP_SECONDS / 3600.0 / dt, "unit": "audio-hours/sec", "cores": workers,
            "kind": "port",
            "sample": f"{n_clips} x 60 s clips: oracle log-mel (per-frame float64 rFFT, 3-D np.dot mel, dB) in a "
                      f"{workers}-process pool + reference Cnn_AvgPooling(32-64-128-128) eval on torch CPU "
                      f"({cores} threads); {dt:.2f} s"}


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path (oracle port) on this box's host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    n_clips = max(8, min(32, cores))
    workers = min(cores, n_clips)
    for _ in range(min(args.warmup, 1)):
        cpu_reference_pass(min(2, n_clips), min(2, workers))
    t = []
    steps = max(1, min(args.steps, 5))
    for _ in range(steps):
        dt, _ = cpu_reference_pass(n_clips, workers)
        t.append(dt)
    per_step = sum(t) / len(t)
    value = n_clips * CLIP_SECONDS / 3600.0 / per_step
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "audio-hours/sec", "n_gpus": args.gpus,
            "steps": steps, "warmup": min(args.warmup, 1), "ms_per_step": per_step * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64/f32", "data": "synthetic",
            "config": {"workload": f"{n_clips} x 60 s TAU-SED-2019-shaped clips per step (bounded sample of the "
                                   f"{args.clips}-clip GPU workload): log-mel + Cnn_AvgPooling(32-64-128-128) frame "
                                   f"probabilities on host CPU"},
            "cpu_baseline": {"value": value, "unit": "audio-hours/sec", "cores": workers, "kind": "port",
                             "sample": f"{n_clips} clips/step, {workers} processes + {cores} torch threads"},
            "e2e": {"value": value, "unit": "audio-hours/sec", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)



# ------------------------------------------------------------------------------------------------- host <-> device path
def bind_to_gpu_numa(index):
    """Pin this process to the CPUs of the GPU's NUMA node BEFORE any pinned allocation, so that the staging memory of
    the end-to-end path is node-local (first touch).  Returns a short description (or why it was skipped)."""
    try:
        import torch
        pr = torch.cuda.get_device_properties(index)
        bdf = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        base = f"/sys/bus/pci/devices/{bdf}"
        node = int(open(f"{base}/numa_node").read().strip())
        cpus = set()
        for part in open(f"{base}/local_cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if node < 0 or not cpus:
            return f"{bdf}: no NUMA information"
        os.sched_setaffinity(0, cpus)
        return f"{bdf}: NUMA node {node}, {len(cpus)} CPUs"
    except Exception as e:                                               # noqa: BLE001
        return f"skipped ({type(e).__name__})"


def h2d_ceiling(host, dev, chunk_bytes=64 << 20, reps=3):
    """Pure pinned host -> device copy rate (GB/s, this rank) with the chunking of the end-to-end pipeline and every
    rank copying at the same time: the ceiling the host-buffer `e2e` figure can reach on this box."""
    import torch
    from sed_b200 import parallel
    flat = host.view(torch.uint8).reshape(-1)
    stage = [torch.empty(chunk_bytes, dtype=torch.uint8, device=dev) for _ in range(2)]
    st = torch.cuda.Stream(device=dev)
    best = 0.0
    for _ in range(reps):
        torch.cuda.synchronize()
        parallel.barrier()
        t0 = time.perf_counter()
        with torch.cuda.stream(st):
            for i, off in enumerate(range(0, flat.numel(), chunk_bytes)):
                n = min(chunk_bytes, flat.numel() - off)
                stage[i & 1][:n].copy_(flat[off:off + n], non_blocking=True)
        st.synchronize()
        dt = parallel.max_over_ranks(time.perf_counter() - t0, dev)
        best = max(best, flat.numel() / dt / 1e9)
    return best


# ------------------------------------------------------------------------------------------------- other BASELINE configs
M5_FLOP_PER_FRAME = 237_570_560
TRAIN_FLOP_PER_CROP = 3 * 153_281_280   # forward + data gradient + weight gradient of the T = 30 crop


def _timed(fn, steps, warmup, dev, sync_ranks=True):
    """CUDA-event time per call (ms), max over ranks."""
    import torch
    from sed_b200 import parallel
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    if sync_ranks:
        parallel.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        fn()
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / steps
    return parallel.max_over_ranks(ms, dev) if sync_ranks else ms


def bench_config3(model, dev, rank, world, steps, tf_peak):
    """128 clips in total, sharded by clip over the ranks; log-mel images resident in HBM; CNN forward only."""
    import torch
    from sed_b200 import parallel
    total = 128
    lo, hi = parallel.shard_range(total, rank, world)
    n = hi - lo
    g = torch.Generator(device=dev).manual_seed(77 + rank)
    x = torch.randn(max(n, 1), 1, FRAMES, 64, device=dev, generator=g)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)            # > L2: evicts activations between calls

    def fwd():
        if n > 0:
            with torch.no_grad():
                model.logits(x[:n])

    def fwd_flushed():
        flush.zero_()
        fwd()

    ms = _timed(fwd, max(steps, 20), 5, dev)
    ms_fl = _timed(fwd_flushed, 10, 2, dev) - _timed(lambda: flush.zero_(), 10, 2, dev)
    tfl = total * CNN_FLOP_PER_CLIP / (ms * 1e-3) / 1e12
    return {"workload": f"{total} x 60 s clips in total, {n} on rank 0: Cnn_AvgPooling(32-64-128-128) frame probabilities, "
                        f"log-mel input resident in HBM", "clips_total": total, "scaling": "strong",
            "ms_per_step": ms, "ms_per_step_l2_flushed": ms_fl,
            "audio_hours_per_sec": total * CLIP_SECONDS / 3600.0 / (ms * 1e-3),
            "roofline": {"bound": "tensor", "achieved": tfl, "peak": tf_peak * world, "unit": "TFLOP/s",
                         "frac": tfl / (tf_peak * world),
                         "note": "algorithmic FLOPs (965.77 MFLOP/clip); the kernels execute 2 MMAs per product "
                                 "(fp16 activations x fp16 hi+lo weights)"}}


def bench_config5(dev, rank, world, steps, tf_peak):
    import torch
    import refmodels
    from sed_b200 import parallel
    m5, _ = refmodels.seeded_m5()
    m5 = m5.to(dev).eval()
    out = {}
    for total in (128, 1024):
        lo, hi = parallel.shard_range(total, rank, world)
        n = hi - lo
        g = torch.Generator(device=dev).manual_seed(99 + rank)
        x = torch.randn(max(n, 1), 1, 31680, device=dev, generator=g) * 0.1

        def fwd():
            if n > 0:
                with torch.no_grad():
                    m5(x[:n])

        ms = _timed(fwd, max(steps, 20), 5, dev)
        tfl = total * M5_FLOP_PER_FRAME / (ms * 1e-3) / 1e12
        out[f"frames_{total}"] = {"ms_per_step": ms, "frames_per_sec": total / (ms * 1e-3),
                                  "audio_hours_per_sec": total / (ms * 1e-3) * 0.33 / 3600.0,
                                  "roofline": {"bound": "tensor", "achieved": tfl, "peak": tf_peak * world, "unit": "TFLOP/s",
                                               "frac": tfl / (tf_peak * world)}}
    out["workload"] = "M5 (models/waveform_models.py) on raw 0.66 s frames [n, 1, 31680], frames split over the ranks"
    out["scaling"] = "strong"
    return out


def bench_config4(dev, rank, world, steps, tf_peak):
    """One data-parallel training step per GPU: 64 crops x 10 s -> fused log-mel -> native train step (graph)."""
    import torch
    import torch.distributed as dist
    from sed_b200 import parallel
    from sed_b200.dataset.spectogram import preprocess as P
    from sed_b200.models.spectogram_models import Cnn_AvgPooling
    from sed_b200.train import DataParallelTrainer, allreduce_sum_
    from sed_b200.utils.common import WeightedBCE
    import refmodels
    torch.manual_seed(0)
    model = Cnn_AvgPooling(1, model_config=refmodels.MAIN_CFG).to(dev)
    crit = WeightedBCE(recall_factor=5, multi_frame=True)
    g = torch.Generator(device=dev).manual_seed(4321 + rank)
    wave = (torch.randn(64, 480000, device=dev, generator=g) * 0.1).clamp_(-1, 1)
    target = (torch.rand(64, 30, 1, device=dev, generator=g) > 0.8).float()
    mean = torch.full((64,), 18.0, device=dev)
    std = torch.full((64,), 6.0, device=dev)
    tr = DataParallelTrainer(model, crit, lr=1e-6, graph=True)

    def step():
        x = P.waveform_to_log_mel(wave, mean=mean, std=std)[:, None, :30].contiguous()
        return tr.step(x, target)

    ms = _timed(step, max(steps, 20), 5, dev)
    loss = float(step())
    # the split, eagerly (no graph), with events between the stages
    tr2 = DataParallelTrainer(model, crit, lr=1e-6, graph=False)
    names = ("logmel", "forward", "loss", "backward", "allreduce", "update")
    acc = {k: 0.0 for k in names}
    reps = 10
    from sed_b200 import _ext
    import ctypes
    lib = _ext.load()
    for it in range(reps + 3):
        e = [torch.cuda.Event(enable_timing=True) for _ in range(7)]
        e[0].record()
        x = P.waveform_to_log_mel(wave, mean=mean, std=std)[:, None, :30].contiguous()
        e[1].record()
        model.train()
        model._train_forward_native(x)
        e[2].record()
        out = model._train_out
        dl = torch.empty_like(out)
        p = lambda t: ctypes.c_void_p(t.data_ptr())    # noqa: E731
        _ext.check(lib.sedb_bce_with_logits(p(out), p(target), 64, out.shape[1], 30, 1, 5.0, 1.0, p(tr2._loss), p(dl),
                                            _ext.stream_ptr()))
        e[3].record()
        grads = model._train_backward_native(x, dl, model._train_token)
        e[4].record()
        allreduce_sum_(tr2.flat.grad)
        e[5].record()
        tr2._apply_update_dev()
        e[6].record()
        torch.cuda.synchronize()
        if it >= 3:
            for k, a, b in zip(names, e[:-1], e[1:]):
                acc[k] += a.elapsed_time(b) / reps
        del grads
    tfl = world * 64 * TRAIN_FLOP_PER_CROP / (ms * 1e-3) / 1e12
    return {"workload": "per GPU: 64 waveform crops x 10 s -> fused log-mel -> Cnn_AvgPooling(32-64-128-128) train-mode "
                        "forward (batch-stat BN) + WeightedBCE(5) + backward (native tcgen05 dgrad/wgrad) + one NCCL "
                        "all-reduce of the 2.33 MB bucket + fused Adam-amsgrad; whole step replayed as a CUDA graph",
            "crops_per_gpu": 64, "scaling": "weak", "ms_per_step": ms, "steps_per_sec": 1e3 / ms,
            "audio_hours_per_sec": world * 64 * 10.0 / 3600.0 / (ms * 1e-3), "loss": loss,
            "split_ms_eager": acc,
            "roofline": {"bound": "tensor", "achieved": tfl, "peak": tf_peak * world, "unit": "TFLOP/s",
                         "frac": tfl / (tf_peak * world),
                         "note": "algorithmic FLOPs 3 x 153.28 MFLOP per crop; launch/latency-bound at this size"}}


# ------------------------------------------------------------------------------------------------- GPU arm
def run_ours(args):
    import numpy as np
    import torch
    import sed_b200  # noqa: F401
    from sed_b200 import _ext, parallel
    from sed_b200.dataset.spectogram import preprocess as P
    import refmodels

    rank, local_rank, world = parallel.init_process_group("nccl" if int(os.environ.get("WORLD_SIZE", "1")) > 1 else None)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the hot path has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa = bind_to_gpu_numa(local_rank) if not args.no_numa_bind else "disabled"
    lib = _ext.load()
    C = args.clips

    # synthetic inputs, resident in HBM before the timed region (2.95 GB per rank at C=256: far larger than L2)
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    wave = torch.empty(C, CLIP_SAMPLES, device=dev, dtype=torch.float32)
    for i in range(C):
        wave[i] = (torch.randn(CLIP_SAMPLES, device=dev, generator=g) * 0.1).clamp_(-1, 1)
    model, _ = refmodels.seeded_cnn(refmodels.MAIN_CFG)
    model = model.to(dev)
    mean = torch.full((64,), 18.0, device=dev)
    std = torch.full((64,), 6.0, device=dev)

    ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731
    lm_t, cnn_t = [], []

    def step(record):
        e0, e1, e2 = ev(), ev(), ev()
        e0.record()
        x = P.waveform_to_log_mel(wave, mean=mean, std=std)
        e1.record()
        probs = model.logits(x[:, None])
        e2.record()
        if record:
            lm_t.append((e0, e1))
            cnn_t.append((e1, e2))
        return probs

    sampler = ClockSampler(local_rank)
    for _ in range(args.warmup):
        step(False)
    torch.cuda.synchronize()
    parallel.barrier()
    sampler.mark_start()
    launches0 = lib.sedb_launch_count()
    t0, t1 = ev(), ev()
    torch.cuda.synchronize()
    t0.record()
    for _ in range(args.steps):
        probs = step(True)
    t1.record()
    torch.cuda.synchronize()
    parallel.barrier()
    launches = lib.sedb_launch_count() - launches0
    clocks = sampler.stop()
    ms_total = parallel.max_over_ranks(t0.elapsed_time(t1), dev)
    ms_step = ms_total / args.steps
    lm_ms = sum(a.elapsed_time(b) for a, b in lm_t) / len(lm_t)
    cnn_ms = sum(a.elapsed_time(b) for a, b in cnn_t) / len(cnn_t)
    assert probs.shape == (C, 176, 1) and bool(torch.isfinite(probs).all())
    value = world * C * CLIP_SECONDS / 3600.0 / (ms_step * 1e-3)

    # ---- end to end through the C-ABI host-buffer entry point (pinned host memory, H2D + D2H in the timed region)
    e2e_steps = max(1, min(args.steps, 5))
    Ce = min(C, args.e2e_clips)
    if Ce <= 0:
        Ce = 1
    host_wave = torch.empty(Ce, CLIP_SAMPLES, dtype=torch.float32).pin_memory()
    host_wave.copy_(wave[:Ce])
    host_probs = torch.empty(Ce, 176, 1, dtype=torch.float32).pin_memory()
    host_norm = torch.cat([mean, std]).cpu().pin_memory()
    handle = model._native.get(dev, model._native_tensors())

    def e2e_step():
        _ext.check(lib.sedb_sed_host_f32(_ext.context(), handle, ctypes.c_void_p(host_wave.data_ptr()), Ce,
                                         CLIP_SAMPLES, CLIP_SAMPLES, ctypes.c_void_p(host_norm.data_ptr()),
                                         ctypes.c_void_p(host_probs.data_ptr())))

    e2e_step()
    torch.cuda.synchronize()
    parallel.barrier()
    w0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()                                   # synchronises internally (result is on the host)
    e2e_ms = parallel.max_over_ranks((time.perf_counter() - w0) * 1e3 / e2e_steps, dev)
    e2e_ok = bool(torch.allclose(host_probs, probs[:Ce].cpu(), atol=1e-4))
    e2e_value = world * Ce * CLIP_SECONDS / 3600.0 / (e2e_ms * 1e-3)
    h2d_gbs = h2d_ceiling(host_wave, dev)                      # per rank, all ranks copying concurrently
    e2e_gbs = Ce * CLIP_SAMPLES * 4 / (e2e_ms * 1e-3) / 1e9

    # ---- the same clips as 16-bit PCM (the WAV data chunk): not the headline configuration, reported
    # beside it because it halves the bytes per clip in HBM and over PCIe
    pcm16 = None
    if not args.no_pcm16:
        pcm_dev = (wave * 32767.0).round_().to(torch.int16)
        from sed_b200.dataset.spectogram.preprocess import pcm16_to_log_mel
        for _ in range(2):
            pcm16_to_log_mel(pcm_dev, mean, std)
        pa, pb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        pa.record()
        for _ in range(e2e_steps):
            pcm16_to_log_mel(pcm_dev, mean, std)
        pb.record()
        torch.cuda.synchronize()
        pcm_lm_ms = pa.elapsed_time(pb) / e2e_steps
        host_pcm = torch.empty(Ce, CLIP_SAMPLES, dtype=torch.int16).pin_memory()
        host_pcm.copy_(pcm_dev[:Ce])
        del pcm_dev

        def pcm_step():
            _ext.check(lib.sedb_sed_host_pcm16(_ext.context(), handle, ctypes.c_void_p(host_pcm.data_ptr()), Ce,
                                               CLIP_SAMPLES, CLIP_SAMPLES, 1, ctypes.c_void_p(host_norm.data_ptr()),
                                               ctypes.c_void_p(host_probs.data_ptr())))

        pcm_step()
        torch.cuda.synchronize()
        parallel.barrier()
        w0 = time.perf_counter()
        for _ in range(e2e_steps):
            pcm_step()
        pcm_ms = parallel.max_over_ranks((time.perf_counter() - w0) * 1e3 / e2e_steps, dev)
        pcm16 = {"note": "same clips quantised to 16-bit PCM mono (sedb_logmel_pcm16 / sedb_sed_host_pcm16); "
                         "not the headline configuration",
                 "logmel_ms_device_resident": pcm_lm_ms,
                 "e2e": {"value": world * Ce * CLIP_SECONDS / 3600.0 / (pcm_ms * 1e-3), "unit": "audio-hours/sec",
                         "ms_per_step": pcm_ms, "h2d_bytes_per_step": Ce * CLIP_SAMPLES * 2 + 512,
                         "d2h_bytes_per_step": Ce * 176 * 4, "clips_per_step": Ce}}

    # ---- the other BASELINE configurations (every rank takes part: sharding, all-reduce), reported in the same line
    hbm_peak, tf_peak, src = measured_peaks()
    del wave, host_wave
    torch.cuda.empty_cache()
    extra = {}
    if not args.no_configs:
        extra["config3"] = bench_config3(model, dev, rank, world, args.steps, tf_peak)
        extra["config5"] = bench_config5(dev, rank, world, args.steps, tf_peak)
        extra["config4"] = bench_config4(dev, rank, world, args.steps, tf_peak)

    if rank != 0:
        return
    prof = profiled_logmel(C)
    burst = burst_tflops()
    achieved = C * ALGO_BYTES_PER_CLIP / (lm_ms * 1e-3) / 1e9
    tflops = C * FRAMES * LOGMEL_MMA_FLOP_PER_FRAME / (lm_ms * 1e-3) / 1e12
    line = {
        "metric": METRIC, "value": value, "unit": "audio-hours/sec", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16x3-split (fp32 accumulate)" if not lib.sedb_split_is_fp16() else
        "fp16x3-split (fp32 accumulate)", "data": "synthetic",
        "config": {"workload": f"{C} x 60 s TAU-SED-2019-shaped clips per GPU: fused log-mel (win 31680, hop 15840, "
                               f"NFFT 32768, 64 mel) + Cnn_AvgPooling(32-64-128-128) frame probabilities",
                   "clips_per_gpu": C, "l2_policy": "inputs (2.95 GB/GPU) larger than L2",
                   "stage_ms": {"logmel": lm_ms, "cnn": cnn_ms}},
        "roofline": {"kernel": "logmel_fused_kernel", "bound": "hbm", "achieved": achieved, "peak": hbm_peak,
                     "unit": "GB/s", "frac": achieved / hbm_peak, "traffic": prof["traffic"] if prof else None,
                     "traffic_stale": prof["stale"] if prof else None, "traffic_capture": prof["capture"] if prof else None,
                     "peak_source": src, "algorithmic_bytes_per_launch": C * ALGO_BYTES_PER_CLIP, "ms_per_launch": lm_ms},
        "tensor": {"kernel": "logmel_fused_kernel", "unit": "TFLOP/s", "peak_burst": burst,
                   "useful_tflops": tflops / 3.0, "frac_useful_of_burst_peak": tflops / 3.0 / burst,
                   "executed_tflops": tflops, "frac_executed_of_burst_peak": tflops / burst,
                   "pipe_active_pct_ncu": prof["pipe_active_pct"] if prof else None,
                   "pipe_active_stale": prof["stale"] if prof else None,
                   "note": "useful = the factored DFT's GEMM FLOPs once (16.8 MFLOP/frame, both stages); executed = "
                           "x3 for the hi/lo split products; pipe_active = sm__pipe_tensor_cycles_active from the committed "
                           "ncu capture (stale = kernel sources changed since)"},
        "cnn": {"ms": cnn_ms, "algorithmic_tflops": C * CNN_FLOP_PER_CLIP / (cnn_ms * 1e-3) / 1e12},
        "e2e": {"value": e2e_value, "unit": "audio-hours/sec", "h2d_bytes_per_step": Ce * CLIP_SAMPLES * 4 + 512,
                "d2h_bytes_per_step": Ce * 176 * 4, "ms_per_step": e2e_ms, "clips_per_step": Ce,
                "matches_device_path": e2e_ok, "h2d_gbs_per_gpu": e2e_gbs, "h2d_ceiling_gbs_per_gpu": h2d_gbs,
                "frac_of_h2d_ceiling": e2e_gbs / h2d_gbs if h2d_gbs else None, "numa_binding": numa,
                "note": "h2d_ceiling = pure pinned host->device copies of the same buffer in the same 64 MB chunks with "
                        "every rank copying at once; the end-to-end path cannot beat it on this box"},
        "gpu_launches": int(launches),
        "clocks": clocks,
    }
    if pcm16 is not None:
        line["pcm16"] = pcm16
    line.update(extra)
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline()
    emit(line)


_REAL_STDOUT = None


def _claim_stdout():
    """Everything except the final JSON line goes to stderr: libraries (NCCL prints its version banner to fd 1) must
    not pollute the one-line contract."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line):
    out = _REAL_STDOUT if _REAL_STDOUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--clips", type=int, default=256, help="60 s clips per GPU per step")
    ap.add_argument("--e2e-clips", type=int, default=256)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-pcm16", action="store_true", help="skip the 16-bit PCM variant of the measurement")
    ap.add_argument("--no-numa-bind", action="store_true", help="do not pin the process to the GPU's NUMA node")
    ap.add_argument("--no-configs", action="store_true", help="skip the config3 / config4 / config5 sub-measurements")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)
    try:
        import torch.distributed as dist
        if dist.is_initialized():
            dist.destroy_process_group()
    except Exception:
        pass


if __name__ == "__main__":
    main()
