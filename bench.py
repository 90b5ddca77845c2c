#!/usr/bin/env python
"""Benchmark of the frame-rate analysis hot path (BASELINE.json metric: frames/sec at fl=400, fp=80,
n_fft=512; HBM GB/s vs roofline).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload stft|lpc|mfcc|mcep|...] [--impl reference]

A "step" is one pass of the hot path over one batch of synthetic 16 kHz waveforms that are already
resident in HBM (``value``); ``e2e`` is the same metric through the host-buffer C-ABI pipeline with the
H2D / D2H copies inside the timed region.  One JSON line is printed by rank 0.

``--impl reference`` times the reference's own CPU implementation of the same workload on the host cores:
the unmodified ``diffsptk`` modules when the package is importable from ``$DIFFSPTK_REFERENCE_ROOT`` or
``baseline/_ref`` (``cpu_baseline.kind = "reference"``, torch intra-op threads = all cores), else the numpy
oracle port with one process per core (``kind = "port"``).  Both arms print the same ``config`` object.
"""

from __future__ import annotations

import argparse
import hashlib
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

FL, FP, NFFT = 400, 80, 512
METRIC = "frames/sec (fl=400 fp=80 n_fft=512)"
CFG1 = "Frame+Window+STFT fl=400 fp=80 n_fft=512 on 1x16 kHz 10 s utterance, CPU reference"

WORKLOADS = {
    # name: (BASELINE.json config, utterances per GPU, samples per utterance, bytes read / written per frame)
    "stft": ("Batched STFT->Spectrum: 256 utterances x 10 s @16 kHz, n_fft=512", 256, 160000, 320, 1028),
    "lpc": ("LPC pipeline (Frame->Window->acorr->levdur, M=24): 1024 utt x 5 s", 1024, 80000, 320, 100),
    "mfcc": ("MFCC (fbank->DCT, 40 mel / 13 cep) from the waveform: 1024 utt x 10 s per GPU", 1024, 160000, 320, 52),
    "mcep": ("MelCepstralAnalysis (n_iter=10, M=24, alpha=0.42) over STFT power, 512 utt x 10 s", 512, 160000,
             1028, 100),
    "stft_grad": ("STFT power forward + backward (d/dwaveform, native adjoint kernel): 256 utt x 10 s", 256, 160000,
                  320 + 1028 + 320, 1028 + 320),
    "delta": ("Delta + delta-delta features (widths 2, 2) of 13-dim cepstra: 4096 utt x 2001 frames", 4096, 160000,
              52, 156),
    "lpc2par": ("LPC -> PARCOR (step-down recursion, M=24) on the LPC rows of config 3: 1024 utt x 5 s", 1024, 80000,
                100, 100),
    "lpc2lsp": ("LPC -> line spectral pairs (M=24) on the LPC rows of config 3: 1024 utt x 5 s", 1024, 80000,
                100, 100),
    "stft1024": ("STFT power at frame length = fft length = 1024, hop 160 (the reference's CREPE front end, pitch.py:"
                 "245-256): 256 utt x 10 s", 256, 160000, 640, 2052),
    "stft2048": ("STFT power at frame length = fft length = 2048, hop 441 (yingram.py:97): 256 utt x 10 s", 256, 160000,
                 1764, 4100),
    "istft": ("Inverse STFT (ifftr -> window -> overlap-add, one kernel): 256 utt x 10 s of complex spectra", 256,
              160000, 2056, 320),
}

# utterances per step of the CPU arms (a bounded sample of the workload; frames/s is batch-size-flat on the CPU)
CPU_SAMPLE = {"stft1024": 32, "stft2048": 32, "stft": 64, "lpc": 64, "mfcc": 64, "mcep": 4, "stft_grad": 32, "delta": 256, "lpc2par": 64,
              "lpc2lsp": 1, "istft": 32}


_HOPS = {"stft1024": 160, "stft2048": 441}


def n_frames(T, hop=FP):
    return (T - 1) // hop + 1


def workload_config(workload, collective="none (batch-sharded)"):
    """The ``config`` object of the JSON line -- built by ONE function so that both arms print identical keys."""
    cfg, B, T, _, _ = WORKLOADS[workload]
    return {"workload": cfg, "frame_length": FL, "frame_period": FP, "fft_length": NFFT,
            "utterances_per_gpu": B, "samples_per_utterance": T,
            "frames_per_step_per_gpu": B * n_frames(T, _HOPS.get(workload, FP)),
            "l2": "per-step inputs + outputs exceed the 126 MB L2; two rotating input buffers",
            "collective": collective}


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (burst copy)"
        except Exception:
            pass
    return 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"


def kernel_source_digest(files):
    h = hashlib.sha256()
    for f in files:
        with open(os.path.join(ROOT, "diffsptk_b200", "csrc", f), "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()[:16]


def ncu_record(workload):
    """Per-launch ncu numbers of the dominant kernel (profiles/traffic.json), valid only for the kernel sources they
    were captured from: every entry carries the digest of those sources and is dropped when they changed."""
    try:
        rec = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(workload)
        if isinstance(rec, dict) and rec.get("source_digest") == kernel_source_digest(rec.get("sources", [])):
            return rec
    except Exception:
        pass
    return None


class ClockSampler:
    """Polls SM clock / throttle reasons through NVML while the GPU is under load."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml as nv
            nv.nvmlInit()
            self.nv, self.h = nv, nv.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {}
        for attr in dir(nv):
            if attr.startswith("nvmlClocksEventReason") or attr.startswith("nvmlClocksThrottleReason"):
                v = getattr(nv, attr)
                if isinstance(v, int) and v not in (0,):
                    names[v] = attr.replace("nvmlClocksEventReason", "").replace("nvmlClocksThrottleReason", "")
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit and "GpuIdle" not in name and "None" not in name and "All" not in name:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.002)

    def start(self):
        if self.nv is not None:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()

    def stop(self):
        self._stop.set()
        if self._thread is not None:
            self._thread.join(timeout=2)
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None,
                "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ------------------------------------------------------------------- CPU arm 1: the reference's own modules
def load_reference_package():
    """The unmodified reference (pure Python on torch), importable from $DIFFSPTK_REFERENCE_ROOT or baseline/_ref
    (git-ignored install that travels to the GPU box).  Returns (module, root) or (None, None)."""
    import types
    for root in (os.environ.get("DIFFSPTK_REFERENCE_ROOT"), os.path.join(ROOT, "baseline", "_ref")):
        if not root or not os.path.isdir(os.path.join(root, "diffsptk")):
            continue
        if "soundfile" not in sys.modules:   # the reference's only missing import-time dependency (utils/public.py:18)
            try:
                import soundfile  # noqa: F401
            except Exception:
                sys.modules["soundfile"] = types.ModuleType("soundfile")
        if root not in sys.path:
            sys.path.insert(0, root)
        try:
            import diffsptk
            return diffsptk, root
        except Exception:
            continue
    return None, None


def reference_step(ref, workload, utterances, T, device="cpu", seed=1234):
    """(frames per step, step()) running the reference's stock modules on `device` (CPU arm, or the same-GPU
    comparator when device is a CUDA device)."""
    import torch
    g = torch.Generator().manual_seed(seed)
    N = n_frames(T)
    x = torch.randn(utterances, T, generator=g).to(device)
    kw = dict(device=device)
    stft = ref.STFT(FL, FP, NFFT, **kw)
    if workload == "stft":
        return utterances * N, lambda: stft(x)
    if workload in ("stft1024", "stft2048"):
        nn, hop = (1024, 160) if workload == "stft1024" else (2048, 441)
        big = ref.STFT(nn, hop, nn, window="hanning", norm="none", **kw)
        return utterances * n_frames(T, hop), lambda: big(x)
    if workload == "lpc":
        fr, wi, lp = ref.Frame(FL, FP), ref.Window(FL, **kw), ref.LPC(FL, 24, **kw)
        return utterances * N, lambda: lp(wi(fr(x)))
    if workload == "mfcc":
        mf = ref.MFCC(fft_length=NFFT, mfcc_order=13, n_channel=40, sample_rate=16000, **kw)
        return utterances * N, lambda: mf(stft(x))
    if workload == "mcep":
        mc = ref.MelCepstralAnalysis(fft_length=NFFT, cep_order=24, alpha=0.42, n_iter=10, **kw)
        with torch.no_grad():
            P = stft(x)
        return utterances * N, lambda: mc(P)
    if workload == "stft_grad":
        gy = torch.randn(utterances, N, NFFT // 2 + 1, generator=g).to(device)

        def fwd_bwd():
            with torch.enable_grad():
                xr = x.detach().requires_grad_(True)
                stft(xr).backward(gy)
            return xr.grad
        return utterances * N, fwd_bwd
    if workload == "istft":
        st = ref.STFT(FL, FP, NFFT, out_format="complex", **kw)
        it = ref.ISTFT(FL, FP, NFFT, **kw)
        with torch.no_grad():
            Y = st(x)
        return utterances * N, lambda: it(Y, out_length=T)
    if workload == "delta":
        d = ref.Delta([2, 2], True, **kw)   # same regression windows as the native arm (Delta([2, 2], True))
        c = torch.randn(utterances, N, 13, generator=g).to(device)
        return utterances * N, lambda: d(c)
    k = torch.empty(utterances, N, 25).uniform_(-0.9, 0.9, generator=g).to(device)
    a = ref.ParcorCoefficientsToLinearPredictiveCoefficients(24)(k)
    a[..., 0].abs_().add_(0.1)
    if workload == "lpc2par":
        m = ref.LinearPredictiveCoefficientsToParcorCoefficients(24)
    else:
        m = ref.LinearPredictiveCoefficientsToLineSpectralPairs(24, **kw)
    return utterances * N, lambda: m(a)


def time_cpu(step, steps, warmup, budget_s):
    times, t_start = [], time.perf_counter()
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        step()
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
        if time.perf_counter() - t_start > budget_s and times:
            break
    return 1e3 * sum(times) / len(times), len(times)


def reference_cpu_throughput(ref, workload, T, steps, warmup, budget_s):
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    utt = CPU_SAMPLE[workload]
    with torch.no_grad():
        frames, step = reference_step(ref, workload, utt, T)
        ms, n = time_cpu(step, steps, warmup, budget_s)
    desc = (f"{utt} utterances x {T / 16000:g} s ({frames} frames) per step, {n} timed steps, the reference's own "
            f"modules (diffsptk {getattr(ref, '__version__', '?')}, torch {torch.__version__} CPU, fp32, "
            f"torch.set_num_threads({cores}))")
    return frames / (ms / 1e3), ms, cores, desc


# ---------------------------------------------------------------------------- CPU arm 2: the oracle port
_CPU_X = None


def _cpu_task(args):
    workload, lo, hi = args
    import numpy as np
    from oracle import np_oracle as O
    x = _CPU_X[lo:hi]
    if workload == "stft":
        y = O.stft(x)
    elif workload in ("stft1024", "stft2048"):
        nn = 1024 if workload == "stft1024" else 2048
        y = O.stft(x, frame_length=nn, frame_period=_HOPS[workload], fft_length=nn, window="hanning", norm="none")
    elif workload == "lpc":
        y = O.lpc(O.window(O.frame(x, FL, FP), None), 24)
    elif workload == "mfcc":
        y = O.mfcc(O.stft(x), 13, 40, 16000)
    elif workload == "istft":  # x holds complex spectra
        y = O.istft(x)
    elif workload == "delta":  # x holds 13-dim features
        y = O.delta(x, [2, 2], True)
    elif workload == "lpc2par":  # x holds LPC rows
        y = O.lpc2par(x)
    elif workload == "lpc2lsp":
        y = O.lpc2lsp(x)
    elif workload == "stft_grad":  # the oracle has no autograd: forward only (a lower bound on the CPU cost)
        y = O.stft(x)
    else:  # mcep: x holds power spectra
        y = O.mcep(x, 24, 0.42, 10)
    return float(np.sum(y[..., :1]))


def port_cpu_throughput(workload, utterances, T, steps, warmup, budget_s=25.0):
    """frames/s of the numpy oracle with one process per host thread (fork pool, inputs shared
    copy-on-write, only a checksum returns).  Returns (frames_per_s, ms_per_step, cores, sample_desc)."""
    global _CPU_X
    import multiprocessing as mp

    import numpy as np
    cores = os.cpu_count() or 1
    rng = np.random.default_rng(1234)
    if workload == "mcep":
        from oracle import np_oracle as O
        utterances = min(utterances, max(1, cores // 16))  # ~1.5k frames/s/core: keep a step to seconds
        _CPU_X = O.stft(rng.standard_normal((utterances, T)).astype(np.float32))
    elif workload == "istft":
        from oracle import np_oracle as O
        _CPU_X = O.stft(rng.standard_normal((utterances, T)).astype(np.float32), out_format="complex")
    elif workload == "delta":
        _CPU_X = rng.standard_normal((utterances, n_frames(T), 13)).astype(np.float32)
    elif workload in ("lpc2par", "lpc2lsp"):
        from oracle import np_oracle as O
        if workload == "lpc2lsp":
            utterances = min(utterances, max(1, cores // 8))   # an eigenvalue problem per row: keep a step to seconds
        k = rng.uniform(-0.9, 0.9, (utterances, n_frames(T), 25)).astype(np.float32)
        _CPU_X = O.par2lpc(k)
    else:
        _CPU_X = rng.standard_normal((utterances, T)).astype(np.float32)
    per = max(1, utterances // (cores * 2))
    tasks = [(workload, lo, min(utterances, lo + per)) for lo in range(0, utterances, per)]
    frames = utterances * n_frames(T, _HOPS.get(workload, FP))
    ctx = mp.get_context("fork")
    with ctx.Pool(min(cores, len(tasks))) as pool:
        ms, n = time_cpu(lambda: pool.map(_cpu_task, tasks, chunksize=1), steps, warmup, budget_s)
    desc = (f"{utterances} utterances x {T / 16000:g} s ({frames} frames) per step, {n} timed steps, "
            f"numpy oracle port, {min(cores, len(tasks))} worker processes")
    return frames / (ms / 1e3), ms, min(cores, len(tasks)), desc


def run_reference(args):
    """The reference arm: rank 0 alone computes and prints; no CUDA is touched."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    cfg, B, T, rd, wr = WORKLOADS[args.workload]
    ref, root = (None, None) if args.port else load_reference_package()
    extra = {}
    if ref is not None:
        value, ms, cores, desc = reference_cpu_throughput(ref, args.workload, T, args.steps, args.warmup, args.budget)
        kind = "reference"
        if args.workload == "stft":   # BASELINE.json config 1: one 10 s utterance through the reference on the CPU
            import torch
            with torch.no_grad():
                frames1, step1 = reference_step(ref, "stft", 1, 160000)
                ms1, n1 = time_cpu(step1, 20, 3, 10.0)
            extra["configs"] = {"cfg1": {"workload": CFG1, "frames_per_s": frames1 / (ms1 / 1e3), "ms_per_step": ms1,
                                         "steps": n1, "cores": cores}}
    else:
        value, ms, cores, desc = port_cpu_throughput(args.workload, min(B, 256), T, args.steps, args.warmup,
                                                     budget_s=args.budget)
        kind = "port"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "frames/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.workload),
        "cpu_baseline": {"value": value, "unit": "frames/s", "cores": cores, "kind": kind, "sample": desc},
        "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    line.update(extra)
    print(json.dumps(line), flush=True)


def cpu_baseline_subprocess(workload, budget_s=20.0):
    """Run the CPU arm in a FRESH process (before this process creates a CUDA context): forking a worker pool out of
    a process that holds a CUDA context with eight visible devices and ~700 MB of pinned memory took > 3 minutes on
    the 8-GPU node in round 1 (SCALE_r01: N=1 wall 220 s vs 13 s on the 1-GPU box)."""
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "LOCAL_RANK", "WORLD_SIZE")}
    try:
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--workload", workload,
                            "--steps", "3", "--warmup", "1", "--budget", str(budget_s)],
                           capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
        line = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1])
        out = dict(line["cpu_baseline"])
        if "configs" in line:
            out["cfg1"] = line["configs"]["cfg1"]
        return out
    except Exception as e:  # the baseline must never take the headline down
        return {"error": repr(e)[:200]}


# ------------------------------------------------------------------------------------ GPU legs
def make_step(workload, B, T, dev):
    """Returns (inputs[2], step(i) -> output tensor) for the resident-in-HBM measurement."""
    import torch

    import diffsptk_b200 as D
    import diffsptk_b200.functional as F
    g = torch.Generator(device=dev).manual_seed(1234 + int(os.environ.get("RANK", "0")))
    if workload == "mcep":
        stft = D.STFT(FL, FP, NFFT).to(dev)
        mcep = D.MelCepstralAnalysis(fft_length=NFFT, cep_order=24, alpha=0.42, n_iter=10).to(dev)
        with torch.no_grad():
            xs = [stft(torch.randn(B, T, generator=g, device=dev)) for _ in range(2)]
        return xs, lambda i: mcep(xs[i & 1])
    if workload == "delta":
        dl = D.Delta([2, 2], True).to(dev)
        xs = [torch.randn(B, n_frames(T), 13, generator=g, device=dev) for _ in range(2)]
        return xs, lambda i: dl(xs[i & 1])
    if workload in ("lpc2par", "lpc2lsp"):
        with torch.no_grad():   # stable LPC rows: step-up recursion from random PARCOR coefficients
            xs = [F.par2lpc(torch.empty(B, n_frames(T), 25, device=dev).uniform_(-0.9, 0.9, generator=g))
                  for _ in range(2)]
            for x in xs:
                x[..., 0].abs_().add_(0.1)   # positive gain
        fn = F.lpc2par if workload == "lpc2par" else F.lpc2lsp
        return xs, lambda i: fn(xs[i & 1])
    if workload == "istft":
        stft = D.STFT(FL, FP, NFFT, out_format="complex").to(dev)
        istft = D.ISTFT(FL, FP, NFFT).to(dev)
        with torch.no_grad():
            xs = [stft(torch.randn(B, T, generator=g, device=dev)) for _ in range(2)]
        return xs, lambda i: istft(xs[i & 1], T)
    xs = [torch.randn(B, T, generator=g, device=dev) for _ in range(2)]
    if workload == "stft_grad":
        m = D.STFT(FL, FP, NFFT).to(dev)
        gy = torch.randn(B, n_frames(T), NFFT // 2 + 1, generator=g, device=dev)

        def grad_step(i):
            with torch.enable_grad():
                x = xs[i & 1].detach().requires_grad_(True)
                m(x).backward(gy)
            return x.grad
        return xs, grad_step
    if workload == "stft":
        m = D.STFT(FL, FP, NFFT).to(dev)
        return xs, lambda i: m(xs[i & 1])
    if workload in ("stft1024", "stft2048"):
        nn = 1024 if workload == "stft1024" else 2048
        m = D.STFT(nn, _HOPS[workload], nn, window="hanning", norm="none").to(dev)
        return xs, lambda i: m(xs[i & 1])
    if workload == "lpc":
        return xs, lambda i: F.lpc_from_waveform(xs[i & 1], lpc_order=24)
    return xs, lambda i: F.mfcc_from_waveform(xs[i & 1])


def timed_steps(step, steps, warmup, dist_on):
    import torch
    import torch.distributed as dist
    with torch.no_grad():
        for i in range(warmup):
            step(i)
        torch.cuda.synchronize()
        if dist_on:
            dist.barrier()
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
        ev[0].record()
        for i in range(steps):
            step(i)
            ev[i + 1].record()
        torch.cuda.synchronize()
        if dist_on:
            dist.barrier()
    per = [ev[i].elapsed_time(ev[i + 1]) for i in range(steps)]
    return ev[0].elapsed_time(ev[-1]), per


def max_over_ranks(v, dev, dist_on):
    import torch
    import torch.distributed as dist
    t = torch.tensor([float(v)], device=dev, dtype=torch.float64)
    if dist_on:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def oracle_spot_check(workload, x_dev, y_dev, utterances=(0, -1)):
    """Recompute a few utterances of the LAST timed step with the numpy oracle (checker only; outside the timed
    region) at the reference's tolerance (rtol 1e-4, atol 1e-6 scaled by the output's magnitude)."""
    import numpy as np
    from oracle import np_oracle as O
    fns = {"stft": lambda x: O.stft(x),
           "mfcc": lambda x: O.mfcc(O.stft(x), 13, 40, 16000),
           "lpc": None}   # ill-conditioned rows need the conditioned criterion of tests/helpers.py: tests only
    fn = fns.get(workload)
    if fn is None:
        return None
    idx = sorted({u % x_dev.shape[0] for u in utterances})
    x = x_dev[idx].double().cpu().numpy()
    got = y_dev[idx].cpu().numpy()
    want = fn(x)
    err = np.abs(got - want)
    tol = 1e-6 * max(1.0, float(np.max(np.abs(want)))) + 1e-4 * np.abs(want)
    return {"ok": bool(np.all(err <= tol)), "utterances": idx, "max_abs_err": float(err.max()),
            "max_abs_value": float(np.max(np.abs(want))), "tolerance": "rtol 1e-4, atol 1e-6 x max|want| vs the numpy oracle (float64)"}


def raw_copy_ms(xh, yh, dev, reps=3):
    """Plain cudaMemcpyAsync of the same pinned buffers on two streams (H2D and D2H overlapped), no kernel: the floor
    of the end-to-end step that the host side (PCIe root, host DRAM, IOMMU) allows on this box at this rank count."""
    import torch
    xd = torch.empty(xh.shape, dtype=xh.dtype, device=dev)
    yd = torch.empty(yh.shape, dtype=yh.dtype, device=dev)
    s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    out = {}
    for name, do_in, do_out in (("h2d", True, False), ("d2h", False, True), ("both", True, True)):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            if do_in:
                with torch.cuda.stream(s1):
                    xd.copy_(xh, non_blocking=True)
            if do_out:
                with torch.cuda.stream(s2):
                    yh.copy_(yd, non_blocking=True)
        torch.cuda.synchronize()
        out[name] = 1e3 * (time.perf_counter() - t0) / reps
    return out


def gpu_composite_comparators(workloads, dev, peak):
    """COMPARATORS ONLY (never on the product path): the reference's own modules moved to the same GPU
    (`diffsptk.X(...).to('cuda')`: torch kernels + cuFFT + cuSOLVER/cuBLAS), same batch, same device."""
    import torch
    out = {}
    ref, root = load_reference_package()
    for wl in workloads:
        _, B, T, rd, wr = WORKLOADS[wl]
        try:
            torch.cuda.empty_cache()
            if ref is not None:
                with torch.no_grad():
                    frames, stepf = reference_step(ref, wl, B, T, device=dev)
                what = f"the reference's modules on the same GPU ({wl}: diffsptk on cuda, fp32, same batch)"
            elif wl == "stft":
                import torch.nn.functional as TF

                from diffsptk_b200 import tables
                xc = torch.randn(B, T, device=dev)
                wc = tables.make_window(FL, device=dev, dtype=torch.float32)
                frames = B * n_frames(T)

                def stepf():
                    fr = TF.pad(xc, (FL // 2, (FL - 1) // 2)).unfold(-1, FL, FP) * wc
                    return torch.fft.rfft(TF.pad(fr, (0, NFFT - FL))).abs().square() + 1e-9
                what = "F.pad + unfold + window + torch.fft.rfft + abs().square() + eps, fp32, same batch"
            else:
                continue
            _, per = timed_steps(lambda i: stepf(), 3, 2, False)
            ms = statistics.mean(per)
            out[wl] = {"frames_per_s": frames / (ms / 1e3), "ms_per_step": ms, "what": what}
            del stepf
        except Exception as e:
            out[wl] = {"error": repr(e)[:200]}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", default="stft", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--no-extras", action="store_true", help="skip the secondary workloads / cpu baseline")
    ap.add_argument("--gather", default="none", choices=["none", "nccl", "fused"],
                    help="mfcc/lpc workloads at N > 1: all-gather the features (chunked in-place NCCL, or fused "
                         "into the kernel's stores over NVLink peer memory)")
    ap.add_argument("--port", action="store_true", help="reference arm: force the numpy oracle port")
    ap.add_argument("--budget", type=float, default=150.0, help="reference arm: wall-clock budget in seconds")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    if args.impl == "reference":
        return run_reference(args)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    extras_on = not args.no_extras

    # CPU baseline first, in a fresh process, before this one owns a CUDA context (rank 0 at N = 1 only)
    cpu_baseline = cpu_baseline_subprocess(args.workload) if (world == 1 and extras_on) else None

    # Libraries (NCCL prints its version line at init) write to file descriptor 1 behind Python's back: keep the
    # contract "rank 0 prints ONE JSON line" by pointing fd 1 at stderr for the run and printing the line through
    # a saved copy of the real stdout.
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        sys.stdout.flush()
        os.write(real_stdout, (json.dumps(obj) + "\n").encode())

    import torch
    import torch.distributed as dist

    from diffsptk_b200 import _native, ops, tables
    from diffsptk_b200.distributed import sharded_features

    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist_on = world > 1
    if dist_on:
        import datetime
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=120))

    cfg, B, T, rd, wr = WORKLOADS[args.workload]
    N = n_frames(T, _HOPS.get(args.workload, FP))
    frames_per_step = B * N
    xs, step = make_step(args.workload, B, T, dev)
    collective = "none (batch-sharded)"
    if args.gather != "none" and args.workload in ("mfcc", "lpc") and dist_on:
        import diffsptk_b200.functional as F
        fn = (lambda x: F.mfcc_from_waveform(x)) if args.workload == "mfcc" else (lambda x: F.lpc_from_waveform(x, lpc_order=24))
        step = lambda i: sharded_features(fn, xs[i & 1], n_chunks=4)  # noqa: E731
        collective = "all-gather of features (chunked in-place NCCL all_gather_into_tensor)"

    sampler = ClockSampler(local) if rank == 0 else None
    timed_steps(step, 2, args.warmup, dist_on)  # extra warm-up pass: allocator, twiddle cache, clocks
    if sampler:
        sampler.start()
    l0 = _native.launch_count()
    total_ms, per = timed_steps(step, args.steps, args.warmup, dist_on)
    launches = _native.launch_count() - l0
    with torch.no_grad():
        check = oracle_spot_check(args.workload, xs[(args.steps - 1) & 1], step(args.steps - 1)) if rank == 0 else None
    total_ms = max_over_ranks(total_ms, dev, dist_on)
    # Keep every GPU under the same load a little longer so that NVML sees it.  ALL ranks run the SAME
    # number of extra steps (derived from the max-reduced time): a step may contain a collective.
    n_extra = max(8, min(4000, int(400.0 / max(total_ms / args.steps, 1e-3))))
    with torch.no_grad():
        for i in range(n_extra):
            step(i)
        torch.cuda.synchronize()
    clocks = sampler.stop() if sampler else None
    ms_per_step = total_ms / args.steps
    value = world * frames_per_step / (ms_per_step / 1e3)

    # ---- end to end through the host-buffer C-ABI pipeline (stft) or pinned copies + op (others)
    e2e_steps = max(2, min(args.steps, 8))
    if args.workload == "stft":
        w = tables.make_window(FL, device=dev, dtype=torch.float32)
        pipe = ops.HostStftPipeline(w, T, FP, NFFT, chunk_utterances=16)
        xh = torch.randn(B, T).pin_memory()
        yh = torch.empty(pipe.out_shape(B), dtype=torch.float32).pin_memory()
        h2d, d2h = xh.numel() * 4, yh.numel() * 4

        def e2e_step():
            pipe(xh, yh)
    else:
        xh = xs[0].cpu().pin_memory()
        h2d = xh.numel() * xh.element_size()
        xin = torch.empty_like(xs[0])
        probe = step(0)
        yh = torch.empty(probe.shape, dtype=probe.dtype).pin_memory()
        d2h = yh.numel() * yh.element_size()

        def e2e_step():
            xin.copy_(xh, non_blocking=True)
            xs[0] = xin
            yh.copy_(step(0), non_blocking=True)
            torch.cuda.synchronize()
    with torch.no_grad():
        for _ in range(2):
            e2e_step()
        torch.cuda.synchronize()
        if dist_on:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_step()
        torch.cuda.synchronize()
        e2e_ms = 1e3 * (time.perf_counter() - t0) / e2e_steps
    e2e_ms = max_over_ranks(e2e_ms, dev, dist_on)
    e2e_value = world * frames_per_step / (e2e_ms / 1e3)
    # the raw copies of the same pinned buffers, all ranks at once (is the end-to-end step host-limited?)
    if dist_on:
        dist.barrier()
    raw = raw_copy_ms(xh, yh, dev)
    raw = {k: max_over_ranks(v, dev, dist_on) for k, v in raw.items()}
    del xh, yh

    peak, peak_src = hbm_peak()
    algo_bytes = frames_per_step * (rd + wr)
    kernel_ms = statistics.mean(per)
    achieved = algo_bytes / (kernel_ms / 1e3) / 1e9
    rec = ncu_record(args.workload)
    line = {
        "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.workload, collective),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": rec.get("dram_bytes_per_launch") if rec else None, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": algo_bytes, "kernel_ms": kernel_ms,
                     "traffic_source": (rec.get("capture") if rec else
                                        "no ncu capture of the current kernel sources (profiles/traffic.json)")},
        "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": e2e_ms, "steps": e2e_steps,
                "raw_copy_ms": raw,
                "raw_copy_note": "plain cudaMemcpyAsync of the same pinned buffers (H2D, D2H, both overlapped), max over "
                                 "ranks, all ranks copying at once: the host-side floor of ms_per_step on this box"},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "check": check,
    }
    if cpu_baseline is not None:
        line["cpu_baseline"] = cpu_baseline
    if rec and rec.get("fma_lane_slots_per_frame") and clocks and clocks.get("sm_mhz"):
        # the kernel sits at the HBM / FP32 ridge: report the FP32 side as well; the
        # lane-slot count comes from the ncu instruction mix of the same sources (profiles/traffic.json)
        slots = float(rec["fma_lane_slots_per_frame"])
        peak_slots = 148 * 128 * float(clocks["sm_mhz"]) * 1e6
        line["roofline"]["fp32_pipe"] = {"lane_slots_per_frame": slots,
                                         "frac": frames_per_step / (kernel_ms / 1e3) * slots / peak_slots,
                                         "peak": "148 SMs x 128 lanes x measured SM clock"}

    if extras_on:
        # ---- secondary workloads: every rank runs them (batch-sharded), times are the max over ranks
        extras = {}
        del xs, step
        names = ["lpc", "mfcc", "mcep"] + (["istft", "stft_grad", "delta"] if world == 1 else [])
        for wl in names:
            if wl == args.workload:
                continue
            try:
                _, Bw, Tw, rdw, wrw = WORKLOADS[wl]
                torch.cuda.empty_cache()
                xw, stepw = make_step(wl, Bw, Tw, dev)
                _, perw = timed_steps(stepw, 5, 3, dist_on)
                kms = max_over_ranks(statistics.mean(perw), dev, dist_on)
                fw = Bw * n_frames(Tw)
                extras[wl] = {"frames_per_s": world * fw / (kms / 1e3), "ms_per_step": kms,
                              "hbm_frac": fw * (rdw + wrw) / (kms / 1e3) / 1e9 / peak}
                del xw, stepw
            except Exception as e:  # an extra must never take the headline down
                extras[wl] = {"error": repr(e)[:200]}
        if dist_on:
            # BASELINE.json config 5: MFCC on 1024 utterances per GPU, features all-gathered over NVLink so that
            # every rank holds the [world x 1024, 2000, 13] tensor
            import diffsptk_b200.functional as F
            _, Bw, Tw, rdw, wrw = WORKLOADS["mfcc"]
            fw = Bw * n_frames(Tw)
            base = (extras.get("mfcc") or {}).get("ms_per_step")
            torch.cuda.empty_cache()
            xw, _ = make_step("mfcc", Bw, Tw, dev)
            recv_bytes = (world - 1) * fw * wrw

            def gather_entry(kms, what):
                e = {"frames_per_s": world * fw / (kms / 1e3), "ms_per_step": kms, "what": what,
                     "gathered_bytes_per_gpu": recv_bytes}
                if base:
                    e["efficiency_vs_no_gather"] = base / kms
                    if kms > base:
                        e["exposed_gather_ms"] = kms - base
                        e["nvlink_GBps_if_serial"] = recv_bytes / ((kms - base) / 1e3) / 1e9
                return e
            try:
                stepw = lambda i: sharded_features(lambda x: F.mfcc_from_waveform(x), xw[i & 1], n_chunks=4)  # noqa: E731
                _, perw = timed_steps(stepw, 5, 3, True)
                extras["mfcc_gather_nccl"] = gather_entry(
                    max_over_ranks(statistics.mean(perw), dev, True),
                    "4 utterance chunks, in-place all_gather_into_tensor of chunk k overlapped with the kernel of chunk k+1")
            except Exception as e:
                extras["mfcc_gather_nccl"] = {"error": repr(e)[:200]}
            try:
                from diffsptk_b200.distributed import FusedGatherMfcc
                fg = FusedGatherMfcc(Bw, Tw, device=dev)
                ok = max_over_ranks(0.0 if fg.available else 1.0, dev, True) == 0.0   # all ranks or none
                if ok:
                    stepw = lambda i: fg(xw[i & 1])  # noqa: E731
                    _, perw = timed_steps(stepw, 5, 3, True)
                    extras["mfcc_gather_fused"] = gather_entry(
                        max_over_ranks(statistics.mean(perw), dev, True),
                        "ONE kernel: the MFCC epilogue stores every feature row into all ranks' output tensors "
                        f"({fg.mode}) + one cross-rank barrier; no NCCL call on the data path")
                    want = sharded_features(lambda x: F.mfcc_from_waveform(x), xw[0], n_chunks=1)
                    extras["mfcc_gather_fused"]["equals_nccl_gather"] = bool(torch.equal(fg(xw[0]), want))
                else:
                    extras["mfcc_gather_fused"] = {"unavailable": fg.reason}
            except Exception as e:
                extras["mfcc_gather_fused"] = {"error": repr(e)[:200]}
            del xw
        line["extra_workloads"] = extras
        if world == 1 and rank == 0:
            which = [w for w in ("stft", "lpc", "mfcc", "mcep")]
            line["comparators"] = {"reference_modules_same_gpu": gpu_composite_comparators(which, dev, peak)}
            own = {"stft": line["ms_per_step"] if args.workload == "stft" else None}
            for wl in ("lpc", "mfcc", "mcep"):
                own[wl] = (extras.get(wl) or {}).get("ms_per_step")
            for wl, c in line["comparators"]["reference_modules_same_gpu"].items():
                if own.get(wl) and "ms_per_step" in c:
                    c["speedup_of_this_repo"] = c["ms_per_step"] / own[wl]
    if rank == 0:
        emit(line)
    if dist_on:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
