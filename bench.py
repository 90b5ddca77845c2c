#!/usr/bin/env python
"""Benchmark of the frame-rate analysis hot path (BASELINE.json metric: frames/sec at fl=400, fp=80,
n_fft=512; HBM GB/s vs roofline).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload stft|lpc|mfcc|mcep] [--impl reference]

A "step" is one pass of the hot path over one batch of synthetic 16 kHz waveforms that are already
resident in HBM (``value``); ``e2e`` is the same metric through the host-buffer C-ABI pipeline with the
H2D / D2H copies inside the timed region.  One JSON line is printed by rank 0.  ``--impl reference``
times the CPU oracle port (the reference itself is pure Python + torch and cannot travel to the GPU
box) with all host threads on the same workload.
"""

from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

FL, FP, NFFT = 400, 80, 512

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
    "istft": ("Inverse STFT (ifftr -> window -> overlap-add, one kernel): 256 utt x 10 s of complex spectra", 256,
              160000, 2056, 320),
}


def n_frames(T):
    return (T - 1) // FP + 1


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (burst copy)"
        except Exception:
            pass
    return 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"


def ncu_traffic(workload):
    """dram bytes per launch of the dominant kernel from the committed ncu capture (or None)."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        return json.load(open(p)).get(workload)
    except Exception:
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


# ------------------------------------------------------------------------------- CPU oracle leg
_CPU_X = None


def _cpu_task(args):
    workload, lo, hi = args
    import numpy as np
    from oracle import np_oracle as O
    x = _CPU_X[lo:hi]
    if workload == "stft":
        y = O.stft(x)
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


def cpu_oracle_throughput(workload, utterances, T, steps, warmup, budget_s=25.0):
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
    frames = utterances * n_frames(T)
    ctx = mp.get_context("fork")
    with ctx.Pool(min(cores, len(tasks))) as pool:
        times = []
        t_start = time.perf_counter()
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            pool.map(_cpu_task, tasks, chunksize=1)
            dt = time.perf_counter() - t0
            if i >= warmup:
                times.append(dt)
            if time.perf_counter() - t_start > budget_s and len(times) >= 1:
                break
    ms = 1e3 * sum(times) / len(times)
    desc = (f"{utterances} utterances x {T / 16000:g} s ({frames} frames) per step, {len(times)} timed steps, "
            f"numpy oracle port, {min(cores, len(tasks))} worker processes")
    return frames / (ms / 1e3), ms, min(cores, len(tasks)), desc


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg, B, T, rd, wr = WORKLOADS[args.workload]
    value, ms, cores, desc = cpu_oracle_throughput(args.workload, B, T, args.steps, args.warmup, budget_s=150.0)
    line = {
        "impl": "reference", "metric": "frames/sec (fl=400 fp=80 n_fft=512)", "value": value, "unit": "frames/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": cfg, "frame_length": FL, "frame_period": FP, "fft_length": NFFT},
        "cpu_baseline": {"value": value, "unit": "frames/s", "cores": cores, "kind": "port", "sample": desc},
        "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", default="stft", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--no-extras", action="store_true", help="skip the secondary workloads / cpu baseline")
    ap.add_argument("--gather", action="store_true", help="mfcc/lpc workloads: all-gather the features (NCCL)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    if args.impl == "reference":
        return run_reference(args)

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

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist_on = world > 1
    if dist_on:
        import datetime
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=90))

    cfg, B, T, rd, wr = WORKLOADS[args.workload]
    N = n_frames(T)
    frames_per_step = B * N
    xs, step = make_step(args.workload, B, T, dev)
    if args.gather and args.workload in ("mfcc", "lpc") and dist_on:
        import diffsptk_b200.functional as F
        fn = (lambda x: F.mfcc_from_waveform(x)) if args.workload == "mfcc" else (lambda x: F.lpc_from_waveform(x, lpc_order=24))
        step = lambda i: sharded_features(fn, xs[i & 1], n_chunks=4)  # noqa: E731

    sampler = ClockSampler(local) if rank == 0 else None
    timed_steps(step, 2, args.warmup, dist_on)  # extra warm-up pass: allocator, twiddle cache, clocks
    if sampler:
        sampler.start()
    l0 = _native.launch_count()
    total_ms, per = timed_steps(step, args.steps, args.warmup, dist_on)
    launches = _native.launch_count() - l0
    t = torch.tensor([total_ms], device=dev, dtype=torch.float64)
    if dist_on:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
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
        src = xs[0].cpu().pin_memory()
        h2d = src.numel() * src.element_size()
        xin = torch.empty_like(xs[0])
        probe = step(0)
        yh = torch.empty(probe.shape, dtype=probe.dtype).pin_memory()
        d2h = yh.numel() * yh.element_size()

        def e2e_step():
            xin.copy_(src, non_blocking=True)
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
    t = torch.tensor([e2e_ms], device=dev, dtype=torch.float64)
    if dist_on:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * frames_per_step / (float(t.item()) / 1e3)

    if rank != 0:
        if dist_on:
            dist.destroy_process_group()
        return

    peak, peak_src = hbm_peak()
    algo_bytes = frames_per_step * (rd + wr)
    kernel_ms = statistics.mean(per)
    achieved = algo_bytes / (kernel_ms / 1e3) / 1e9
    line = {
        "metric": "frames/sec (fl=400 fp=80 n_fft=512)", "value": value, "unit": "frames/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": cfg, "frame_length": FL, "frame_period": FP, "fft_length": NFFT,
                   "utterances_per_gpu": B, "samples_per_utterance": T, "frames_per_step_per_gpu": frames_per_step,
                   "l2": "per-step inputs + outputs exceed the 126 MB L2; two rotating input buffers",
                   "collective": "all-gather of features (NCCL)" if (args.gather and dist_on) else "none (batch-sharded)"},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": ncu_traffic(args.workload), "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": algo_bytes, "kernel_ms": kernel_ms},
        "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": float(t.item()), "steps": e2e_steps},
        "gpu_launches": int(launches),
        "clocks": clocks,
    }
    if world == 1 and not args.no_extras:
        v, ms, cores, desc = cpu_oracle_throughput(args.workload, min(B, 256), T, 3, 1)
        line["cpu_baseline"] = {"value": v, "unit": "frames/s", "cores": cores, "kind": "port", "sample": desc}
        extras = {}
        for wl in ("lpc", "mfcc", "mcep", "istft", "stft_grad", "delta"):
            if wl == args.workload:
                continue
            try:
                _, Bw, Tw, rdw, wrw = WORKLOADS[wl]
                del xs, step
                torch.cuda.empty_cache()
                xs, step = make_step(wl, Bw, Tw, dev)
                tot, perw = timed_steps(step, 5, 3, False)
                kms = statistics.mean(perw)
                fw = Bw * n_frames(Tw)
                extras[wl] = {"frames_per_s": fw / (kms / 1e3), "ms_per_step": kms,
                              "hbm_frac": fw * (rdw + wrw) / (kms / 1e3) / 1e9 / peak}
            except Exception as e:  # an extra must never take the headline down
                extras[wl] = {"error": repr(e)[:200]}
        line["extra_workloads"] = extras
        if args.workload == "stft":
            # COMPARATOR ONLY (never on the product path): the route the reference's modules take on a GPU --
            # pad + unfold + window multiply + torch.fft.rfft (cuFFT) + abs/square/add as separate torch kernels
            # (stft.py:237-241) -- on the same device and the same batch.
            try:
                import torch.nn.functional as TF
                del xs, step
                torch.cuda.empty_cache()
                xc = torch.randn(B, T, device=dev)
                wc = tables.make_window(FL, device=dev, dtype=torch.float32)

                def composite(i):
                    fr = TF.pad(xc, (FL // 2, (FL - 1) // 2)).unfold(-1, FL, FP) * wc
                    return torch.fft.rfft(TF.pad(fr, (0, NFFT - FL))).abs().square() + 1e-9
                _, perc = timed_steps(composite, 5, 3, False)
                cms = statistics.mean(perc)
                line["comparators"] = {"torch_composite_same_gpu": {
                    "frames_per_s": frames_per_step / (cms / 1e3), "ms_per_step": cms,
                    "what": "F.pad + unfold + window + torch.fft.rfft + abs().square() + eps, fp32, same batch"}}
            except Exception as e:
                line["comparators"] = {"torch_composite_same_gpu": {"error": repr(e)[:200]}}
    if args.workload == "stft" and clocks and clocks.get("sm_mhz"):
        # the kernel sits at the HBM / FP32 ridge: report the FP32 side as well
        slots = 8320.0   # FMA-pipe lane-slots per frame: 1 040 pipe cycles x 32 lanes per quad of 4 frames
        peak_slots = 148 * 128 * float(clocks["sm_mhz"]) * 1e6
        line["roofline"]["fp32_pipe"] = {"lane_slots_per_frame": slots, "frac": frames_per_step / (kernel_ms / 1e3) * slots / peak_slots,
                                         "peak": "148 SMs x 128 lanes x measured SM clock"}
    emit(line)
    if dist_on:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
