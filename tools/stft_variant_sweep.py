"""A/B sweep of the stft512 tuning knobs on one B200 (run under gpurun).

    python tools/stft_variant_sweep.py            # parent: runs every combination in a child process
    python tools/stft_variant_sweep.py --child    # child: parity vs the default build + timing, prints one JSON line

The knobs (DSB200_STFT_V, DSB200_STFT_W, DSB200_STFT_STORE) are read once per process, hence the children.
Parity: neither knob changes the arithmetic of a frame, so every variant must reproduce the default variant's
output BIT FOR BIT on a batch with ragged edges (the default itself is pinned by tests/test_gpu_parity.py).
Timing: CUDA events over `steps` launches of BASELINE config 2 (256 x 10 s), two rotating inputs.
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "gpurun_out")


def child():
    import torch

    import diffsptk_b200 as D

    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev).manual_seed(7)
    res = {k[12:]: v for k, v in os.environ.items() if k.startswith("DSB200_STFT_")}
    # parity batch: ragged length (last quad partial), several utterances, all real formats + complex
    outs = {}
    for T in (16000, 16084, 400, 81):
        x = torch.randn(5, T, device=dev, generator=g)
        for fmt in ("power", "magnitude", "db", "log-magnitude", "complex"):
            m = D.STFT(400, 80, 512, out_format=fmt).to(dev)
            with torch.no_grad():
                y = m(x)
            outs[f"{T}_{fmt}"] = torch.view_as_real(y).cpu() if y.is_complex() else y.cpu()
        with torch.no_grad():
            outs[f"{T}_mfcc"] = D.mfcc_from_waveform(x).cpu()
    ref_path = os.path.join(OUT, "sweep_ref.pt")
    if res == {"V": "0"}:
        torch.save(outs, ref_path)
        res["parity"] = "reference"
    else:
        ref = torch.load(ref_path)
        bad = [k for k in outs if not torch.equal(outs[k], ref[k])]
        res["parity"] = "bit-exact" if not bad else "MISMATCH " + ",".join(bad)
    # timing at config 2
    steps = int(os.environ.get("SWEEP_STEPS", "60"))
    xs = [torch.randn(256, 160000, device=dev, generator=g) for _ in range(2)]
    m = D.STFT(400, 80, 512).to(dev)
    with torch.no_grad():
        for i in range(6):
            y = m(xs[i & 1])
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
        ev[0].record()
        for i in range(steps):
            y = m(xs[i & 1])
            ev[i + 1].record()
        torch.cuda.synchronize()
    ts = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(steps))
    res["ms_median"] = ts[len(ts) // 2]
    res["ms_min"] = ts[0]
    res["ms_mean"] = sum(ts) / len(ts)
    # sustained run with NVML sampling: is the kernel running into the board's power cap?
    sustained = int(os.environ.get("SWEEP_SUSTAINED", "3000"))
    if sustained:
        import threading
        import time
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(0)
            mhz, watts, stop = [], [], threading.Event()

            def poll():
                while not stop.is_set():
                    mhz.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                    watts.append(nv.nvmlDeviceGetPowerUsage(h) / 1000.0)
                    time.sleep(0.005)
            th = threading.Thread(target=poll, daemon=True)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            with torch.no_grad():
                for i in range(200):
                    y = m(xs[i & 1])
                torch.cuda.synchronize()
                th.start()
                e0.record()
                for i in range(sustained):
                    y = m(xs[i & 1])
                e1.record()
                torch.cuda.synchronize()
            stop.set()
            th.join(timeout=2)
            k = len(mhz) // 4    # drop the ramp
            res["sustained_ms"] = e0.elapsed_time(e1) / sustained
            res["sm_mhz"] = sorted(mhz[k:])[len(mhz[k:]) // 2] if mhz[k:] else None
            res["watts"] = sum(watts[k:]) / max(1, len(watts[k:]))
            res["power_limit_w"] = nv.nvmlDeviceGetEnforcedPowerLimit(h) / 1000.0
        except Exception as exc:   # NVML missing: timing only
            res["nvml_error"] = repr(exc)[:100]
    print("SWEEP " + json.dumps(res), flush=True)


def main():
    """SWEEP_COMBOS="V=1;V=7,W=16;V=7,W=20": ';'-separated configurations, each a ','-list of DSB200_STFT_<K>=<v>.
    The first configuration must be the parity reference (V=0)."""
    os.makedirs(OUT, exist_ok=True)
    spec = os.environ.get("SWEEP_COMBOS", "V=0;V=1;V=7,W=16;V=7,W=20")
    combos = [dict(kv.split("=") for kv in c.split(",")) for c in spec.split(";")]
    rows = []
    for rep in range(2):   # second pass: the reference and the three fastest again (clocks drift between processes)
        if rep == 1:
            ok = sorted((r for r in rows if "ms_median" in r), key=lambda r: r["ms_median"])[:3]
            combos = [combos[0]] + [r["combo"] for r in ok if r["combo"] != combos[0]]
        for c in combos:
            env = dict(os.environ, **{f"DSB200_STFT_{k}": v for k, v in c.items()})
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--child"], env=env, capture_output=True,
                               text=True, timeout=600)
            line = [ln for ln in r.stdout.splitlines() if ln.startswith("SWEEP ")]
            row = json.loads(line[0][6:]) if line else {"error": (r.stderr or r.stdout)[-400:]}
            row["combo"] = c
            rows.append(row)
            print(row, flush=True)
    with open(os.path.join(OUT, "stft_variant_sweep.json"), "w") as f:
        json.dump(rows, f, indent=1)
    ok = [r for r in rows if r.get("parity") in ("bit-exact", "reference")]
    best = min(ok, key=lambda r: r["ms_median"])
    print("BEST", best)
    with open(os.path.join(OUT, "sweep_best.env"), "w") as f:
        f.write("export " + " ".join(f"DSB200_STFT_{k}={v}" for k, v in best["combo"].items()) + "\n")


if __name__ == "__main__":
    child() if "--child" in sys.argv else main()
