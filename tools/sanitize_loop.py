"""Shapes large enough that every persistent warp runs its main loop several times (double-buffered staging, parked
output rows, next-unit prefetch), for compute-sanitizer racecheck / memcheck on the GPU box:

    compute-sanitizer --tool racecheck --kernel-regex kns=dsb200 python tools/sanitize_loop.py [workload ...]

tools/sanitize_smoke.py covers the ragged edges; at its sizes most warps take one quad / unit and never reach the
steady state of their pipelines."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import diffsptk_b200 as B  # noqa: E402
import diffsptk_b200.functional as F  # noqa: E402
from oracle import np_oracle as O  # noqa: E402


def close(a, b, what, rtol=1e-3, atol=1e-4):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    s = max(1.0, float(np.max(np.abs(b))))
    assert np.allclose(a, b, rtol=rtol, atol=atol * s), f"{what}: max err {np.max(np.abs(a - b)):.3e}"


def main():
    which = set(sys.argv[1:]) or {"stft", "mfcc", "lpc", "mcep", "stftn", "istft", "grad"}
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev).manual_seed(5)
    x = torch.randn(80, 160000, generator=g, device=dev)       # 160 080 frames: 17 quads / 2.2 LPC units per warp
    probe = (0, 41, 79)
    x64 = {b: x[b].cpu().numpy().astype(np.float64) for b in probe}
    with torch.no_grad():
        if "stft" in which:
            P = F.stft(x)
            for b in probe:
                close(P[b].cpu().numpy(), O.stft(x64[b]), f"stft {b}")
        if "mfcc" in which:
            c = F.mfcc_from_waveform(x)
            for b in probe:
                close(c[b].cpu().numpy(), O.mfcc(O.stft(x64[b]), 13, 40, 16000), f"mfcc {b}")
        if "lpc" in which:
            a = F.lpc_from_waveform(x, lpc_order=24)
            a12 = F.lpc_from_waveform(x, lpc_order=12)
            a320 = F.lpc_from_waveform(x, lpc_order=16, frame_length=320, frame_period=160)
            for b in probe:
                close(a[b].cpu().numpy(), O.lpc(O.window(O.frame(x64[b]), None), 24, eps=1e-5), f"lpc {b}", 2e-2, 2e-3)
                close(a12[b].cpu().numpy(), O.lpc(O.window(O.frame(x64[b]), None), 12, eps=1e-5), f"lpc12 {b}", 2e-2, 2e-3)
                close(a320[b].cpu().numpy(), O.lpc(O.window(O.frame(x64[b], 320, 160), None), 16, eps=1e-5),
                      f"lpc320 {b}", 2e-2, 2e-3)
        if "mcep" in which:
            P = F.stft(x[:40])
            mc = B.MelCepstralAnalysis(fft_length=512, cep_order=24, alpha=0.42, n_iter=10).to(dev)(P)
            close(mc[0, :64].cpu().numpy(), O.mcep(O.stft(x64[0])[:64], 24, 0.42, 10), "mcep 0")
        if "stftn" in which:
            for nn, hop in ((1024, 160), (2048, 441)):
                Pn = B.STFT(nn, hop, nn, window="hanning", norm="none", zmean=True).to(dev)(x[:48])
                close(Pn[0].cpu().numpy(), O.stft(x64[0], frame_length=nn, frame_period=hop, fft_length=nn, zmean=True,
                                                  window="hanning", norm="none"), f"stft {nn}")
        if "istft" in which:
            Y = F.stft(x[:48], out_format="complex")
            xr = F.istft(Y, out_length=160000)
            close(xr[0].cpu().numpy(), x64[0], "istft 0")
    if "grad" in which:
        xg = x[:48].clone().requires_grad_(True)
        F.stft(xg).sum().backward()
        assert torch.isfinite(xg.grad).all()
    torch.cuda.synchronize()
    print("sanitize_loop ok:", " ".join(sorted(which)))


if __name__ == "__main__":
    main()
