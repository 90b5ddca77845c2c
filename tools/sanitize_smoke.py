"""Small-shape pass over every specialised kernel, meant to run under compute-sanitizer (GPU box):

    compute-sanitizer --tool memcheck|racecheck|synccheck|initcheck --kernel-regex kns=dsb200 \
        python tools/sanitize_smoke.py

Ragged shapes on purpose: partial quads / octets / tiles, utterance ends inside a staged span, unaligned batch
offsets.  Results are compared against the generic kernels / the oracle so that the run also fails on wrong data.
"""
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
    dev = torch.device("cuda", 0)
    rng = np.random.default_rng(0)
    for (Bn, T) in ((3, 1237), (2, 4000), (5, 333), (1, 81)):
        x = rng.standard_normal((Bn, T)).astype(np.float32)
        xd = torch.from_numpy(x).to(dev)
        x64 = x.astype(np.float64)
        with torch.no_grad():
            for fmt in ("power", "db", "complex"):
                P = B.STFT(400, 80, 512, out_format=fmt).to(dev)(xd)
                want = O.stft(x64, out_format=fmt)
                if fmt == "complex":
                    close(torch.view_as_real(P).cpu().numpy(), np.stack([want.real, want.imag], -1), f"stft {fmt} {Bn}x{T}")
                else:
                    close(P.cpu().numpy(), want, f"stft {fmt} {Bn}x{T}")
            Pw = B.STFT(400, 80, 512).to(dev)(xd)
            close(F.mfcc_from_waveform(xd, out_format="ycE").cpu().numpy(),
                  O.mfcc(O.stft(x64), 13, 40, 16000, out_format="ycE"), f"mfcc_wave {Bn}x{T}")
            a = F.lpc_from_waveform(xd, lpc_order=24)          # lag-pair kernel, 16 warps
            close(a.cpu().numpy(), O.lpc(O.window(O.frame(x64), None), 24, eps=1e-5), f"lpc_wave {Bn}x{T}", 2e-2, 2e-3)
            a = F.lpc_from_waveform(xd, lpc_order=12)          # lag-pair kernel with the guarded recursion, 12 warps
            close(a.cpu().numpy(), O.lpc(O.window(O.frame(x64), None), 12, eps=1e-5), f"lpc_wave M=12 {Bn}x{T}", 2e-2, 2e-3)
            a = F.lpc_from_waveform(xd, lpc_order=16, frame_length=320, frame_period=160)   # frame-pair kernel
            close(a.cpu().numpy(), O.lpc(O.window(O.frame(x64, 320, 160), None), 16, eps=1e-5),
                  f"lpc_wave fl=320 {Bn}x{T}", 2e-2, 2e-3)
            for nn, hop in ((1024, 160), (2048, 441)):         # shared-memory FFT kernel
                if T >= 333:
                    Pn = B.STFT(nn, hop, nn, window="hanning", norm="none", zmean=(nn == 1024)).to(dev)(xd)
                    close(Pn.cpu().numpy(), O.stft(x64, frame_length=nn, frame_period=hop, fft_length=nn, zmean=(nn == 1024),
                                                   window="hanning", norm="none"),
                          f"stft {nn} {Bn}x{T}")
            Pz = B.STFT(400, 80, 512, zmean=True, relative_floor=-60.0).to(dev)(xd)
            close(Pz.cpu().numpy(), O.stft(x64, zmean=True, relative_floor=-60.0), f"stft zmean+floor {Bn}x{T}")
            mc = B.MelCepstralAnalysis(fft_length=512, cep_order=24, alpha=0.42, n_iter=10).to(dev)(Pw)
            close(mc.cpu().numpy(), O.mcep(O.stft(x64), 24, 0.42, 10), f"mcep {Bn}x{T}")
            Y = B.STFT(400, 80, 512, out_format="complex").to(dev)(xd)
            xr = B.ISTFT(400, 80, 512).to(dev)(Y, T)
            close(xr.cpu().numpy(), x, f"istft {Bn}x{T}")
        xg = xd.clone().requires_grad_(True)
        B.STFT(400, 80, 512).to(dev)(xg).sum().backward()
        assert torch.isfinite(xg.grad).all()
    torch.cuda.synchronize()
    print("sanitize_smoke ok")


if __name__ == "__main__":
    main()
