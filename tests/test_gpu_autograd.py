"""Gradients of the differentiable (spectral) ops against torch autograd of a plain composite.

The composite below is a float64 torch restatement used ONLY as the gradient checker (pad + unfold +
window + torch.fft.rfft + formatter, i.e. what the reference's modules do, stft.py:237-241); the product
path never uses torch.fft.  float64 kernels must match it tightly, float32 kernels to float32 accuracy.
"""

import numpy as np
import pytest
import torch
import torch.nn.functional as TF

pytestmark = pytest.mark.gpu


def dev():
    return torch.device("cuda", 0)


def composite_stft(x, w, fl, fp, nfft, center, zmean, mode, eps, rf, fmt):
    pad = (fl // 2, (fl - 1) // 2) if center else (0, fl - 1)
    xp = TF.pad(x.unsqueeze(0), pad, mode=mode).squeeze(0) if mode != "constant" else TF.pad(x, pad)
    f = xp.unfold(-1, fl, fp)
    if zmean:
        f = f - f.mean(-1, keepdim=True)
    g = f * w
    g = TF.pad(g, (0, nfft - fl)) if nfft >= fl else g[..., :nfft]
    X = torch.fft.rfft(g, n=nfft)
    if fmt == "complex":
        return X
    s = X.abs().square() + eps
    if rf is not None:
        s = torch.maximum(s, s.amax(-1, keepdim=True) * 10 ** (rf / 10))
    return {"db": lambda v: 10 * torch.log10(v), "log-magnitude": lambda v: 0.5 * torch.log(v),
            "magnitude": torch.sqrt, "power": lambda v: v}[fmt](s)


CASES = [
    dict(fl=400, fp=80, nfft=512, fmt="power"),
    dict(fl=400, fp=80, nfft=512, fmt="complex"),
    dict(fl=400, fp=80, nfft=512, fmt="db", eps=1e-3),
    dict(fl=400, fp=80, nfft=512, fmt="magnitude", eps=1e-3, rf=-20.0),
    dict(fl=12, fp=10, nfft=16, fmt="log-magnitude", eps=1e-2, center=False, zmean=True),
    dict(fl=40, fp=10, nfft=48, fmt="power", mode="reflect"),
    dict(fl=40, fp=10, nfft=32, fmt="power", mode="circular", zmean=True),
    dict(fl=30, fp=7, nfft=64, fmt="complex", mode="replicate"),
]


@pytest.mark.parametrize("prec", ["f64", "f32"])
@pytest.mark.parametrize("case", CASES, ids=lambda c: "-".join(f"{k}{v}" for k, v in c.items()))
def test_stft_gradients(case, prec):
    import diffsptk_b200 as B
    dt = torch.float64 if prec == "f64" else torch.float32
    fl, fp, nfft, fmt = case["fl"], case["fp"], case["nfft"], case["fmt"]
    center, zmean, mode = case.get("center", True), case.get("zmean", False), case.get("mode", "constant")
    eps, rf = case.get("eps", 1e-9), case.get("rf")
    g = torch.Generator().manual_seed(3)
    x0 = torch.randn(2, 700, generator=g, dtype=torch.float64)
    mod = B.STFT(fl, fp, nfft, center=center, zmean=zmean, mode=mode, eps=eps, relative_floor=rf, out_format=fmt,
                 window="hamming", norm="none", learnable=["window"], dtype=dt).to(dev())
    x = x0.to(dev(), dt).requires_grad_(True)
    y = mod(x)
    wgt = torch.randn(y.shape, generator=g, dtype=torch.float64).to(dev()) if not y.is_complex() else \
        torch.complex(torch.randn(y.shape, generator=g, dtype=torch.float64),
                      torch.randn(y.shape, generator=g, dtype=torch.float64)).to(dev())
    loss = (y.to(wgt.dtype) * wgt).real.sum() if y.is_complex() else (y.double() * wgt).sum()
    loss.backward()
    gx, gw = x.grad.double().cpu(), mod.window.window.grad.double().cpu()

    xr = x0.to(dev()).requires_grad_(True)
    wr = mod.window.window.detach().double().requires_grad_(True)
    yr = composite_stft(xr, wr, fl, fp, nfft, center, zmean, mode, eps, rf, fmt)
    lr = (yr * wgt).real.sum() if yr.is_complex() else (yr * wgt).sum()
    lr.backward()
    tol = dict(rtol=1e-8, atol=1e-9) if prec == "f64" else dict(rtol=2e-3, atol=2e-3)
    scale = max(1.0, float(xr.grad.abs().max()))
    np.testing.assert_allclose(gx.numpy() / scale, xr.grad.cpu().numpy() / scale, **tol)
    wscale = max(1.0, float(wr.grad.abs().max()))
    np.testing.assert_allclose(gw.numpy() / wscale, wr.grad.cpu().numpy() / wscale, **tol)


def test_leaf_op_gradients():
    """frame / window / fftr / spec / freqt / dct, float64, against torch composites."""
    import diffsptk_b200 as B
    import diffsptk_b200.functional as F
    d = dev()
    g = torch.Generator().manual_seed(5)
    x = torch.randn(3, 333, generator=g, dtype=torch.float64).to(d).requires_grad_(True)

    def grad_of(fn, *inputs):
        outs = fn(*inputs)
        w = torch.randn(outs.shape, generator=g, dtype=torch.float64).to(d)
        if outs.is_complex():
            w = torch.complex(w, torch.randn(outs.shape, generator=g, dtype=torch.float64).to(d))
            loss = (outs * w).real.sum()
        else:
            loss = (outs * w).sum()
        return torch.autograd.grad(loss, inputs, allow_unused=True), w

    # frame (+zmean, reflect): adjoint of pad + unfold
    (ga,), w = grad_of(lambda t: F.frame(t, 50, 13, zmean=True, mode="reflect"), x)
    xp = TF.pad(x.unsqueeze(0), (25, 24), mode="reflect").squeeze(0).unfold(-1, 50, 13)
    ref = torch.autograd.grad(((xp - xp.mean(-1, keepdim=True)) * w).sum(), x)[0]
    assert torch.allclose(ga, ref, rtol=1e-10, atol=1e-11)
    # window with a learnable table
    fr = torch.randn(4, 6, 20, generator=g, dtype=torch.float64).to(d).requires_grad_(True)
    win = B.Window(20, 32, window="hanning", norm="power", learnable=True, dtype=torch.float64).to(d)
    y = win(fr)
    w = torch.randn(y.shape, generator=g, dtype=torch.float64).to(d)
    gfr, gwin = torch.autograd.grad((y * w).sum(), (fr, win.window))
    assert torch.allclose(gfr, w[..., :20] * win.window.detach(), rtol=1e-12, atol=1e-13)
    assert torch.allclose(gwin, (w[..., :20] * fr.detach()).reshape(-1, 20).sum(0), rtol=1e-10, atol=1e-11)
    # fftr, every output format, odd input length shorter than the FFT
    v = torch.randn(5, 13, generator=g, dtype=torch.float64).to(d).requires_grad_(True)
    for fmt in ("complex", "real", "imaginary", "amplitude", "power"):
        (ga,), w = grad_of(lambda t: F.fftr(t, 16, fmt), v)
        X = torch.fft.rfft(v, n=16)
        out = {"complex": X, "real": X.real, "imaginary": X.imag, "amplitude": X.abs(), "power": X.abs().square()}[fmt]
        ref = torch.autograd.grad((out * w).real.sum() if out.is_complex() else (out * w).sum(), v)[0]
        assert torch.allclose(ga, ref, rtol=1e-9, atol=1e-10), fmt
    # spec (numerator), relative floor + dB
    (ga,), w = grad_of(lambda t: F.spec(t, fft_length=16, eps=1e-2, relative_floor=-10.0, out_format="db"), v)
    s = torch.fft.rfft(v, n=16).abs().square() + 1e-2
    s = torch.maximum(s, s.amax(-1, keepdim=True) * 10 ** (-1.0))
    ref = torch.autograd.grad((10 * torch.log10(s) * w).sum(), v)[0]
    assert torch.allclose(ga, ref, rtol=1e-8, atol=1e-9)
    # freqt and dct: x @ A
    c = torch.randn(7, 20, generator=g, dtype=torch.float64).to(d).requires_grad_(True)
    fq = B.FrequencyTransform(19, 29, 0.1, dtype=torch.float64).to(d)
    (ga,), w = grad_of(fq, c)
    assert torch.allclose(ga, w @ fq.A.t(), rtol=1e-10, atol=1e-11)
    dc = B.DCT(20, dtype=torch.float64).to(d)
    (ga,), w = grad_of(dc, c)
    assert torch.allclose(ga, w @ dc.W.t(), rtol=1e-10, atol=1e-11)


def test_gradient_flows_through_a_pipeline():
    """A small training step through STFT (fused kernel + native backward) moves the input towards a target."""
    import diffsptk_b200 as B
    d = dev()
    torch.manual_seed(0)
    stft = B.STFT(400, 80, 512, out_format="log-magnitude", eps=1e-5).to(d)
    target = stft(torch.randn(2, 4000, device=d)).detach()
    x = torch.randn(2, 4000, device=d, requires_grad=True)
    opt = torch.optim.Adam([x], lr=0.05)
    losses = []
    for _ in range(25):
        opt.zero_grad()
        loss = (stft(x) - target).square().mean()
        loss.backward()
        opt.step()
        losses.append(float(loss))
    assert np.isfinite(losses).all() and losses[-1] < 0.6 * losses[0]


# ---------------------------------------------------------------------------------------------------------
# Feature ops behind the spectrum.  Composites restate the reference forward passes in torch float64
# (fbank.py:305-330, mfcc.py:243-256, acorr.py:112-121, levdur.py:113-127) and are differentiated by torch.
def _vjp(fn, inputs, gen):
    out = fn(*inputs)
    outs = out if isinstance(out, (tuple, list)) else (out,)
    ws = [torch.randn(o.shape, generator=gen, dtype=torch.float64).to(o.device) for o in outs]
    loss = sum((o.double() * w).sum() for o, w in zip(outs, ws))
    return torch.autograd.grad(loss, inputs), ws


def _ref_vjp(fn, inputs, ws):
    out = fn(*inputs)
    outs = out if isinstance(out, (tuple, list)) else (out,)
    return torch.autograd.grad(sum((o * w).sum() for o, w in zip(outs, ws)), inputs)


def composite_fbank(x, Hm, floor, gamma, use_power):
    y = x if use_power else torch.sqrt(x)
    y = torch.clip(y @ Hm, min=floor)
    y = torch.log(y) if gamma == 0 else (torch.pow(y, gamma) - 1) / gamma
    E = (2 * x[..., 1:-1]).sum(-1) + x[..., 0] + x[..., -1]
    return y, torch.log(E / (2 * (x.size(-1) - 1))).unsqueeze(-1)


def composite_levdur(r, eps):
    M = r.size(-1) - 1
    idx = (torch.arange(M)[:, None] - torch.arange(M)[None, :]).abs().to(r.device)
    R = r[..., :-1][..., idx] + eps * torch.eye(M, dtype=r.dtype, device=r.device)
    a = torch.linalg.solve(R, -r[..., 1:].unsqueeze(-1)).squeeze(-1)
    K = torch.sqrt((r[..., 1:] * a).sum(-1, keepdim=True) + r[..., :1])
    return torch.cat((K, a), -1)


def composite_acorr(x, M, fmt):
    L = x.size(-1)
    r = torch.stack([(x[..., : L - k] * x[..., k:]).sum(-1) for k in range(M + 1)], -1)
    if fmt == "normalized":
        r = r / r[..., :1]
    elif fmt == "biased":
        r = r / L
    elif fmt == "unbiased":
        r = r / torch.arange(L, L - M - 1, -1, device=x.device)
    return r


@pytest.mark.parametrize("gamma,use_power", [(0.0, False), (-0.5, True), (0.3, False)])
def test_fbank_and_mfcc_gradients(gamma, use_power):
    import diffsptk_b200 as B
    d = dev()
    g = torch.Generator().manual_seed(11)
    x = (torch.rand(3, 5, 33, generator=g, dtype=torch.float64) + 0.05).to(d).requires_grad_(True)
    fb = B.MelFilterBankAnalysis(fft_length=64, n_channel=10, sample_rate=8000, floor=0.2, gamma=gamma,
                                 use_power=use_power, out_format="y,E", dtype=torch.float64).to(d)
    (gx,), ws = _vjp(fb, (x,), g)
    (ref,) = _ref_vjp(lambda t: composite_fbank(t, fb.H, 0.2, gamma, use_power), (x,), ws)
    assert torch.allclose(gx, ref, rtol=1e-9, atol=1e-11)
    if use_power:
        return
    for fmt in ("y", "yE", "yc", "ycE"):
        mf = B.MFCC(fft_length=64, mfcc_order=6, n_channel=10, sample_rate=8000, lifter=5, floor=0.2, gamma=gamma,
                    out_format=fmt, dtype=torch.float64).to(d)

        def ref_mfcc(t):
            y, E = composite_fbank(t, mf.fbank.H, 0.2, gamma, False)
            c = (y @ mf.dct.W)[..., :7] * mf.liftering_vector
            c0, c = c[..., :1], c[..., 1:]
            return {"y": c, "yE": torch.cat((c, E), -1), "yc": torch.cat((c, c0), -1),
                    "ycE": torch.cat((c, c0, E), -1)}[fmt]
        (gx,), ws = _vjp(mf, (x,), g)
        (ref,) = _ref_vjp(ref_mfcc, (x,), ws)
        assert torch.allclose(gx, ref, rtol=1e-9, atol=1e-11), fmt


@pytest.mark.parametrize("prec", ["f64", "f32"])
def test_lpc_chain_gradients(prec):
    import diffsptk_b200 as B
    import diffsptk_b200.functional as F
    d = dev()
    dt = torch.float64 if prec == "f64" else torch.float32
    g = torch.Generator().manual_seed(13)
    tol = dict(rtol=1e-7, atol=1e-9) if prec == "f64" else dict(rtol=5e-3, atol=5e-3)

    def close(a, b, what):
        scale = max(1.0, float(b.abs().max()))
        assert torch.allclose(a.double() / scale, b / scale, **tol), (what, float((a.double() - b).abs().max()), scale)

    x64 = torch.randn(4, 6, 50, generator=g, dtype=torch.float64).to(d)
    x = x64.to(dt).requires_grad_(True)
    xr = x64.clone().requires_grad_(True)
    for fmt in ("naive", "normalized", "biased", "unbiased"):
        (gx,), ws = _vjp(lambda t: F.acorr(t, 7, fmt), (x,), g)
        (ref,) = _ref_vjp(lambda t: composite_acorr(t, 7, fmt), (xr,), ws)
        close(gx, ref, f"acorr {fmt}")
    r64 = composite_acorr(x64, 7, "naive").detach()
    r = r64.to(dt).requires_grad_(True)
    rr = r64.clone().requires_grad_(True)
    (gr,), ws = _vjp(lambda t: F.levdur(t, eps=1e-3), (r,), g)
    (ref,) = _ref_vjp(lambda t: composite_levdur(t, 1e-3), (rr,), ws)
    close(gr, ref, "levdur")
    (gx,), ws = _vjp(lambda t: F.lpc(t, 7, eps=1e-3), (x,), g)
    (ref,) = _ref_vjp(lambda t: composite_levdur(composite_acorr(t, 7, "naive"), 1e-3), (xr,), ws)
    close(gx, ref, "lpc")
    # module path + order 0 and 1 corner cases
    for M in (0, 1):
        lpc = B.LPC(50, M, eps=1e-3, dtype=dt).to(d)
        (gx,), ws = _vjp(lpc, (x,), g)
        (ref,) = _ref_vjp(lambda t: composite_levdur(composite_acorr(t, M, "naive"), 1e-3), (xr,), ws)
        close(gx, ref, f"LPC module order {M}")


def test_fused_waveform_pipelines_are_differentiable():
    """lpc_from_waveform / mfcc_from_waveform (fused forward kernels) against frame->window->... composites."""
    import diffsptk_b200 as B
    import diffsptk_b200.functional as F
    d = dev()
    g = torch.Generator().manual_seed(17)
    x64 = torch.randn(2, 1500, generator=g, dtype=torch.float64).to(d)
    win = B.Window(400, window="hamming", norm="power", dtype=torch.float64).to(d).window.double()

    def frames(t):
        return TF.pad(t, (200, 199)).unfold(-1, 400, 80) * win

    for dt, tol in ((torch.float64, dict(rtol=1e-7, atol=1e-9)), (torch.float32, dict(rtol=5e-3, atol=5e-3))):
        x = x64.to(dt).requires_grad_(True)
        xr = x64.clone().requires_grad_(True)
        (gx,), ws = _vjp(lambda t: F.lpc_from_waveform(t, lpc_order=8, window="hamming", eps=1e-4), (x,), g)
        (ref,) = _ref_vjp(lambda t: composite_levdur(composite_acorr(frames(t), 8, "naive"), 1e-4), (xr,), ws)
        scale = float(ref.abs().max())
        assert torch.allclose(gx.double() / scale, ref / scale, **tol), ("lpc_wave", dt)

        mf = B.MFCC(fft_length=512, mfcc_order=12, n_channel=20, sample_rate=16000, lifter=22, out_format="ycE",
                    dtype=torch.float64).to(d)

        def ref_mfcc(t):
            P = torch.fft.rfft(frames(t), n=512).abs().square() + 1e-6
            y, E = composite_fbank(P, mf.fbank.H, 1e-5, 0.0, False)
            c = (y @ mf.dct.W)[..., :13] * mf.liftering_vector
            return torch.cat((c[..., 1:], c[..., :1], E), -1)
        (gx,), ws = _vjp(lambda t: F.mfcc_from_waveform(t, mfcc_order=12, n_channel=20, lifter=22, window="hamming",
                                                        eps=1e-6, out_format="ycE"), (x,), g)
        (ref,) = _ref_vjp(ref_mfcc, (xr,), ws)
        scale = float(ref.abs().max())
        assert torch.allclose(gx.double() / scale, ref / scale, **tol), ("mfcc_wave", dt)


@pytest.mark.parametrize("fmt", ["power", "db", "log-magnitude", "magnitude"])
def test_fused_stft_backward_kernel(fmt):
    """stft512_bwd.cu (one kernel, no atomics) against the general adjoint kernel and torch autograd of the composite."""
    import os

    import diffsptk_b200.functional as F
    from diffsptk_b200 import _native
    d = dev()
    g = torch.Generator().manual_seed(23)
    for T, kw in ((9000, dict()), (1234, dict(frame_length=320, frame_period=160, window="hamming")),
                  (4001, dict(frame_length=512, frame_period=128, center=False, window="hanning", norm="none")),
                  (700, dict(frame_length=400, frame_period=80)), (90, dict(frame_length=100, frame_period=50))):
        x64 = torch.randn(3, T, generator=g, dtype=torch.float64)
        fl, fp = kw.get("frame_length", 400), kw.get("frame_period", 80)

        def run(x):
            y = F.stft(x, fft_length=512, eps=1e-3, out_format=fmt, **kw)
            w = torch.randn(y.shape, generator=torch.Generator().manual_seed(1), dtype=torch.float64).to(d)
            (gx,) = torch.autograd.grad((y.double() * w).sum(), x)
            return gx, w
        x = x64.to(d, torch.float32).requires_grad_(True)
        F.stft(x.detach(), fft_length=512, **kw)          # warm the twiddle cache
        n0 = _native.launch_count()
        fast, w = run(x)
        assert _native.launch_count() - n0 == 2, "one forward and one backward kernel"
        os.environ["DSB200_STFT_BWD_GENERIC"] = "1"
        try:
            slow, _ = run(x)
        finally:
            del os.environ["DSB200_STFT_BWD_GENERIC"]
        xr = x64.to(d).requires_grad_(True)
        win = F.window(torch.ones(fl, dtype=torch.float64, device=d), None, window=kw.get("window", "blackman"),
                       norm=kw.get("norm", "power"))
        yr = composite_stft(xr, win, fl, fp, 512, kw.get("center", True), False, "constant", 1e-3, None, fmt)
        (ref,) = torch.autograd.grad((yr * w).sum(), xr)
        scale = float(ref.abs().max())
        assert float((fast.double() - ref).abs().max()) < 2e-3 * scale, (fmt, T, kw)
        assert float((fast - slow).abs().max()) < 2e-4 * scale, (fmt, T, kw)


def test_learnable_filter_bank_gradients():
    """learnable=True turns H into a Parameter (fbank.py:112): its gradient is amp^T (g dy/dz), x's comes from the
    native kernel; the MFCC module chains both."""
    import diffsptk_b200 as B
    d = dev()
    g = torch.Generator().manual_seed(29)
    x = (torch.rand(4, 6, 33, generator=g, dtype=torch.float64) + 0.05).to(d).requires_grad_(True)
    fb = B.MelFilterBankAnalysis(fft_length=64, n_channel=10, sample_rate=8000, floor=0.2, out_format="yE",
                                 learnable=True, dtype=torch.float64).to(d)
    assert isinstance(fb.H, torch.nn.Parameter)
    (gx, gH), ws = _vjp(lambda t, Hm: fb(t), (x, fb.H), g)
    Hr = fb.H.detach().clone().requires_grad_(True)
    rx, rH = _ref_vjp(lambda t, Hm: torch.cat(composite_fbank(t, Hm, 0.2, 0.0, False), -1), (x, Hr), ws)
    assert torch.allclose(gx, rx, rtol=1e-9, atol=1e-11) and torch.allclose(gH, rH, rtol=1e-9, atol=1e-11)

    mf = B.MFCC(fft_length=64, mfcc_order=6, n_channel=10, sample_rate=8000, lifter=5, floor=0.2, out_format="yc",
                learnable=True, dtype=torch.float64).to(d)
    (gx, gH), ws = _vjp(lambda t, Hm: mf(t), (x, mf.fbank.H), g)

    def ref(t, Hm):
        y, _ = composite_fbank(t, Hm, 0.2, 0.0, False)
        c = (y @ mf.dct.W)[..., :7] * mf.liftering_vector
        return torch.cat((c[..., 1:], c[..., :1]), -1)
    Hr = mf.fbank.H.detach().clone().requires_grad_(True)
    rx, rH = _ref_vjp(ref, (x, Hr), ws)
    assert torch.allclose(gx, rx, rtol=1e-9, atol=1e-11) and torch.allclose(gH, rH, rtol=1e-9, atol=1e-11)
    # an optimiser step on a learnable front end: window and filter bank both receive gradients
    front = torch.nn.Sequential(B.STFT(400, 80, 512, learnable=["window"]),
                                B.MFCC(fft_length=512, mfcc_order=12, n_channel=24, sample_rate=16000,
                                       learnable=True)).to(d)
    wav = torch.randn(2, 3000, device=d)
    front(wav).square().mean().backward()
    grads = [p.grad for p in front.parameters()]
    assert len(grads) == 2 and all(gr is not None and bool(torch.isfinite(gr).all()) and float(gr.abs().max()) > 0
                                   for gr in grads)


def composite_unframe(fr, w, P, center, out_length):
    """unframe.py:164-211 restated with torch.nn.functional.fold (differentiated by torch)."""
    N, L = fr.shape[-2], fr.shape[-1]
    span = (N - 1) * P + L
    lead = fr.shape[:-2]
    x = (fr * w).reshape(-1, N, L).transpose(-2, -1)
    num = TF.fold(x, (1, span), (1, L), stride=(1, P)).reshape(*lead, span)
    den = TF.fold((w * w).reshape(1, L, 1).expand(1, L, N), (1, span), (1, L), stride=(1, P)).reshape(span)
    s = L // 2 if center else 0
    if out_length is None:
        out_length = N * P if center else span
    return (num / (den + 1e-16))[..., s:s + out_length]


@pytest.mark.parametrize("prec", ["f64", "f32"])
def test_inverse_path_gradients(prec):
    import diffsptk_b200.functional as F
    d = dev()
    g = torch.Generator().manual_seed(31)
    cdt = torch.complex128 if prec == "f64" else torch.complex64
    tol = dict(rtol=1e-8, atol=1e-10) if prec == "f64" else dict(rtol=2e-3, atol=2e-3)

    def close(a, b, what):
        scale = max(float(b.abs().max()), 1e-30)
        a, b = torch.view_as_real(a.to(torch.complex128)) if a.is_complex() else a.double(), \
            torch.view_as_real(b) if b.is_complex() else b
        assert torch.allclose(a / scale, b / scale, **tol), (what, float((a - b).abs().max()), scale)

    def cplx(*shape):
        return torch.complex(torch.randn(*shape, generator=g, dtype=torch.float64),
                             torch.randn(*shape, generator=g, dtype=torch.float64)).to(d)
    # ifftr: every out_length, non power of two as well
    for K, ol in ((9, None), (9, 5), (13, 24), (257, 400)):
        Y64 = cplx(3, K)
        Y = Y64.to(cdt).requires_grad_(True)
        Yr = Y64.clone().requires_grad_(True)
        n = 2 * (K - 1)
        (gy,), ws = _vjp(lambda t: F.ifftr(t, ol), (Y,), g)
        (ref,) = _ref_vjp(lambda t: torch.fft.irfft(t, n=n)[..., :ol], (Yr,), ws)
        close(gy, ref, f"ifftr K={K} out_length={ol}")
    # unframe: windows with and without zeros at the ends, short requested lengths, no centring
    for shape, P, kw, ol in (((2, 9, 12), 5, dict(), None), ((2, 9, 12), 5, dict(), 20),
                             ((3, 30, 400), 80, dict(window="hamming", norm="power"), 2000),
                             ((2, 11, 33), 7, dict(center=False, window="hamming", norm="none"), None)):
        fr64 = torch.randn(*shape, generator=g, dtype=torch.float64).to(d)
        fr = fr64.to(torch.float64 if prec == "f64" else torch.float32).requires_grad_(True)
        frr = fr64.clone().requires_grad_(True)
        w = F.window(torch.ones(shape[-1], dtype=torch.float64, device=d), None,
                     window=kw.get("window", "rectangular"), norm=kw.get("norm", "none"))
        (gf,), ws = _vjp(lambda t: F.unframe(t, ol, frame_period=P, **kw), (fr,), g)
        (ref,) = _ref_vjp(lambda t: composite_unframe(t, w, P, kw.get("center", True), ol), (frr,), ws)
        close(gf, ref, f"unframe {shape} P={P} {kw} out_length={ol}")
    # istft: the fused kernel forward, the fused complex STFT as its adjoint
    for shape, kw, ol in (((2, 13, 257), dict(), 1000), ((2, 13, 257), dict(), None), ((1, 40, 257), dict(), 700),
                          ((3, 7, 9), dict(frame_length=12, frame_period=5, fft_length=16, center=False,
                                           window="hamming"), None),
                          ((2, 9, 25), dict(frame_length=30, frame_period=7, fft_length=48, window="hanning"), 50)):
        Y64 = cplx(*shape)
        Y = Y64.to(cdt).requires_grad_(True)
        Yr = Y64.clone().requires_grad_(True)
        fl, fp, n = kw.get("frame_length", 400), kw.get("frame_period", 80), kw.get("fft_length", 512)
        w = F.window(torch.ones(fl, dtype=torch.float64, device=d), None, window=kw.get("window", "blackman"),
                     norm=kw.get("norm", "power"))
        (gy,), ws = _vjp(lambda t: F.istft(t, out_length=ol, **kw), (Y,), g)
        (ref,) = _ref_vjp(lambda t: composite_unframe(torch.fft.irfft(t, n=n)[..., :fl], w, fp,
                                                      kw.get("center", True), ol), (Yr,), ws)
        close(gy, ref, f"istft {shape} {kw} out_length={ol}")
    # analysis -> synthesis round trip is differentiable end to end: d sum(istft(stft(x))) / dx == 1
    x = torch.randn(2, 4000, device=d, dtype=torch.float64 if prec == "f64" else torch.float32, requires_grad=True)
    F.istft(F.stft(x, out_format="complex"), out_length=4000).sum().backward()
    assert float((x.grad - 1).abs().max()) < (1e-9 if prec == "f64" else 2e-4)


def test_fftcep_gradients():
    import diffsptk_b200.functional as F
    d = dev()
    g = torch.Generator().manual_seed(37)
    x = (torch.rand(3, 4, 33, generator=g, dtype=torch.float64) + 0.1).to(d).requires_grad_(True)

    def ref(t, M, accel, n_iter):   # fftcep.py:116-136
        N, Hn = M + 1, t.size(-1)
        e = torch.fft.irfft(torch.log(t))
        v = e[..., :N]
        e = TF.pad(e[..., N:Hn], (N, 0))
        for _ in range(n_iter):
            e = torch.fft.hfft(e).clamp(min=0)
            e = torch.fft.ihfft(e).real
            tt = e[..., :N] * (1 + accel)
            v = v + tt
            e = e - TF.pad(tt, (0, Hn - N))
        scale = torch.ones(N, dtype=t.dtype, device=t.device)
        scale[0] = 0.5
        if Hn == N:
            scale[N - 1] = 0.5
        return v * scale
    for M, accel, n_iter in ((8, 0.0, 0), (8, 0.5, 3), (32, 0.0, 2)):
        (gx,), ws = _vjp(lambda t: F.fftcep(t, M, accel, n_iter), (x,), g)
        (rx,) = _ref_vjp(lambda t: ref(t, M, accel, n_iter), (x,), ws)
        assert torch.allclose(gx, rx, rtol=1e-8, atol=1e-10), (M, accel, n_iter)


def test_delta_gradients():
    import diffsptk_b200 as B
    d = dev()
    g = torch.Generator().manual_seed(41)
    for shape, seed, so in (((2, 9, 3), [[-0.5, 0, 0.5], [1, -2, 1]], True), ((7, 4), [3, 2], False),
                            ((3, 1, 5), [2, 2], True), ((2, 2, 6), [[1, -1, 2, 0.5, 3]], True),
                            ((2, 20, 3), [5], True), ((1, 300, 40), [4, 3], True),     # widths 11 and 9
                            ((2, 6, 700), [2], True)):                                # wide features: fallback kernel
        x = torch.randn(*shape, generator=g, dtype=torch.float64).to(d).requires_grad_(True)
        mod = B.Delta(seed, so, dtype=torch.float64).to(d)

        def ref(t):   # delta.py:172-194
            t3 = t if t.dim() == 3 else t.unsqueeze(0)
            Bn, Tn, _ = t3.shape
            W = mod.window.size(-1)
            p = (W - 1) // 2
            y = TF.conv2d(TF.pad(t3.unsqueeze(1), (0, 0, p, p), mode="replicate"), mod.window.view(-1, 1, W, 1))
            y = y.permute(0, 2, 1, 3).reshape(Bn, Tn, -1)
            return y if t.dim() == 3 else y.squeeze(0)
        (gx,), ws = _vjp(mod, (x,), g)
        (rx,) = _ref_vjp(ref, (x,), ws)
        assert torch.allclose(mod(x), ref(x), rtol=1e-12, atol=1e-12), (shape, seed)
        assert torch.allclose(gx, rx, rtol=1e-10, atol=1e-12), (shape, seed)


def test_converter_gradients():
    """lpc2par / par2lpc / gnorm / ignorm / norm0 / mc2b / b2mc: gradients against torch.autograd.gradcheck-style
    finite differences of the kernels themselves (float64) -- the backward recomputes a torch composite, so a
    mismatch between kernel and composite shows up here."""
    import diffsptk_b200.functional as F
    d = dev()
    g = torch.Generator().manual_seed(5)
    k = torch.empty(6, 9, dtype=torch.float64).uniform_(-0.8, 0.8, generator=g)
    k[:, 0] = torch.empty(6, dtype=torch.float64).uniform_(0.5, 2.0, generator=g)
    k = k.to(d)
    with torch.no_grad():
        a = F.par2lpc(k)
    cep = (0.3 * torch.randn(6, 9, generator=g, dtype=torch.float64)).to(d)
    cep[:, 0] = cep[:, 0].abs() + 0.2
    fns = [("lpc2par", lambda t: F.lpc2par(t), a), ("lpc2par_g", lambda t: F.lpc2par(t, gamma=0.5), a),
           ("par2lpc", lambda t: F.par2lpc(t), k), ("par2lpc_c", lambda t: F.par2lpc(t, c=2), k),
           ("gnorm0", lambda t: F.gnorm(t), cep), ("gnorm", lambda t: F.gnorm(t, gamma=-0.5), cep),
           ("ignorm0", lambda t: F.ignorm(t), cep), ("ignorm", lambda t: F.ignorm(t, gamma=0.5), cep),
           ("norm0", lambda t: F.norm0(t), a), ("mc2b", lambda t: F.mc2b(t, 0.42), cep),
           ("b2mc", lambda t: F.b2mc(t, 0.42), cep)]
    for name, fn, x0 in fns:
        x = x0.clone().requires_grad_(True)
        y = fn(x)
        w = torch.randn(y.shape, generator=g, dtype=torch.float64).to(d)
        (gx,) = torch.autograd.grad((y * w).sum(), x)
        num = torch.zeros_like(x0)
        h = 1e-6
        with torch.no_grad():
            for j in range(x0.shape[-1]):
                e = torch.zeros_like(x0)
                e[:, j] = h
                num[:, j] = ((fn(x0 + e) - fn(x0 - e)) * w).sum(-1) / (2 * h)
        assert torch.allclose(gx, num, rtol=1e-5, atol=1e-6), name


def test_thsolve_matches_a_dense_solve_and_its_gradients():
    """dsb200_thsolve against torch.linalg.solve of the explicit Toeplitz + Hankel matrix (float64), forward and
    the three gradients (the backward reuses the kernel: the matrix is symmetric)."""
    from diffsptk_b200 import ops
    d = dev()
    g = torch.Generator().manual_seed(11)
    for M in (1, 2, 8, 24, 40):
        i = torch.arange(M)
        t0 = torch.randn(5, M, generator=g, dtype=torch.float64) * 0.1
        t0[:, 0] = 3.0 + torch.rand(5, generator=g, dtype=torch.float64)          # diagonally dominant: well posed
        h0 = torch.randn(5, 2 * M - 1, generator=g, dtype=torch.float64) * 0.1
        r0 = torch.randn(5, M, generator=g, dtype=torch.float64)
        w = torch.randn(5, M, generator=g, dtype=torch.float64).to(d)
        t, h, r = (v.to(d).requires_grad_(True) for v in (t0, h0, r0))
        x = ops.thsolve(t, h, r)
        gt, gh, gr = torch.autograd.grad((x * w).sum(), (t, h, r))
        t2, h2, r2 = (v.to(d).requires_grad_(True) for v in (t0, h0, r0))
        A = t2[..., (i[:, None] - i[None, :]).abs().to(d)] + h2[..., (i[:, None] + i[None, :]).to(d)]
        x2 = torch.linalg.solve(A, r2.unsqueeze(-1)).squeeze(-1)
        rt, rh, rr = torch.autograd.grad((x2 * w).sum(), (t2, h2, r2))
        assert torch.allclose(x, x2, rtol=1e-9, atol=1e-11), M
        for a, b, what in ((gt, rt, "t"), (gh, rh, "h"), (gr, rr, "r")):
            assert torch.allclose(a, b, rtol=1e-8, atol=1e-10), (M, what)


def test_mcep_gradients():
    """mcep backward = recompute of the Newton iteration on differentiable kernels (rowmat, thsolve) + autograd:
    checked against central differences of the fused forward kernel itself (float64)."""
    import diffsptk_b200.functional as F
    d = dev()
    g = torch.Generator().manual_seed(23)
    for K, M, alpha, n_iter in ((17, 4, 0.1, 2), (17, 4, 0.3, 0), (257, 24, 0.42, 3)):
        base = torch.randn(3, 2 * (K - 1), generator=g, dtype=torch.float64)
        x0 = (torch.fft.rfft(base).abs().square() + 0.5).to(d)
        w = torch.randn(3, M + 1, generator=g, dtype=torch.float64).to(d)
        x = x0.clone().requires_grad_(True)
        (gx,) = torch.autograd.grad((F.mcep(x, M, alpha, n_iter) * w).sum(), x)
        cols = torch.randperm(K, generator=g)[:12].tolist()
        with torch.no_grad():
            for j in cols:
                h = 1e-6 * float(x0[:, j].abs().max())
                e = torch.zeros_like(x0)
                e[:, j] = h
                num = ((F.mcep(x0 + e, M, alpha, n_iter) - F.mcep(x0 - e, M, alpha, n_iter)) * w).sum(-1) / (2 * h)
                assert torch.allclose(gx[:, j], num, rtol=2e-4, atol=1e-7), (K, M, n_iter, j)
    # float32 and the module path run, and mgcep with gamma = 0 (which is this kernel) is differentiable too
    import diffsptk_b200 as B
    xs = (torch.rand(5, 257, generator=g) + 0.1).to(d).requires_grad_(True)
    B.MelCepstralAnalysis(fft_length=512, cep_order=24, alpha=0.42, n_iter=2).to(d)(xs).sum().backward()
    assert torch.isfinite(xs.grad).all() and float(xs.grad.abs().max()) > 0
    xs.grad = None
    B.MelGeneralizedCepstralAnalysis(fft_length=512, cep_order=12, alpha=0.42, gamma=-0.5, n_iter=2).to(d)(xs).sum().backward()
    assert torch.isfinite(xs.grad).all() and float(xs.grad.abs().max()) > 0


def test_lpc2lsp_gradients():
    """Implicit-differentiation backward of lpc2lsp against central differences of the kernel (float64)."""
    import diffsptk_b200.functional as F
    d = dev()
    g = torch.Generator().manual_seed(31)
    for M, kw in ((1, {}), (7, dict(log_gain=True, sample_rate=8000, out_format="khz")), (8, {}), (16, dict(out_format="cycle"))):
        k = torch.empty(4, M + 1, dtype=torch.float64).uniform_(-0.7, 0.7, generator=g)
        k[:, 0] = torch.empty(4, dtype=torch.float64).uniform_(0.5, 2.0, generator=g)
        with torch.no_grad():
            a0 = F.par2lpc(k.to(d))        # a stable filter
        w = torch.randn(4, M + 1, generator=g, dtype=torch.float64).to(d)
        a = a0.clone().requires_grad_(True)
        (ga,) = torch.autograd.grad((F.lpc2lsp(a, **kw) * w).sum(), a)
        num = torch.zeros_like(a0)
        h = 1e-6
        with torch.no_grad():
            for j in range(M + 1):
                e = torch.zeros_like(a0)
                e[:, j] = h
                num[:, j] = ((F.lpc2lsp(a0 + e, **kw) - F.lpc2lsp(a0 - e, **kw)) * w).sum(-1) / (2 * h)
        assert torch.allclose(ga, num, rtol=1e-5, atol=1e-6), (M, (ga - num).abs().max())


def test_composite_consumer_gradients():
    """mgc2mgc / mgc2sp / plp / mgcep chain kernels that each carry their own backward: the chain rule through the
    whole composite against central differences of the forward (float64)."""
    import diffsptk_b200 as B
    import diffsptk_b200.functional as F
    d = dev()
    g = torch.Generator().manual_seed(77)
    c = (0.2 * torch.randn(3, 9, generator=g, dtype=torch.float64))
    c[:, 0] = 0.6
    P = (torch.fft.rfft(torch.randn(3, 64, generator=g, dtype=torch.float64)).abs().square() + 0.1)
    mg = B.MelGeneralizedCepstralAnalysis(fft_length=64, cep_order=6, alpha=0.2, gamma=-0.5, n_iter=2,
                                          dtype=torch.float64).to(d)
    cases = [("mgc2mgc", lambda t: F.mgc2mgc(t, 10, in_alpha=0.1, out_alpha=0.3, in_gamma=-0.5, out_gamma=-0.25, n_fft=128), c),
             ("mgc2mgc_norm", lambda t: F.mgc2mgc(t, 8, in_gamma=0.0, out_gamma=-1.0, out_norm=True, out_mul=True, n_fft=64), c),
             ("mgc2sp", lambda t: F.mgc2sp(t, 32, alpha=0.3, gamma=-0.5, n_fft=128, out_format="log-magnitude"), c),
             ("plp", lambda t: F.plp(t, 5, 10, 8000, lifter=20, floor=1e-3, out_format="ycE"), P),
             ("mgcep", mg, P)]
    for name, fn, x0 in cases:
        x0 = x0.to(d)
        x = x0.clone().requires_grad_(True)
        y = fn(x)
        w = torch.randn(y.shape, generator=g, dtype=torch.float64).to(d)
        (gx,) = torch.autograd.grad((y * w).sum(), x)
        num = torch.zeros_like(x0)
        with torch.no_grad():
            for j in range(x0.shape[-1]):
                h = 1e-6 * max(1.0, float(x0[:, j].abs().max()))
                e = torch.zeros_like(x0)
                e[:, j] = h
                num[:, j] = ((fn(x0 + e) - fn(x0 - e)) * w).sum(-1) / (2 * h)
        scale = float(num.abs().max())
        err = float((gx - num).abs().max())
        assert err <= 2e-5 * scale + 1e-7, (name, err, scale)


@pytest.mark.parametrize("fmt,rf", [("power", None), ("db", -40.0), ("magnitude", None)])
def test_polezero_spectrum_gradients_reach_numerator_and_denominator(fmt, rf):
    """Spectrum(b, a) = K |B| / |A| under autograd (spec.py:162-178): ADVICE round 1 -- the fused kernel differentiates
    the numerator only; the module now routes such calls through the differentiable numerator kernel so that gradients
    reach both polynomials.  Checker: a float64 torch composite (rfft + abs), used only here."""
    import diffsptk_b200 as B
    g = torch.Generator().manual_seed(9)
    b0 = torch.randn(3, 5, generator=g, dtype=torch.float64)
    a0 = torch.randn(3, 7, generator=g, dtype=torch.float64) * 0.3
    a0[:, 0] = a0[:, 0].abs() + 1.0
    a0[:, 1:] *= 0.2
    gy = torch.randn(3, 17, generator=g, dtype=torch.float64)
    mod = B.Spectrum(32, eps=1e-6, relative_floor=rf, out_format=fmt)

    def ref(b, a):
        K, a1 = a[..., :1], TF.pad(a[..., 1:], (1, 0), value=1.0)
        X = K * (torch.fft.rfft(b, n=32).abs() / torch.fft.rfft(a1, n=32).abs())
        s = X.square() + 1e-6
        if rf is not None:
            s = torch.maximum(s, s.amax(-1, keepdim=True) * 10 ** (rf / 10))
        return {"db": lambda v: 10 * torch.log10(v), "magnitude": torch.sqrt, "power": lambda v: v}[fmt](s)

    br, ar = b0.clone().requires_grad_(True), a0.clone().requires_grad_(True)
    (ref(br, ar) * gy).sum().backward()
    for which in ("both", "b_only", "a_only"):
        b = b0.to(dev()).requires_grad_(which != "a_only")
        a = a0.to(dev()).requires_grad_(which != "b_only")
        y = mod(b, a)
        assert torch.allclose(y.detach().cpu(), ref(b0, a0), rtol=1e-9, atol=1e-10)
        (y * gy.to(dev())).sum().backward()
        if b.requires_grad:
            assert torch.allclose(b.grad.cpu(), br.grad, rtol=1e-7, atol=1e-9)
        if a.requires_grad:
            assert torch.allclose(a.grad.cpu(), ar.grad, rtol=1e-7, atol=1e-9)
    # without gradients the fused pole-zero kernel runs, and agrees
    with torch.no_grad():
        assert torch.allclose(mod(b0.to(dev()), a0.to(dev())).cpu(), ref(b0, a0), rtol=1e-9, atol=1e-10)


@pytest.mark.parametrize("center", [True, False])
def test_learnable_synthesis_window_gradients(center):
    """Unframe / ISTFT with learnable=True (unframe.py:150-152, istft.py:186-193 of the reference): the window is a
    (1, L, 1) parameter like the reference's and receives its gradient (it used to raise).  Checker:
    the reference's fold formula restated in float64 torch."""
    import diffsptk_b200 as B
    L, P, Nf = 12, 5, 9
    g = torch.Generator().manual_seed(13)
    y0 = torch.randn(2, Nf, L, generator=g, dtype=torch.float64)

    def ref(y, w, out_length):
        span = (Nf - 1) * P + L
        x = TF.fold((y * w).transpose(-2, -1), (1, span), (1, L), stride=(1, P))[..., 0, 0, :]
        d = TF.fold((w * w).reshape(1, L, 1).expand(1, L, Nf), (1, span), (1, L), stride=(1, P))[..., 0, 0, :]
        s = L // 2 if center else 0
        return (x / (d + 1e-16))[..., s:s + out_length]

    un = B.Unframe(L, P, center=center, window="hanning", norm="none", learnable=True, dtype=torch.float64).to(dev())
    assert tuple(un.window.shape) == (1, L, 1) and isinstance(un.window, torch.nn.Parameter)
    T = Nf * P if center else (Nf - 1) * P + 3
    gy = torch.randn(2, T, generator=g, dtype=torch.float64)
    w_ref = un.window.detach().cpu().reshape(-1).clone().requires_grad_(True)
    y_ref = y0.clone().requires_grad_(True)
    (ref(y_ref, w_ref, T) * gy).sum().backward()
    y = y0.to(dev()).requires_grad_(True)
    out = un(y, T)
    assert torch.allclose(out.detach().cpu(), ref(y0, w_ref.detach(), T), rtol=1e-10, atol=1e-12)
    (out * gy.to(dev())).sum().backward()
    assert torch.allclose(y.grad.cpu(), y_ref.grad, rtol=1e-9, atol=1e-11)
    assert torch.allclose(un.window.grad.cpu().reshape(-1), w_ref.grad, rtol=1e-9, atol=1e-11)

    # ISTFT: frames = irfft(Y); the synthesis window of the unframe sub-layer is the learnable one
    ist = B.ISTFT(L, P, 16, center=center, window="hanning", norm="none", learnable=["window"],
                  dtype=torch.float64).to(dev())
    Y0 = torch.randn(2, Nf, 9, generator=g, dtype=torch.complex128)
    w_ref = ist.unframe.window.detach().cpu().reshape(-1).clone().requires_grad_(True)
    Y_ref = Y0.clone().requires_grad_(True)
    (ref(torch.fft.irfft(Y_ref, n=16)[..., :L], w_ref, T) * gy).sum().backward()
    Y = Y0.to(dev()).requires_grad_(True)
    (ist(Y, T) * gy.to(dev())).sum().backward()
    assert torch.allclose(ist.unframe.window.grad.cpu().reshape(-1), w_ref.grad, rtol=1e-9, atol=1e-11)
    # torch's irfft ignores the imaginary parts of DC / Nyquist and so does the kernel: compare the rest
    assert torch.allclose(torch.view_as_real(Y.grad.cpu()), torch.view_as_real(Y_ref.grad), rtol=1e-9, atol=1e-11)


def test_learnable_dft_basis_matches_the_fft_and_trains():
    """fftr / ifftr / Spectrum / STFT / ISTFT with a trainable DFT basis (fftr.py:123-131,146-150, ifftr.py:117-124 of
    the reference; round 1 raised NotImplementedError).  At initialisation the basis IS the DFT, so the outputs must
    agree with the fused kernels; gradients must reach the basis."""
    import diffsptk_b200 as B
    d = dev()
    g = torch.Generator().manual_seed(17)
    x = torch.randn(3, 700, generator=g, dtype=torch.float64).to(d)
    f_fix = B.RealValuedFastFourierTransform(32, dtype=torch.float64).to(d)
    f_lrn = B.RealValuedFastFourierTransform(32, learnable=True, dtype=torch.float64).to(d)
    assert isinstance(f_lrn.W, torch.nn.Parameter) and tuple(f_lrn.W.shape) == (32, 34)
    fr = x[:, :32]
    assert torch.allclose(torch.view_as_real(f_lrn(fr)), torch.view_as_real(f_fix(fr)), rtol=1e-9, atol=1e-11)
    i_fix = B.RealValuedInverseFastFourierTransform(32, 20, dtype=torch.float64).to(d)
    i_lrn = B.RealValuedInverseFastFourierTransform(32, 20, learnable=True, dtype=torch.float64).to(d)
    Y = f_fix(fr)
    assert torch.allclose(i_lrn(Y), i_fix(Y), rtol=1e-9, atol=1e-11)
    torch.set_default_dtype(torch.float64)
    try:
        s_fix = B.Spectrum(32, eps=1e-6, out_format="db").to(d)
        s_lrn = B.Spectrum(32, eps=1e-6, out_format="db", learnable=True).to(d)
    finally:
        torch.set_default_dtype(torch.float32)
    b, a = x[:, :5], torch.cat([x[:, 5:6].abs() + 1, 0.1 * x[:, 6:9]], -1)
    assert torch.allclose(s_lrn(b, a), s_fix(b, a), rtol=1e-8, atol=1e-9)
    st_fix = B.STFT(40, 10, 64, dtype=torch.float64).to(d)
    st_lrn = B.STFT(40, 10, 64, learnable=["basis"], dtype=torch.float64).to(d)
    # like the reference's, Spectrum takes no dtype: the basis inside STFT's spectrum layer is created in the default
    # dtype (float32) even for a float64 STFT (stft.py:225-235 -> spec.py:134-142), hence float32 accuracy here
    assert st_lrn.spec.fftr.W.dtype == torch.float32
    assert torch.allclose(st_lrn(x), st_fix(x), rtol=1e-4, atol=1e-5)
    names = [n for n, _ in st_lrn.named_parameters()]
    assert names == ["spec.fftr.W"], names
    st_lrn(x).sum().backward()
    W = st_lrn.spec.fftr.W
    assert W.grad is not None and bool(torch.isfinite(W.grad).all()) and float(W.grad.abs().max()) > 0
    is_lrn = B.ISTFT(40, 10, 64, learnable=True, dtype=torch.float64).to(d)
    is_fix = B.ISTFT(40, 10, 64, dtype=torch.float64).to(d)
    Yc = B.STFT(40, 10, 64, out_format="complex", dtype=torch.float64).to(d)(x)
    out = is_lrn(Yc, out_length=700)
    assert torch.allclose(out, is_fix(Yc, out_length=700), rtol=1e-8, atol=1e-10)
    out.sum().backward()
    assert sorted(n for n, p in is_lrn.named_parameters() if p.grad is not None) == ["ifftr.W", "unframe.window"]
