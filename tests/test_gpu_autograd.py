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
