"""Host logic of the composite modules (mgc2mgc / mgc2sp / plp) on the CPU.

The composites chain this package's kernels with a few elementwise torch ops.  Here every kernel entry in
``diffsptk_b200.ops`` is replaced by a plain torch stand-in of the same contract (TEST ONLY -- the product has no
CPU path), so the step sequences, table construction, flag handling and output packing are checked against the
reference's golden vectors without a GPU; the GPU parity tests then run the same cases on the real kernels.
"""

import numpy as np
import pytest
import torch

import helpers as H


def _rfft(x, fft_length, out_format):
    X = torch.fft.rfft(x[..., :fft_length], n=fft_length)
    return torch.view_as_real(X).contiguous() if out_format == 0 else X.real


def _ifftr(y, out_length):
    return torch.fft.irfft(y, n=2 * (y.shape[-1] - 1))[..., :out_length]


def _fbank(x, Hm, cb, ce, floor, gamma, use_power, want_energy):
    a = x if use_power else torch.sqrt(x)
    y = torch.clip(a @ Hm, min=floor)
    y = torch.log(y) if gamma == 0 else (y ** gamma - 1) / gamma
    E = torch.log((2 * x[..., 1:-1].sum(-1) + x[..., 0] + x[..., -1]) / (2 * (x.shape[-1] - 1))).unsqueeze(-1)
    return y, (E if want_energy else x.new_empty(0))


def _levdur(r, eps):
    M = r.shape[-1] - 1
    idx = (torch.arange(M)[:, None] - torch.arange(M)[None, :]).abs()
    R = r[..., :-1][..., idx] + eps * torch.eye(M, dtype=r.dtype)
    a = torch.linalg.solve(R, -r[..., 1:].unsqueeze(-1)).squeeze(-1)
    K = torch.sqrt((r[..., 1:] * a).sum(-1, keepdim=True) + r[..., :1])
    return torch.cat((K, a), dim=-1)


def _thsolve(t, h, r):
    M = t.shape[-1]
    i = torch.arange(M)
    A = t[..., (i[:, None] - i[None, :]).abs()] + h[..., i[:, None] + i[None, :]]
    return torch.linalg.solve(A, r.unsqueeze(-1)).squeeze(-1)


@pytest.fixture()
def host_ops(monkeypatch):
    from diffsptk_b200 import ops
    monkeypatch.setattr(ops, "rfft", _rfft)
    monkeypatch.setattr(ops, "ifftr", _ifftr)
    monkeypatch.setattr(ops, "rowmat", lambda x, W: x @ W.to(x.dtype))
    monkeypatch.setattr(ops, "rowconv", lambda x, op, g: ops.rowconv_composite(x, op, g))
    monkeypatch.setattr(ops, "fbank", _fbank)
    monkeypatch.setattr(ops, "levdur", _levdur)
    monkeypatch.setattr(ops, "thsolve", _thsolve)
    monkeypatch.setattr(ops, "gc2gc", ops.gc2gc_composite)   # the kernel's contract, on the stand-in transforms
    return ops


CASES = H.case_names(["mgc2mgc", "mgc2sp", "plp"])


@pytest.mark.parametrize("prec", ["f64", "f32"])
@pytest.mark.parametrize("name", CASES)
def test_functional_host_logic_matches_reference(host_ops, name, prec):
    import diffsptk_b200.functional as F
    op, params, ins, outs = H.load_case(name, prec)
    got = getattr(F, op)(*[torch.from_numpy(np.ascontiguousarray(a)) for a in ins], **params)
    H.assert_close(got.numpy(), outs[0], prec, what=f"{name}[{prec}]", scale_atol=True)


@pytest.mark.parametrize("name", CASES[::3])
def test_module_host_logic_matches_reference(host_ops, name):
    import diffsptk_b200 as B
    op, params, ins, outs = H.load_case(name, "f64")
    n = ins[0].shape[-1]
    p = dict(params)
    if op == "mgc2mgc":
        mod = B.MelGeneralizedCepstrumToMelGeneralizedCepstrum(n - 1, **p, dtype=torch.float64)
    elif op == "mgc2sp":
        mod = B.MelGeneralizedCepstrumToSpectrum(n - 1, p.pop("fft_length"), **p, dtype=torch.float64)
    else:
        mod = B.PLP(fft_length=2 * n - 2, **p, dtype=torch.float64)
    got = mod(torch.from_numpy(np.ascontiguousarray(ins[0])))
    H.assert_close(got.numpy(), outs[0], "f64", what=name, scale_atol=True)


@pytest.mark.parametrize("prec", ["f64", "f32"])
@pytest.mark.parametrize("name", [n for n in H.case_names(["mgcep"]) if H.load_case(n, "f64")[1].get("gamma") != 0
                                  or H.load_case(n, "f64")[1].get("c")])
def test_mgcep_host_logic_matches_reference(host_ops, name, prec):
    """gamma != 0 only: gamma = 0 is the fused mcep kernel, which has no stand-in."""
    import diffsptk_b200 as B
    op, params, ins, outs = H.load_case(name, prec)
    mod = B.MelGeneralizedCepstralAnalysis(**params, dtype=torch.float64 if prec == "f64" else torch.float32)
    got = mod(torch.from_numpy(np.ascontiguousarray(ins[0])))
    if prec == "f32":
        H.assert_close_conditioned(got.numpy(), outs[0], H.load_case(name, "f64")[3][0], what=f"{name}[f32]")
    else:
        H.assert_close(got.numpy(), outs[0], prec, what=f"{name}[{prec}]", scale_atol=True)


def test_learnable_dft_basis_host_logic(host_ops, monkeypatch):
    """Wiring of the trainable DFT basis (fftr.py:123-131, ifftr.py:117-124, spec.py:134-178, stft.py:198-241 of the
    reference) with the kernels replaced by torch stand-ins: at initialisation the basis is the DFT, so every module
    must reproduce torch.fft; parameter names follow the reference."""
    import torch.nn.functional as TF

    import diffsptk_b200 as B
    ops = host_ops
    monkeypatch.setattr(ops, "frame", lambda x, L, P, center, zmean, pad: TF.pad(x, (L // 2, (L - 1) // 2)).unfold(-1, L, P))
    monkeypatch.setattr(ops, "window", lambda x, w, n: TF.pad(x * w, (0, n - x.shape[-1])))

    def _unframe(y, w, T, P, center):
        L, Nf = y.shape[-1], y.shape[-2]
        span = (Nf - 1) * P + L
        num = TF.fold((y * w).transpose(-2, -1), (1, span), (1, L), stride=(1, P))[..., 0, 0, :]
        den = TF.fold((w * w).reshape(1, L, 1).expand(1, L, Nf), (1, span), (1, L), stride=(1, P))[..., 0, 0, :]
        s = L // 2 if center else 0
        return (num / (den + 1e-16))[..., s:s + T]
    monkeypatch.setattr(ops, "unframe", _unframe)
    g = torch.Generator().manual_seed(1)
    x = torch.randn(2, 300, generator=g, dtype=torch.float64)
    f = B.RealValuedFastFourierTransform(32, learnable=True, dtype=torch.float64)
    assert tuple(f.W.shape) == (32, 34) and isinstance(f.W, torch.nn.Parameter)
    assert torch.allclose(f(x[:, :20]), torch.fft.rfft(x[:, :20], n=32))
    for fmt, fn in (("real", lambda v: v.real), ("amplitude", torch.abs), ("power", lambda v: v.abs() ** 2)):
        fm = B.RealValuedFastFourierTransform(32, out_format=fmt, learnable=True, dtype=torch.float64)
        assert torch.allclose(fm(x[:, :32]), fn(torch.fft.rfft(x[:, :32])))
    Y = torch.fft.rfft(x[:, :32])
    inv = B.RealValuedInverseFastFourierTransform(32, 20, learnable=True, dtype=torch.float64)
    assert tuple(inv.W.shape) == (34, 20)
    assert torch.allclose(inv(Y), torch.fft.irfft(Y)[..., :20])
    torch.set_default_dtype(torch.float64)
    try:
        sp = B.Spectrum(32, eps=1e-6, relative_floor=-30.0, out_format="db", learnable=True)
    finally:
        torch.set_default_dtype(torch.float32)
    assert [n for n, _ in sp.named_parameters()] == ["fftr.W"]
    b, a = x[:, :5], torch.cat([x[:, 5:6].abs() + 1, 0.1 * x[:, 6:9]], -1)
    a1 = TF.pad(a[:, 1:], (1, 0), value=1.0)
    s = (a[:, :1] * torch.fft.rfft(b, n=32).abs() / torch.fft.rfft(a1, n=32).abs()) ** 2 + 1e-6
    s = torch.maximum(s, s.amax(-1, keepdim=True) * 10 ** (-3.0))
    assert torch.allclose(sp(b, a), 10 * torch.log10(s))
    st = B.STFT(40, 10, 64, learnable=True, dtype=torch.float64)
    assert sorted(n for n, _ in st.named_parameters()) == ["spec.fftr.W", "window.window"]
    fr = TF.pad(x, (20, 19)).unfold(-1, 40, 10) * st.window.window.detach()
    assert torch.allclose(st(x), torch.fft.rfft(fr, n=64).abs() ** 2 + 1e-9)
    stc = B.STFT(40, 10, 64, out_format="complex", learnable=["basis"], dtype=torch.float64)
    assert [n for n, _ in stc.named_parameters()] == ["spec.W"]
    assert torch.allclose(stc(x), torch.fft.rfft(TF.pad(x, (20, 19)).unfold(-1, 40, 10) * stc.window.window, n=64))
    ist = B.ISTFT(40, 10, 64, learnable=True, dtype=torch.float64)
    assert sorted(n for n, _ in ist.named_parameters()) == ["ifftr.W", "unframe.window"]
    assert torch.allclose(ist(stc(x), out_length=300), x, atol=1e-9)
    ist(stc(x), out_length=300).sum().backward()
    assert ist.ifftr.W.grad is not None and stc.spec.W.grad is not None
