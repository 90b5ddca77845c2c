"""Drop-in contract: class names/aliases, constructor and functional signatures, ``_takes_input_size``
flags, buffer names and the static protocol match the reference (snapshot in golden/api_signatures.json,
taken from /root/reference by golden/make_api_snapshot.py)."""

import json
import os
import sys

import pytest
import torch

import diffsptk_b200 as B

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
from make_api_snapshot import CLASSES, FUNCTIONS, snapshot  # noqa: E402

REF = json.load(open(os.path.join(HERE, "golden", "api_signatures.json")))
MINE = json.loads(json.dumps(snapshot(B)))


@pytest.mark.parametrize("name", CLASSES)
def test_class_signature(name):
    assert MINE["classes"][name] == REF["classes"][name]


@pytest.mark.parametrize("name", FUNCTIONS)
def test_functional_signature(name):
    assert MINE["functions"][name] == REF["functions"][name]


def test_protocol_and_buffers():
    for name in CLASSES:
        cls = getattr(B, name)
        if name == "MelGeneralizedCepstralAnalysis":   # a module without functional interface in the reference too
            continue
        for m in ("_func", "_check", "_precompute", "_forward"):
            assert isinstance(cls.__dict__.get(m) or getattr(cls, m), (staticmethod, type(lambda: 0))), (name, m)
    # buffer names other reference modules / checkpoints rely on
    assert [n for n, _ in B.STFT(400, 80, 512).named_buffers()] == ["window.window"]
    assert [n for n, _ in B.LPC(400, 24).named_buffers()] == ["levdur.eye"]
    names = {n for n, _ in B.MelCepstralAnalysis(fft_length=64, cep_order=8, alpha=0.3, n_iter=1).named_buffers()}
    assert {"alpha_vector", "freqt.A", "ifreqt.A", "rfreqt.A"} <= names
    names = {n for n, _ in B.MFCC(fft_length=64, mfcc_order=4, n_channel=8, sample_rate=8000).named_buffers()}
    assert {"liftering_vector", "fbank.H", "dct.W"} <= names <= {"liftering_vector", "fbank.H", "dct.W",
                                                                "fbank.H_begin", "fbank.H_end"}
    lf = B.FBANK(fft_length=64, n_channel=8, sample_rate=8000, learnable=True)
    assert [n for n, _ in lf.named_parameters()] == ["H"] and not list(lf.named_buffers())
    assert B.STFT(400, 80, 512).state_dict() == {}  # non-persistent buffers, as in the reference
    m = B.STFT(400, 80, 512, learnable=["window"])
    assert [n for n, _ in m.named_parameters()] == ["window.window"]


def test_value_errors_match_reference_messages():
    with pytest.raises(ValueError, match="frame_length must be positive"):
        B.Frame(0, 1)
    with pytest.raises(ValueError, match="fft_length must be positive even"):
        B.RealValuedFastFourierTransform(7)
    with pytest.raises(ValueError, match="relative_floor must be negative"):
        B.Spectrum(8, relative_floor=3.0)
    with pytest.raises(ValueError, match="acr_order must be less than frame_length"):
        B.Autocorrelation(4, 4)
    with pytest.raises(ValueError, match="alpha must be in"):
        B.FrequencyTransform(3, 3, 1.0)
    with pytest.raises(ValueError, match="cep_order must be less than or equal"):
        B.MelCepstralAnalysis(fft_length=8, cep_order=5)
    with pytest.raises(ValueError, match="mfcc_order must be less than n_channel"):
        B.MFCC(fft_length=32, mfcc_order=8, n_channel=8, sample_rate=8000)
    with pytest.raises(ValueError, match="Unexpected input length"):
        B.Window(8)(torch.zeros(9))
    with pytest.raises(ValueError, match="window foo is not supported"):
        B.Window(8, window="foo")
    with pytest.raises(ValueError, match="An unsupported key"):
        B.STFT(8, 2, 8, learnable=["nope"])


def test_cpu_tensors_fail_loudly():
    """No CPU fallback: a CPU tensor must raise, never silently compute."""
    with pytest.raises((NotImplementedError, RuntimeError)):
        B.functional.stft(torch.randn(1000))
    with pytest.raises((NotImplementedError, RuntimeError)):
        B.Frame(400, 80)(torch.randn(1000))
