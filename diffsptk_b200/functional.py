"""Functional API of the hot path: same names, parameters and defaults as the 13 delegates in
``diffsptk/functional.py`` (acorr :23, dct :262, fbank :695, fftr :797, frame :859, freqt :905,
levdur :1659, lpc :1700, mcep :1956, mfcc :2103, spec :2916, stft :2963, window :3142).

Each function is ``<Module>._func(...)``; host tables are memoised (``tables.py``) instead of being
rebuilt per call as in the reference.  ``lpc_from_waveform`` / ``mfcc_from_waveform`` are the fused
pipelines of BASELINE.json (configs 3 and 5); they have no reference counterpart other than the
module cascades they replace.
"""

from __future__ import annotations

from torch import Tensor

from . import modules as nn
from . import fused as _fused


def acorr(x: Tensor, acr_order: int, out_format: str = "naive") -> Tensor:
    """Autocorrelation of framed waveforms ``(..., L) -> (..., M+1)``."""
    return nn.Autocorrelation._func(x, acr_order=acr_order, out_format=out_format)


def dct(x: Tensor, dct_type: int = 2) -> Tensor:
    """DCT ``(..., L) -> (..., L)``."""
    return nn.DiscreteCosineTransform._func(x, dct_type=dct_type)


def fbank(x: Tensor, n_channel: int, sample_rate: int, f_min: float = 0, f_max: float | None = None,
          floor: float = 1e-5, gamma: float = 0, scale: str = "htk", erb_factor: float | None = None,
          use_power: bool = False, out_format: str = "y") -> Tensor | tuple[Tensor, Tensor]:
    """Mel filter-bank analysis of a power spectrum ``(..., L/2+1) -> (..., C)`` [+ energy]."""
    return nn.MelFilterBankAnalysis._func(
        x, n_channel=n_channel, sample_rate=sample_rate, f_min=f_min, f_max=f_max, floor=floor, gamma=gamma,
        scale=scale, erb_factor=erb_factor, use_power=use_power, out_format=out_format)


def fftr(x: Tensor, fft_length: int | None = None, out_format: str = "complex") -> Tensor:
    """Real FFT ``(..., N) -> (..., L/2+1)``."""
    return nn.RealValuedFastFourierTransform._func(x, fft_length=fft_length, out_format=out_format)


def frame(x: Tensor, frame_length: int = 400, frame_period: int = 80, center: bool = True,
          zmean: bool = False, mode: str = "constant") -> Tensor:
    """Framing ``(..., T) -> (..., T/P, L)``."""
    return nn.Frame._func(x, frame_length=frame_length, frame_period=frame_period, center=center, zmean=zmean,
                          mode=mode)


def freqt(c: Tensor, out_order: int, alpha: float = 0) -> Tensor:
    """Frequency transform ``(..., M1+1) -> (..., M2+1)``."""
    return nn.FrequencyTransform._func(c, out_order=out_order, alpha=alpha)


def levdur(r: Tensor, eps: float | None = None) -> Tensor:
    """Levinson-Durbin ``(..., M+1) -> (..., M+1)``."""
    return nn.LevinsonDurbin._func(r, eps=eps)


def lpc(x: Tensor, lpc_order: int, eps: float | None = None) -> Tensor:
    """LPC analysis of framed waveforms ``(..., L) -> (..., M+1)``."""
    return nn.LinearPredictiveCodingAnalysis._func(x, lpc_order=lpc_order, eps=eps)


def mcep(x: Tensor, cep_order: int, alpha: float = 0, n_iter: int = 0) -> Tensor:
    """Mel-cepstral analysis of a power spectrum ``(..., L/2+1) -> (..., M+1)``."""
    return nn.MelCepstralAnalysis._func(x, cep_order=cep_order, alpha=alpha, n_iter=n_iter)


def mfcc(x: Tensor, mfcc_order: int, n_channel: int, sample_rate: int, lifter: int = 1, f_min: float = 0,
         f_max: float | None = None, floor: float = 1e-5, gamma: float = 0, scale: str = "htk",
         erb_factor: float | None = None, out_format: str = "y") -> Tensor:
    """MFCC of a power spectrum ``(..., L/2+1) -> (..., M)`` [+ c0] [+ energy]."""
    return nn.MelFrequencyCepstralCoefficientsAnalysis._func(
        x, mfcc_order=mfcc_order, n_channel=n_channel, sample_rate=sample_rate, lifter=lifter, f_min=f_min,
        f_max=f_max, floor=floor, gamma=gamma, scale=scale, erb_factor=erb_factor, out_format=out_format)


def spec(b: Tensor | None = None, a: Tensor | None = None, *, fft_length: int = 512, eps: float = 0,
         relative_floor: float | None = None, out_format: str = "power") -> Tensor:
    """Spectrum of ``b / a`` ``-> (..., L/2+1)``."""
    return nn.Spectrum._func(b, a, fft_length=fft_length, eps=eps, relative_floor=relative_floor,
                             out_format=out_format)


def stft(x: Tensor, *, frame_length: int = 400, frame_period: int = 80, fft_length: int = 512,
         center: bool = True, zmean: bool = False, mode: str = "constant", window: str = "blackman",
         norm: str = "power", symmetric: bool = True, eps: float = 1e-9, relative_floor: float | None = None,
         out_format: str = "power") -> Tensor:
    """Short-time Fourier transform ``(..., T) -> (..., T/P, N/2+1)`` in one fused kernel."""
    return nn.ShortTimeFourierTransform._func(
        x, frame_length=frame_length, frame_period=frame_period, fft_length=fft_length, center=center,
        zmean=zmean, mode=mode, window=window, norm=norm, symmetric=symmetric, eps=eps,
        relative_floor=relative_floor, out_format=out_format)


def window(x: Tensor, out_length: int | None = None, *, window: str = "blackman", norm: str = "power",
           symmetric: bool = True) -> Tensor:
    """Windowing ``(..., L1) -> (..., L2)``."""
    return nn.Window._func(x, out_length=out_length, window=window, norm=norm, symmetric=symmetric)


def delta(x: Tensor, seed=[[-0.5, 0, 0.5]], static_out: bool = True) -> Tensor:  # noqa: B006 (reference default)
    """Delta features ``(B, T, D) or (T, D) -> (..., T, D x H)``."""
    return nn.Delta._func(x, seed, static_out=static_out)


def b2mc(b: Tensor, alpha: float = 0) -> Tensor:
    """MLSA filter coefficients to mel-cepstrum ``(..., M+1) -> (..., M+1)``."""
    return nn.MLSADigitalFilterCoefficientsToMelCepstrum._func(b, alpha=alpha)


def mc2b(mc: Tensor, alpha: float = 0) -> Tensor:
    """Mel-cepstrum to MLSA filter coefficients ``(..., M+1) -> (..., M+1)``."""
    return nn.MelCepstrumToMLSADigitalFilterCoefficients._func(mc, alpha=alpha)


def gnorm(x: Tensor, gamma: float = 0, c: int | None = None) -> Tensor:
    """Gain normalisation of a generalized cepstrum ``(..., M+1) -> (..., M+1)``."""
    return nn.GeneralizedCepstrumGainNormalization._func(x, gamma=gamma, c=c)


def ignorm(y: Tensor, gamma: float = 0, c: int | None = None) -> Tensor:
    """Inverse gain normalisation ``(..., M+1) -> (..., M+1)``."""
    return nn.GeneralizedCepstrumInverseGainNormalization._func(y, gamma=gamma, c=c)


def lpc2lsp(a: Tensor, log_gain: bool = False, sample_rate: int | None = None, out_format: str = "radian") -> Tensor:
    """LPC to line spectral pairs ``(..., M+1) -> (..., M+1)``."""
    return nn.LinearPredictiveCoefficientsToLineSpectralPairs._func(a, log_gain=log_gain, sample_rate=sample_rate,
                                                                    out_format=out_format)


def lpc2par(a: Tensor, gamma: float = 1, c: int | None = None) -> Tensor:
    """LPC to PARCOR coefficients ``(..., M+1) -> (..., M+1)``."""
    return nn.LinearPredictiveCoefficientsToParcorCoefficients._func(a, gamma=gamma, c=c)


def par2lpc(k: Tensor, gamma: float = 1, c: int | None = None) -> Tensor:
    """PARCOR to LPC coefficients ``(..., M+1) -> (..., M+1)``."""
    return nn.ParcorCoefficientsToLinearPredictiveCoefficients._func(k, gamma=gamma, c=c)


def mgc2mgc(mc: Tensor, out_order: int, in_alpha: float = 0, out_alpha: float = 0, in_gamma: float = 0,
            out_gamma: float = 0, in_norm: bool = False, out_norm: bool = False, in_mul: bool = False,
            out_mul: bool = False, n_fft: int = 512) -> Tensor:
    """Mel-generalized cepstrum conversion ``(..., M1+1) -> (..., M2+1)``."""
    return nn.MelGeneralizedCepstrumToMelGeneralizedCepstrum._func(
        mc, out_order=out_order, in_alpha=in_alpha, out_alpha=out_alpha, in_gamma=in_gamma, out_gamma=out_gamma,
        in_norm=in_norm, out_norm=out_norm, in_mul=in_mul, out_mul=out_mul, n_fft=n_fft)


def mgc2sp(mc: Tensor, fft_length: int, alpha: float = 0, gamma: float = 0, norm: bool = False, mul: bool = False,
           n_fft: int = 512, out_format: str = "power") -> Tensor:
    """Mel-generalized cepstrum to spectrum ``(..., M+1) -> (..., L/2+1)``."""
    return nn.MelGeneralizedCepstrumToSpectrum._func(mc, fft_length=fft_length, alpha=alpha, gamma=gamma, norm=norm,
                                                     mul=mul, n_fft=n_fft, out_format=out_format)


def plp(x: Tensor, plp_order: int, n_channel: int, sample_rate: int, compression_factor: float = 0.33,
        lifter: int = 1, f_min: float = 0, f_max: float | None = None, floor: float = 1e-5, gamma: float = 0,
        scale: str = "htk", erb_factor: float | None = None, n_fft: int = 512, out_format: str = "y") -> Tensor:
    """PLP analysis of a power spectrum ``(..., L/2+1) -> (..., M)`` [+ c0] [+ energy]."""
    return nn.PerceptualLinearPredictiveCoefficientsAnalysis._func(
        x, plp_order=plp_order, n_channel=n_channel, sample_rate=sample_rate,
        compression_factor=compression_factor, lifter=lifter, f_min=f_min, f_max=f_max, floor=floor, gamma=gamma,
        scale=scale, erb_factor=erb_factor, n_fft=n_fft, out_format=out_format)


def norm0(a: Tensor) -> Tensor:
    """All-pole to all-zero filter coefficients ``(..., M+1) -> (..., M+1)``."""
    return nn.AllPoleToAllZeroDigitalFilterCoefficients._func(a)


def fftcep(x: Tensor, cep_order: int, accel: float = 0, n_iter: int = 0) -> Tensor:
    """Cepstral analysis ``(..., L/2+1) -> (..., M+1)`` (improved cepstral method)."""
    return nn.CepstralAnalysis._func(x, cep_order=cep_order, accel=accel, n_iter=n_iter)


def ifftr(y: Tensor, out_length: int | None = None) -> Tensor:
    """Inverse real FFT, complex ``(..., L/2+1) -> (..., N)``."""
    return nn.RealValuedInverseFastFourierTransform._func(y, out_length=out_length)


def unframe(y: Tensor, out_length: int | None = None, *, frame_period: int = 80, center: bool = True,
            window: str = "rectangular", norm: str = "none", symmetric: bool = True) -> Tensor:
    """Windowed overlap-add ``(..., T/P, L) -> (..., T)``."""
    return nn.Unframe._func(y, out_length, frame_period=frame_period, center=center, window=window, norm=norm,
                            symmetric=symmetric)


def istft(y: Tensor, *, out_length: int | None = None, frame_length: int = 400, frame_period: int = 80,
          fft_length: int = 512, center: bool = True, window: str = "blackman", norm: str = "power",
          symmetric: bool = True) -> Tensor:
    """Inverse short-time Fourier transform, complex ``(..., T/P, N/2+1) -> (..., T)``, in one fused kernel."""
    return nn.InverseShortTimeFourierTransform._func(
        y, out_length, frame_length=frame_length, frame_period=frame_period, fft_length=fft_length, center=center,
        window=window, norm=norm, symmetric=symmetric)


# fused pipelines (one kernel chain from the waveform; see fused.py)
lpc_from_waveform = _fused.lpc_from_waveform
mfcc_from_waveform = _fused.mfcc_from_waveform
