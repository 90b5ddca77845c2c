"""nn.Module mirror of the hot-path subset of ``diffsptk.modules`` (same names and aliases as
diffsptk/modules/__init__.py:17-175)."""

from .acorr import Autocorrelation
from .b2mc import MLSADigitalFilterCoefficientsToMelCepstrum
from .dct import DiscreteCosineTransform
from .dct import DiscreteCosineTransform as DCT
from .delta import Delta
from .fbank import MelFilterBankAnalysis
from .fftcep import CepstralAnalysis
from .fbank import MelFilterBankAnalysis as FBANK
from .fftr import RealValuedFastFourierTransform
from .frame import Frame
from .ifftr import RealValuedInverseFastFourierTransform
from .istft import InverseShortTimeFourierTransform
from .istft import InverseShortTimeFourierTransform as ISTFT
from .freqt import FrequencyTransform
from .gnorm import GeneralizedCepstrumGainNormalization
from .ignorm import GeneralizedCepstrumInverseGainNormalization
from .levdur import LevinsonDurbin
from .lpc import LinearPredictiveCodingAnalysis
from .lpc import LinearPredictiveCodingAnalysis as LPC
from .lpc2lsp import LinearPredictiveCoefficientsToLineSpectralPairs
from .lpc2par import LinearPredictiveCoefficientsToParcorCoefficients
from .mc2b import MelCepstrumToMLSADigitalFilterCoefficients
from .mcep import MelCepstralAnalysis
from .mfcc import MelFrequencyCepstralCoefficientsAnalysis
from .mgc2mgc import MelGeneralizedCepstrumToMelGeneralizedCepstrum
from .mgc2sp import MelGeneralizedCepstrumToSpectrum
from .mgcep import MelGeneralizedCepstralAnalysis
from .plp import PerceptualLinearPredictiveCoefficientsAnalysis
from .plp import PerceptualLinearPredictiveCoefficientsAnalysis as PLP
from .mfcc import MelFrequencyCepstralCoefficientsAnalysis as MFCC
from .norm0 import AllPoleToAllZeroDigitalFilterCoefficients
from .par2lpc import ParcorCoefficientsToLinearPredictiveCoefficients
from .spec import Spectrum
from .stft import ShortTimeFourierTransform
from .stft import ShortTimeFourierTransform as STFT
from .unframe import Unframe
from .window import Window

__all__ = [
    "Autocorrelation", "DiscreteCosineTransform", "DCT", "MelFilterBankAnalysis", "FBANK",
    "RealValuedFastFourierTransform", "Frame", "FrequencyTransform", "LevinsonDurbin",
    "LinearPredictiveCodingAnalysis", "LPC", "MelCepstralAnalysis",
    "MelFrequencyCepstralCoefficientsAnalysis", "MFCC", "Spectrum", "ShortTimeFourierTransform", "STFT",
    "Window", "RealValuedInverseFastFourierTransform", "Unframe", "InverseShortTimeFourierTransform", "ISTFT", "CepstralAnalysis", "Delta",
    "MLSADigitalFilterCoefficientsToMelCepstrum", "MelCepstrumToMLSADigitalFilterCoefficients",
    "GeneralizedCepstrumGainNormalization", "GeneralizedCepstrumInverseGainNormalization",
    "LinearPredictiveCoefficientsToParcorCoefficients", "ParcorCoefficientsToLinearPredictiveCoefficients",
    "AllPoleToAllZeroDigitalFilterCoefficients", "MelGeneralizedCepstrumToMelGeneralizedCepstrum",
    "MelGeneralizedCepstrumToSpectrum", "PerceptualLinearPredictiveCoefficientsAnalysis", "PLP",
    "MelGeneralizedCepstralAnalysis", "LinearPredictiveCoefficientsToLineSpectralPairs",
]
