"""Snapshot the reference's public signatures for the hot path (run in the build container).

Writes ``tests/golden/api_signatures.json``: constructor / functional parameter names, kinds and
defaults of the 37 exported classes (incl. aliases) and 29 functional delegates, plus the
``_takes_input_size`` flags -- the drop-in contract of SURVEY.md section 8(b).
"""

import inspect
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

CLASSES = ["Autocorrelation", "DiscreteCosineTransform", "DCT", "MelFilterBankAnalysis", "FBANK",
           "RealValuedFastFourierTransform", "Frame", "FrequencyTransform", "LevinsonDurbin",
           "LinearPredictiveCodingAnalysis", "LPC", "MelCepstralAnalysis",
           "MelFrequencyCepstralCoefficientsAnalysis", "MFCC", "Spectrum", "ShortTimeFourierTransform", "STFT",
           "Window", "RealValuedInverseFastFourierTransform", "Unframe", "InverseShortTimeFourierTransform",
           "ISTFT", "CepstralAnalysis", "Delta",
           "MLSADigitalFilterCoefficientsToMelCepstrum", "MelCepstrumToMLSADigitalFilterCoefficients",
           "GeneralizedCepstrumGainNormalization", "GeneralizedCepstrumInverseGainNormalization",
           "LinearPredictiveCoefficientsToParcorCoefficients", "ParcorCoefficientsToLinearPredictiveCoefficients",
           "AllPoleToAllZeroDigitalFilterCoefficients", "MelGeneralizedCepstrumToMelGeneralizedCepstrum",
           "MelGeneralizedCepstrumToSpectrum", "PerceptualLinearPredictiveCoefficientsAnalysis", "PLP",
           "MelGeneralizedCepstralAnalysis", "LinearPredictiveCoefficientsToLineSpectralPairs"]
FUNCTIONS = ["acorr", "dct", "fbank", "fftr", "frame", "freqt", "levdur", "lpc", "mcep", "mfcc", "spec", "stft",
             "window", "ifftr", "unframe", "istft", "fftcep", "delta", "b2mc", "mc2b", "gnorm", "ignorm", "lpc2par",
             "par2lpc", "norm0", "mgc2mgc", "mgc2sp", "plp", "lpc2lsp"]


def describe(fn):
    out = []
    for name, p in inspect.signature(fn).parameters.items():
        if name == "self":
            continue
        d = None if p.default is inspect.Parameter.empty else repr(p.default)
        out.append([name, p.kind.name, d])
    return out


def snapshot(pkg):
    snap = {"classes": {}, "functions": {}}
    for c in CLASSES:
        cls = getattr(pkg, c)
        snap["classes"][c] = {"init": describe(cls.__init__), "name": cls.__name__,
                              "takes_input_size": bool(getattr(cls, "_takes_input_size", False)),
                              "forward": describe(cls.forward)}
    for f in FUNCTIONS:
        snap["functions"][f] = describe(getattr(pkg.functional, f))
    return snap


if __name__ == "__main__":
    from oracle.ref_shim import load_reference

    D = load_reference()
    if D is None:
        raise SystemExit("reference not importable here")
    with open(os.path.join(HERE, "api_signatures.json"), "w") as f:
        json.dump(snapshot(D), f, indent=1, sort_keys=True)
    print("ok")
