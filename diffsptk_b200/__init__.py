"""diffsptk_b200 -- B200-native (sm_100a) drop-in for the frame-rate analysis path of diffsptk.

    Frame -> Window -> rFFT -> {Spectrum/STFT, LPC (acorr + levdur), freqt, mcep, fbank, MFCC}

Same nn.Module classes and functional signatures as ``diffsptk.modules`` / ``diffsptk.functional``
for that path; the arithmetic runs in hand-written CUDA kernels behind a C ABI
(``include/diffsptk_b200.h``).  CUDA tensors only: there is no CPU fallback.
"""

from . import functional  # noqa: F401
from .modules import *  # noqa: F401,F403
from .modules import __all__ as _module_names
from .fused import fuse, lpc_from_waveform, mfcc_from_waveform  # noqa: F401
from .version import __version__  # noqa: F401

__all__ = [*_module_names, "functional", "fuse", "lpc_from_waveform", "mfcc_from_waveform", "__version__"]
