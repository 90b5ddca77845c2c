"""Delta (regression) features over the frame axis (drop-in for diffsptk/modules/delta.py)."""

from __future__ import annotations

import torch

from .. import ops, tables
from ..utils import filter_values
from .base import BaseFunctionalModule, Precomputed


class Delta(BaseFunctionalModule):
    """``(B, T, D)`` or ``(T, D)`` ``-> (..., T, D x H)``; kernel ``dsb200_delta``.

    The reference pads the frame axis by replication and runs a 2-D convolution with an ``(H, 1, W, 1)`` kernel,
    then permutes (delta.py:172-194).  Here one thread per (frame, feature) pair walks the ``W`` taps of all ``H``
    windows with clamped frame indices and writes the ``H`` outputs in the reference's ``[static | delta | ...]``
    layout: the input is read once from HBM (the halo rows hit the L1/L2), the output written once, no
    permute.  Buffer name ``window`` as in the reference.
    """

    def __init__(self, seed=[[-0.5, 0, 0.5], [1, -2, 1]], static_out: bool = True,  # noqa: B006 (reference default)
                 device: torch.device | None = None, dtype: torch.dtype | None = None) -> None:
        super().__init__()
        self._register_precomputed(self._precompute(**filter_values(locals())))

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self._call_forward(x)

    @staticmethod
    def _func(x: torch.Tensor, *args, **kwargs) -> torch.Tensor:
        pre = Delta._precompute(*args, **kwargs, device=x.device, dtype=x.dtype)
        return Delta._apply_precomputed(pre, x=x)

    @staticmethod
    def _check(seed) -> None:
        if not isinstance(seed, (tuple, list)):
            raise ValueError("seed must be tuple or list.")

    @staticmethod
    def _precompute(seed, static_out: bool, device: torch.device | None, dtype: torch.dtype | None) -> Precomputed:
        Delta._check(seed)
        if dtype is not None and not dtype.is_floating_point:
            dtype = None
        return Precomputed(tensors={"window": tables.make_delta_window(seed, static_out, device, dtype)})

    @staticmethod
    def _forward(x: torch.Tensor, *, window: torch.Tensor) -> torch.Tensor:
        if x.dim() not in (2, 3):
            raise ValueError("Input must be 2D or 3D tensor.")
        return ops.delta(x, window)
