"""Module protocol of the drop-in (mirrors diffsptk/modules/base.py:26-101).

A functional module is described by four static methods -- ``_check`` (validate), ``_precompute``
(parameters -> ``Precomputed``), ``_forward`` (the op itself, state passed by keyword) and ``_func``
(the functional entry: precompute + forward) -- so that composite modules and
``diffsptk_b200.functional`` can reuse one definition, exactly like the reference.
"""

from __future__ import annotations

from abc import ABC, abstractmethod
from typing import Any, Callable, ClassVar, NamedTuple

import torch
from torch import nn


class Precomputed(NamedTuple):
    values: dict[str, Any] = {}
    layers: dict[str, Callable] = {}
    tensors: dict[str, torch.Tensor] = {}


class BaseFunctionalModule(ABC, nn.Module):
    _takes_input_size: ClassVar[bool] = False  # first _precompute parameter is the input size
    _value_names: tuple[str, ...] = ()
    _layer_names: tuple[str, ...] = ()

    def _register_precomputed(self, pre: Precomputed, learnable: bool = False) -> None:
        self._value_names = tuple(pre.values)
        self._layer_names = tuple(pre.layers)
        for name, obj in {**pre.values, **pre.layers}.items():
            setattr(self, name, obj)
        for name, t in pre.tensors.items():
            if learnable:
                setattr(self, name, nn.Parameter(t))
            else:
                self.register_buffer(name, t, persistent=False)  # state_dict stays empty, as in the reference

    def _call_forward(self, *args) -> Any:
        state = {n: getattr(self, n) for n in (*self._value_names, *self._layer_names)}
        state.update(self._buffers)
        state.update(self._parameters)
        return self._forward(*args, **state)

    @classmethod
    def _apply_precomputed(cls, pre: Precomputed, **inputs: Any) -> Any:
        return cls._forward(**inputs, **pre.values, **pre.layers, **pre.tensors)

    @staticmethod
    @abstractmethod
    def _func(*args, **kwargs) -> Any: ...

    @staticmethod
    @abstractmethod
    def _check(*args, **kwargs) -> None: ...

    @staticmethod
    @abstractmethod
    def _precompute(*args, **kwargs) -> Precomputed: ...

    @staticmethod
    @abstractmethod
    def _forward(*args, **kwargs) -> Any: ...


# Differences from the reference's base class are deliberate and small: buffers are registered non-persistent
# exactly like the reference (state_dict stays empty), but sub-layers and values are plain attributes collected
# by name, so that fused forwards (STFT, ISTFT, MFCC, LPC) can read a sub-layer's table (``self.window.window``)
# without calling the sub-layer.
