"""Host-side helpers shared by the module mirror (the subset of the reference's
``diffsptk/utils/private.py`` that the hot path uses: get_layer :45-60,
filter_values :63-72, check_size :97-99)."""

from __future__ import annotations

from typing import Any, Callable


def check_size(actual: int, expected: int, what: str) -> None:
    if actual != expected:
        raise ValueError(f"Unexpected {what} (input {actual} vs target {expected}).")


def filter_values(scope: dict[str, Any], drop_keys: tuple[str, ...] | list[str] = ()) -> dict[str, Any]:
    """Constructor locals -> kwargs for ``_precompute`` (drops self/__class__ and ``drop_keys``)."""
    skip = {"self", "__class__", *drop_keys}
    return {k: v for k, v in scope.items() if k not in skip}


def get_layer(is_module: bool, cls, params: dict[str, Any]) -> Callable:
    """A sub-module instance (module path) or a closure over ``cls._func`` (functional path).

    On the functional path the leading size parameter is dropped for classes that infer it from
    the input, and learnable/device/dtype are dropped because they follow the input tensor.
    """
    if is_module:
        return cls(**params)
    items = list(params.items())
    if cls._takes_input_size:
        items = items[1:]
    kwargs = {k: v for k, v in items if k not in ("learnable", "device", "dtype")}

    def layer(*args, **extra):
        return cls._func(*args, **kwargs, **extra)

    return layer


PAD_MODE_IDS = {"constant": 0, "reflect": 1, "replicate": 2, "circular": 3}


def pad_mode_id(mode: str) -> int:
    try:
        return PAD_MODE_IDS[mode]
    except KeyError:
        raise ValueError(f"mode {mode} is not supported.") from None


def get_gamma(gamma: float, c: int | None) -> float:
    """``gamma`` itself, or ``-1 / c`` when the integer ``c`` is given (diffsptk/utils/private.py:233-238)."""
    if c is None or c == 0:
        return gamma
    if not 1 <= c:
        raise ValueError("c must be an integer greater than or equal to 1.")
    return -1 / c
