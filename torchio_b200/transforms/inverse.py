"""History -> inverse Compose (mirror of transforms/inverse.py:15-61)."""

from __future__ import annotations

import warnings
from typing import Any

from .base import _TRANSFORM_REGISTRY, IntensityTransform
from .compose import Compose


def get_inverse_transform(history: list[Any], *, warn: bool = True,
                          ignore_intensity: bool = False) -> Compose:
    steps = []
    for trace in reversed(history):
        cls = _TRANSFORM_REGISTRY.get(trace.name)
        if cls is None:
            if warn:
                warnings.warn(f"Unknown transform {trace.name!r} in history, skipping",
                              stacklevel=2)
            continue
        if ignore_intensity and issubclass(cls, IntensityTransform):
            continue
        instance = object.__new__(cls)
        if not instance.invertible:
            if warn:
                warnings.warn(f"{trace.name} is not invertible, skipping", stacklevel=2)
            continue
        inverse = instance.inverse(trace.params)
        inverse.include = trace.include
        inverse.exclude = trace.exclude
        steps.append(inverse)
    return Compose(steps)


def apply_inverse_transform(data: Any, *, warn: bool = True, ignore_intensity: bool = False):
    if not hasattr(data, "applied_transforms"):
        return data
    inverse = get_inverse_transform(
        data.applied_transforms, warn=warn, ignore_intensity=ignore_intensity
    )
    result = inverse(data)
    if hasattr(result, "applied_transforms"):
        result.applied_transforms = []
    return result
