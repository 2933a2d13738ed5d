"""Undoing a recorded history (behaviour of transforms/inverse.py:15-100, TorchIO 2.0.0a2).

A history is a list of `AppliedTransform(name, params, include, exclude)`.  Undoing it means:
newest record first, look the class up by name, skip what cannot be inverted (unknown names
and non-invertible classes with a warning, intensity transforms silently when
``ignore_intensity``), and run the remaining inverses as one `Compose` — so a run of
`_BiasFieldInverse` / `_GammaInverse` steps goes through the same kernels as the forward chain.
`plan_inverse` exposes the decisions; the two public functions keep the reference's signatures.
"""

from __future__ import annotations

import warnings
from dataclasses import dataclass, field
from typing import Any

from .base import _TRANSFORM_REGISTRY, IntensityTransform, Transform
from .compose import Compose

_UNKNOWN = "Unknown transform {name!r} in history, skipping"
_ONE_WAY = "{name} is not invertible, skipping"


@dataclass
class InversePlan:
    """Inverse transforms in application order plus the records that were left out and why."""

    steps: list[Transform] = field(default_factory=list)
    skipped: list[tuple[str, str]] = field(default_factory=list)  # (record name, reason)

    def compose(self) -> Compose:
        # copy=True (Compose's default): undoing must not mutate the caller's data in place
        return Compose(self.steps)


def _inverse_of(record: Any) -> tuple[Transform | None, str | None]:
    """(inverse transform, None) for an invertible record, else (None, reason)."""
    cls = _TRANSFORM_REGISTRY.get(record.name)
    if cls is None:
        return None, "unknown"
    # `invertible` and `inverse(params)` are defined on the class and need no constructor state:
    # the reference calls them on a bare instance as well
    bare = object.__new__(cls)
    if not bare.invertible:
        return None, "one-way"
    undo = bare.inverse(record.params)
    undo.include, undo.exclude = record.include, record.exclude
    return undo, None


def plan_inverse(history: list[Any], *, ignore_intensity: bool = False) -> InversePlan:
    plan = InversePlan()
    for record in history[::-1]:
        cls = _TRANSFORM_REGISTRY.get(record.name)
        if ignore_intensity and cls is not None and issubclass(cls, IntensityTransform):
            plan.skipped.append((record.name, "intensity"))
            continue
        undo, reason = _inverse_of(record)
        if undo is None:
            plan.skipped.append((record.name, reason))
        else:
            plan.steps.append(undo)
    return plan


def get_inverse_transform(history: list[Any], *, warn: bool = True,
                          ignore_intensity: bool = False) -> Compose:
    plan = plan_inverse(history, ignore_intensity=ignore_intensity)
    if warn:
        messages = {"unknown": _UNKNOWN, "one-way": _ONE_WAY}
        for name, reason in plan.skipped:
            if reason in messages:
                warnings.warn(messages[reason].format(name=name), stacklevel=2)
    return plan.compose()


def apply_inverse_transform(data: Any, *, warn: bool = True, ignore_intensity: bool = False):
    history = getattr(data, "applied_transforms", None)
    if history is None:  # plain tensors / arrays carry no history: nothing to undo
        return data
    undone = get_inverse_transform(history, warn=warn, ignore_intensity=ignore_intensity)(data)
    if hasattr(undone, "applied_transforms"):
        undone.applied_transforms = []
    return undone
