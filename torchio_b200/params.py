"""Parameter specifications: scalar / (lo, hi) / per-axis / Choice / Distribution.

Host-side mirror of transforms/parameter_range.py (TorchIO 2.0.0a2).  What
matters for drop-in behaviour is the *RNG call sequence* on torch's global CPU
generator: with the same ``torch.manual_seed`` this
module draws the same numbers in the same order as the reference:
  - a plain number or a degenerate (v, v) range draws nothing;
  - a (lo, hi) range draws ``torch.empty(1).uniform_(lo, hi)`` per axis
    (parameter_range.py:103-106), or one size-n ``uniform_`` per axis in
    batched mode (:134-137);
  - ``Choice`` uses ``torch.multinomial`` (:59-74), a ``Distribution`` its own
    ``sample`` (:101-102,132-133).
"""

from __future__ import annotations

from collections.abc import Sequence

import torch
from torch.distributions import Distribution


class Choice:
    """Discrete set of values, optionally weighted."""

    def __init__(self, values: Sequence[float], probabilities: Sequence[float] | None = None):
        if len(values) < 1:
            raise ValueError("Choice requires at least one value")
        self._values = torch.tensor([float(v) for v in values])
        if probabilities is None:
            self._probs = torch.ones(len(values)) / len(values)
        else:
            if len(probabilities) != len(values):
                raise ValueError(
                    f"Expected {len(values)} probabilities, got {len(probabilities)}"
                )
            self._probs = torch.tensor([float(p) for p in probabilities])

    def sample(self) -> float:
        return float(self._values[int(torch.multinomial(self._probs, 1).item())])

    def sample_batched(self, n: int) -> torch.Tensor:
        return self._values[torch.multinomial(self._probs, n, replacement=True)]

    def __repr__(self) -> str:
        vals = ", ".join(f"{v:.1f}" if v == int(v) else f"{v}" for v in self._values.tolist())
        if torch.allclose(self._probs, self._probs[0].expand_as(self._probs)):
            return f"Choice([{vals}])"
        probs = ", ".join(f"{p:.2f}" for p in self._probs.tolist())
        return f"Choice([{vals}], p=[{probs}])"


def _is_number(x: object) -> bool:
    return isinstance(x, (int, float))


def _axis_spec(spec: object):
    if _is_number(spec):
        return float(spec)
    if isinstance(spec, (Choice, Distribution)):
        return spec
    if isinstance(spec, tuple) and len(spec) == 2 and all(_is_number(v) for v in spec):
        return (float(spec[0]), float(spec[1]))
    raise TypeError(
        "Per-axis spec must be a float, (lo, hi) tuple, Choice, or Distribution,"
        f" got {type(spec).__name__}"
    )


def _axes_from_tuple(value: tuple):
    n = len(value)
    if n == 3:
        if all(_is_number(v) for v in value):
            return tuple(float(v) for v in value)
        return tuple(_axis_spec(v) for v in value)
    if not all(_is_number(v) for v in value):
        raise ValueError(f"Mixed per-axis specs require exactly 3 elements, got {n}")
    if n == 1:
        return (float(value[0]),) * 3
    if n == 2:
        return ((float(value[0]), float(value[1])),) * 3
    if n == 6:
        return tuple((float(value[2 * a]), float(value[2 * a + 1])) for a in range(3))
    raise ValueError(f"Tuple must have 1, 2, 3, or 6 elements, got {n}")


def _draw(spec, generator=None) -> float:
    if _is_number(spec):
        return float(spec)
    if isinstance(spec, Choice):
        return spec.sample()
    if isinstance(spec, Distribution):
        return spec.sample().item()
    lo, hi = spec
    if lo == hi:
        return float(lo)
    return torch.empty(1).uniform_(float(lo), float(hi), generator=generator).item()


def _draw_n(spec, n: int, generator=None) -> torch.Tensor:
    if _is_number(spec):
        return torch.full((n,), float(spec))
    if isinstance(spec, Choice):
        return spec.sample_batched(n)
    if isinstance(spec, Distribution):
        return spec.sample((n,)).reshape(n).to(torch.float32)
    lo, hi = spec
    if lo == hi:
        return torch.full((n,), float(lo))
    return torch.empty(n).uniform_(float(lo), float(hi), generator=generator)


class _ParameterRange:
    """Three per-axis specs parsed from the user's value."""

    def __init__(self, value) -> None:
        self._original = value
        if _is_number(value):
            self._axes = (float(value),) * 3
        elif isinstance(value, (Choice, Distribution)):
            self._axes = (value,) * 3
        elif isinstance(value, tuple):
            self._axes = _axes_from_tuple(value)
        else:
            raise TypeError(
                f"Expected float, tuple, Distribution, or Choice, got {type(value).__name__}"
            )

    @property
    def is_deterministic(self) -> bool:
        return all(_is_number(a) for a in self._axes)

    def is_constant(self, value: float) -> bool:
        for a in self._axes:
            if _is_number(a):
                if float(a) != float(value):
                    return False
            elif isinstance(a, tuple):
                if not (a[0] == a[1] == value):
                    return False
            else:
                return False
        return True

    @property
    def _ranges(self):
        out = []
        for a in self._axes:
            if _is_number(a):
                out.append((float(a), float(a)))
            elif isinstance(a, tuple):
                out.append(a)
            else:
                out.append((0.0, 0.0))
        return tuple(out)

    @property
    def _distribution(self):
        return self._axes[0] if isinstance(self._axes[0], Distribution) else None

    def sample(self, n: int | None = None, *, generator=None):
        if n is None:
            return tuple(_draw(a, generator) for a in self._axes)
        return torch.stack([_draw_n(a, n, generator) for a in self._axes], dim=-1)

    def sample_1d(self, n: int | None = None, *, generator=None):
        if n is None:
            return _draw(self._axes[0], generator)
        return _draw_n(self._axes[0], n, generator)

    def __repr__(self) -> str:
        v = self._original
        if isinstance(v, tuple):
            return "(" + ", ".join(repr(x) for x in v) + ")"
        return repr(v) if isinstance(v, (Distribution, Choice)) else str(v)


def to_range(value) -> _ParameterRange:
    return _ParameterRange(value)


def to_nonneg_range(value) -> _ParameterRange:
    pr = _ParameterRange(value)
    if pr._distribution is None and any(lo < 0 or hi < 0 for lo, hi in pr._ranges):
        raise ValueError(f"Value must be non-negative, got {value}")
    return pr


class LazyParams(dict):
    """``params`` dict whose bulky entries (per-element matrices, control grids)
    are materialised to JSON-style nested lists on first access.

    The reference serialises everything eagerly (``control_points.cpu().tolist()``
    per element, transforms/spatial/spatial.py:2450-2468); for a batch of 32 that
    is ~33k Python floats per transform and dominates a step once the kernels are
    fast.  Reads (`[]`, ``get``, ``items``, ``values``, ``==``, ``repr``, pickling,
    ``json.dumps``) force the entry, so observable content is identical."""

    __slots__ = ("_lazy", "_packed")

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self._lazy = {}
        #: numpy form of bulky entries kept by ``make_params`` for ``apply_transform``
        #: (not part of the dict: invisible to ==, json, pickling)
        self._packed = None

    def set_lazy(self, key, thunk) -> None:
        dict.__setitem__(self, key, None)
        self._lazy[key] = thunk

    def _force(self, key=None) -> None:
        if not self._lazy:
            return
        for k in ([key] if key is not None else list(self._lazy)):
            thunk = self._lazy.pop(k, None)
            if thunk is not None:
                dict.__setitem__(self, k, thunk())

    def __getitem__(self, key):
        if key in self._lazy:
            self._force(key)
        return dict.__getitem__(self, key)

    def get(self, key, default=None):
        if key in self._lazy:
            self._force(key)
        return dict.get(self, key, default)

    def __setitem__(self, key, value):
        self._lazy.pop(key, None)
        dict.__setitem__(self, key, value)

    def __iter__(self):
        # a Python-level __iter__ takes `dict(params)`, `{**params}`, `a | params` and
        # `other.update(params)` off CPython's raw-storage fast path (dict_merge checks
        # tp_iter) and onto keys() + __getitem__, which force the lazy entries
        return dict.__iter__(self)

    def keys(self):
        return dict.keys(self)

    def pop(self, key, *default):
        if key in self._lazy:
            self._force(key)
        return dict.pop(self, key, *default)

    def popitem(self):
        self._force()
        return dict.popitem(self)

    def setdefault(self, key, default=None):
        if key in self._lazy:
            self._force(key)
        return dict.setdefault(self, key, default)

    def __or__(self, other):
        self._force()
        return dict(self) | other

    def __ror__(self, other):
        self._force()
        return other | dict(self)

    def items(self):
        self._force()
        return dict.items(self)

    def values(self):
        self._force()
        return dict.values(self)

    def copy(self):
        self._force()
        return dict(self)

    def __eq__(self, other):
        self._force()
        if isinstance(other, LazyParams):
            other._force()
        return dict.__eq__(self, other)

    def __ne__(self, other):
        return not self.__eq__(other)

    __hash__ = None

    def __repr__(self):
        self._force()
        return dict.__repr__(self)

    def __reduce__(self):
        self._force()
        return (dict, (dict(self),))

    def __deepcopy__(self, memo):
        import copy as _copy

        self._force()
        return _copy.deepcopy(dict(self), memo)


def slice_params(params, b0: int, b1: int):
    """Recorded ``params`` restricted to batch elements ``[b0, b1)``: per-instance
    entries (``_batched_keys``, ``_keep``) are sliced, shared entries kept.  The
    same rule the reference applies per element when a batch is unbatched
    (data/batch.py:365-399).  Lazy entries stay lazy."""
    batched = dict.get(params, "_batched_keys")
    if batched is None:
        return params
    out = LazyParams()
    lazy = getattr(params, "_lazy", {})
    for key in dict.keys(params):
        if key in lazy:
            if key in batched:
                out.set_lazy(key, lambda key=key: params[key][b0:b1])
            else:
                out.set_lazy(key, lambda key=key: params[key])
            continue
        value = dict.__getitem__(params, key)
        if key == "_batch_size":
            value = b1 - b0
        elif key == "_keep" and value is not None:
            value = value[b0:b1]
        elif key in batched and isinstance(value, list):
            value = value[b0:b1]
        dict.__setitem__(out, key, value)
    packed = getattr(params, "_packed", None)
    if packed is not None:
        first, second, per_instance = packed
        out._packed = (first[b0:b1], second[b0:b1], True) if per_instance else packed
    return out


def uniform_from_unit(u, lo: float, hi: float):
    """Map unit draws to ``uniform_(lo, hi)`` exactly as ATen's CPU kernel does:
    fp32 ``fma(u, float(hi) - float(lo), float(lo))`` (probe-verified against
    ``torch.empty(n).uniform_(lo, hi)``).  ``u``: float32 numpy array."""
    import numpy as np

    lo32 = np.float32(lo)
    span = np.float32(np.float32(hi) - lo32)
    # product of two fp32 values is exact in fp64; one rounding back to fp32
    return (u.astype(np.float64) * np.float64(span) + np.float64(lo32)).astype(np.float32)
