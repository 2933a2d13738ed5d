"""Compose (host-side mirror of transforms/compose.py:38-98, TorchIO 2.0.0a2).

Semantics kept: deep-copy once at the top, children run with ``copy=False``
in order, each child draws its gate + params from the global CPU RNG exactly
where the reference does.  Difference by design: a CPU-resident batch is
staged to the GPU once for the whole pipeline (not once per child).
"""

from __future__ import annotations

import copy as _copy
from collections.abc import Mapping, Sequence
from typing import Any

import torch as _torch

from .base import Transform, _Staging, _finish, wrap_input
from . import intensity as _int


class Compose(Transform):
    def __init__(self, transforms: Sequence[Transform] | Mapping[str, Transform] | None = None,
                 *, copy: bool = True, **kwargs: Any) -> None:
        super().__init__(copy=copy, **kwargs)
        if transforms is None:
            self.transforms: list[Transform] = []
        elif isinstance(transforms, Mapping):
            self.transforms = list(transforms.values())
        else:
            self.transforms = list(transforms)

    def forward(self, data: Any) -> Any:
        if self.copy:
            data = _copy.deepcopy(data)
        batch, unwrap = wrap_input(data)
        with _Staging(batch):
            batch = self._forward_batch(batch)
        return _finish(batch, unwrap)

    def _forward_batch(self, batch):
        # Children never copy (compose.py:88-92).  Unlike the reference we do
        # not mutate child.copy: Queue calls one Compose from several threads.
        index = 0
        while index < len(self.transforms):
            group = self._fusable_run(index) if self.fuse else []
            if len(group) >= 2:
                batch = _run_fused(group, batch)
                index += len(group)
            else:
                batch = self.transforms[index]._forward_batch(batch)
                index += 1
        return batch

    #: Fuse runs of consecutive intensity transforms into one kernel pair.
    fuse = True

    def _fusable_run(self, start: int) -> list[Transform]:
        """Longest run from ``start`` of BiasField < Blur < Noise < Gamma (each at
        most once, in that order, same include/exclude): exactly the chains whose
        per-voxel arithmetic the fused kernel reproduces."""
        order = {_int.BiasField: 0, _int.Blur: 1, _int.Noise: 2, _int.Gamma: 3}
        run: list[Transform] = []
        last = -1
        for transform in self.transforms[start:]:
            rank = order.get(type(transform))
            if rank is None or rank <= last:
                break
            if run and (transform.include != run[0].include or transform.exclude != run[0].exclude):
                break
            run.append(transform)
            last = rank
        return run

    def __len__(self) -> int:
        return len(self.transforms)

    def __getitem__(self, index: int) -> Transform:
        return self.transforms[index]

    def to_hydra(self) -> dict[str, Any]:
        cfg = super().to_hydra()
        cfg["transforms"] = [t.to_hydra() for t in self.transforms]
        return cfg


def _run_fused(group: list[Transform], batch):
    """Sample every transform of the run exactly as sequential application would
    (gate draw, then make_params, in order — none of them reads voxel data),
    then apply all non-gated stages with one fused launch pair per image."""
    applied = []
    builders = []
    for transform in group:
        # same draws as Transform._forward_batch
        if not transform._per_instance_p_active(batch) and _torch.rand(1).item() >= transform.p:
            continue
        params = transform.make_params(batch)
        applied.append((transform, params))
        if isinstance(transform, _int.BiasField):
            builders.append(lambda ib, index, p=params: _int._bias_stage(
                ib.data.shape, ib.affines, p["std"], p["seed"], p["scale"], divide=False))
        elif isinstance(transform, _int.Blur):
            builders.append(lambda ib, index, p=params: _int._blur_stage(ib, p))
        elif isinstance(transform, _int.Noise):
            builders.append(_int._noise_stage_factory(params))
        else:
            builders.append(lambda ib, index, p=params: {
                "gamma": _int.tables.gamma_values(p["log_gamma"], ib.data.shape[0])})
    if applied:
        _int.run_stages(group[0]._get_images(batch), builders)
        for transform, params in applied:
            transform._record(batch, params)
    return batch
