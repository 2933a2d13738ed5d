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

from .base import Transform, _Staging, _finish, wrap_input


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
        for transform in self.transforms:
            batch = transform._forward_batch(batch)
        return batch

    def __len__(self) -> int:
        return len(self.transforms)

    def __getitem__(self, index: int) -> Transform:
        return self.transforms[index]

    def to_hydra(self) -> dict[str, Any]:
        cfg = super().to_hydra()
        cfg["transforms"] = [t.to_hydra() for t in self.transforms]
        return cfg
