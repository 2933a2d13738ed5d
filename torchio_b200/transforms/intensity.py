"""BiasField / Blur / Noise / Gamma behind the reference API.

Host-side mirror of transforms/intensity/{bias_field,blur,noise,gamma}.py
(TorchIO 2.0.0a2): constructor signatures, ``make_params`` RNG order and the
``params`` schema are the reference's; ``apply_transform`` runs the CUDA
kernels K2-K5 (`torchio_b200.ops`).
"""

from __future__ import annotations

import os
from typing import Any

import numpy as np
import torch
from torch import Tensor

from .. import ops, tables
from ..data import SubjectsBatch
from ..params import to_nonneg_range, to_range
from .base import IntensityTransform


def _as_f32(data: Tensor) -> Tensor:
    return data if data.dtype == torch.float32 else data.float()


# ---- BiasField (intensity/bias_field.py:22-197) ------------------------------


class BiasField(IntensityTransform):
    def __init__(self, *, std=0.5, scale: float = 0.025, **kwargs: Any) -> None:
        super().__init__(**kwargs)
        self.std = to_nonneg_range(std)
        if scale <= 0 or scale > 1:
            raise ValueError(f"scale must be in (0, 1], got {scale}")
        self.scale = scale

    @property
    def supports_per_instance_params(self) -> bool:
        return True

    @property
    def supports_per_instance_p(self) -> bool:
        return True

    def make_params(self, batch: SubjectsBatch) -> dict[str, Any]:
        n = self._resolve_n(batch)
        if n is None:
            std = self.std.sample_1d()
            seed = int(torch.randint(0, 2**31, (1,)).item())
            return {"std": std, "seed": seed, "scale": self.scale}
        keep = self._keep_mask(batch, n)
        std = self._mask_identity(self.std.sample_1d(n), keep, identity=0.0)
        seeds = [int(torch.randint(0, 2**31, (1,)).item()) for _ in range(n)]
        params = {"std": self._serialize_param(std), "seed": seeds, "scale": self.scale}
        self._tag_batched(params, batch, n, keep, ["std", "seed"])
        return params

    def apply_transform(self, batch: SubjectsBatch, params: dict[str, Any]) -> SubjectsBatch:
        _apply_bias(self, batch, params["std"], params["seed"], params["scale"], divide=False)
        return batch

    @property
    def invertible(self) -> bool:
        return True

    def inverse(self, params: dict[str, Any]) -> _BiasFieldInverse:
        return _BiasFieldInverse(std=params["std"], seed=params["seed"], scale=params["scale"],
                                 copy=False)


def _apply_bias(transform, batch, std, seed, scale, *, divide: bool) -> None:
    per_element = isinstance(std, list)
    if not per_element and std == 0:
        return
    if per_element and all(s == 0 for s in std):
        return
    for ib in transform._get_images(batch).values():
        data = ib.data
        b = data.shape[0]
        if per_element and len(std) != b:
            raise RuntimeError(
                f"Per-instance parameters were recorded for {len(std)} elements"
                f" but the batch has {b}"
            )
        coarse = tables.coarse_bias_fields(data.shape, std, seed, scale)
        identity = np.asarray([s == 0 for s in std], dtype=np.uint8) if per_element else None
        coarse_d, identity_d = ops.upload(data.device, coarse, identity)
        out = ops.bias_field(_as_f32(data), coarse_d, identity_d, divide=divide)
        ib.data = out if per_element is False or out.dtype == data.dtype else out.to(data.dtype)


class _BiasFieldInverse(IntensityTransform):
    def __init__(self, *, std, seed, scale: float, **kwargs: Any) -> None:
        super().__init__(**kwargs)
        self._std, self._seed, self._scale = std, seed, scale

    def apply_transform(self, batch: SubjectsBatch, params: dict[str, Any]) -> SubjectsBatch:
        _apply_bias(self, batch, self._std, self._seed, self._scale, divide=True)
        return batch


# ---- Blur (intensity/blur.py:19-126) ------------------------------------------


class Blur(IntensityTransform):
    def __init__(self, *, std=0.0, **kwargs: Any) -> None:
        super().__init__(**kwargs)
        self.std = to_nonneg_range(std)
        self._warn_if_noop(is_noop=self.std.is_constant(0.0), hint="std=(0, 2)")

    @property
    def supports_per_instance_params(self) -> bool:
        return True

    @property
    def supports_per_instance_p(self) -> bool:
        return True

    def make_params(self, batch: SubjectsBatch) -> dict[str, Any]:
        n = self._resolve_n(batch)
        if n is None:
            return {"std": self.std.sample()}
        keep = self._keep_mask(batch, n)
        std = self.std.sample(n)
        if keep is not None:
            std[~keep] = 0.0
        params = {"std": self._serialize_param(std)}
        self._tag_batched(params, batch, n, keep, ["std"])
        return params

    def apply_transform(self, batch: SubjectsBatch, params: dict[str, Any]) -> SubjectsBatch:
        per_instance = self._is_per_instance_params(params)
        for ib in self._get_images(batch).values():
            data = ib.data
            if per_instance:
                mm = np.asarray(params["std"], dtype=np.float64)
                sp = np.asarray([a.spacing for a in ib.affines], dtype=np.float64)
                vox = np.divide(mm, sp, out=np.zeros_like(mm), where=sp > 0)
            else:
                sp = np.asarray(ib.affines[0].spacing, dtype=np.float64)
                vox = [s / q if q > 0 else 0.0 for s, q in zip(params["std"], sp, strict=True)]
            t = tables.blur_tables(vox, data.shape[0])
            if t is None:  # all sigma <= 0: the input tensor itself (blur.py:143-144)
                continue
            taps, radius, identity = ops.upload(data.device, t.taps, t.radius, t.identity)
            out = ops.blur(_as_f32(data), taps, radius, t.big_r, t.axes_mask, identity)
            ib.data = out if out.dtype == data.dtype else out.to(data.dtype)
        return batch


# ---- Noise (intensity/noise.py:18-178) -----------------------------------------


def _noise_mode() -> str:
    """"exact": normals are torch.randn draws of the recorded CPU-generator seed
    (the reference's stream; generated on the host and uploaded, as the
    reference itself does on a GPU batch, noise.py:177).  "philox": in-kernel
    counter-based normals — same distribution, different stream."""
    mode = os.environ.get("TIO_B200_NOISE", "exact").lower()
    if mode not in ("exact", "philox"):
        raise ValueError(f"TIO_B200_NOISE must be 'exact' or 'philox', got {mode!r}")
    return mode


class Noise(IntensityTransform):
    def __init__(self, *, mean=0.0, std=0.25, rician: bool = False, **kwargs: Any) -> None:
        super().__init__(**kwargs)
        self.mean = to_range(mean)
        self.std = to_nonneg_range(std)
        self.rician = rician

    @property
    def supports_per_instance_params(self) -> bool:
        return True

    @property
    def supports_per_instance_p(self) -> bool:
        return True

    def make_params(self, batch: SubjectsBatch) -> dict[str, Any]:
        seed = int(torch.randint(0, 2**31, (1,)).item())  # drawn first (noise.py:75)
        n = self._resolve_n(batch)
        keep = self._keep_mask(batch, n)
        mean = self._mask_identity(self.mean.sample_1d(n), keep, identity=0.0)
        std = self._mask_identity(self.std.sample_1d(n), keep, identity=0.0)
        params = {
            "mean": self._serialize_param(mean), "std": self._serialize_param(std),
            "seed": seed, "rician": self.rician,
        }
        self._tag_batched(params, batch, n, keep, ["mean", "std"])
        return params

    def apply_transform(self, batch: SubjectsBatch, params: dict[str, Any]) -> SubjectsBatch:
        rician = params.get("rician", False)
        keep = params.get("_keep")
        mode = _noise_mode()
        generator = torch.Generator(device="cpu")
        generator.manual_seed(params["seed"])
        for index, ib in enumerate(self._get_images(batch).values()):
            data = ib.data
            b = data.shape[0]
            mean = tables.per_element_vector(params["mean"], b)
            std = tables.per_element_vector(params["std"], b)
            keep_np = None if keep is None else np.asarray(keep, dtype=np.uint8)
            mean_d, std_d, keep_d = ops.upload(data.device, mean, std, keep_np)
            x = _as_f32(data)
            if mode == "philox":
                seed = (int(params["seed"]) << 8) | (index & 0xFF)
                ib.data = ops.noise_philox(x, mean_d, std_d, keep_d, seed, rician)
                continue
            # one CPU generator, consumed in flat (B,C,I,J,K) order, continuing
            # across images and across the second Rician draw (noise.py:166-178)
            pin = torch.cuda.is_available()
            z = torch.empty(data.shape, dtype=torch.float32, pin_memory=pin)
            torch.randn(data.shape, generator=generator, out=z)
            z2 = None
            if rician:
                z2 = torch.empty(data.shape, dtype=torch.float32, pin_memory=pin)
                torch.randn(data.shape, generator=generator, out=z2)
                z2 = z2.to(data.device, non_blocking=True)
            ib.data = ops.noise(x, mean_d, std_d, keep_d, z.to(data.device, non_blocking=True), z2)
        return batch


# ---- Gamma (intensity/gamma.py:17-149) ------------------------------------------


class Gamma(IntensityTransform):
    def __init__(self, *, log_gamma=0.0, **kwargs: Any) -> None:
        super().__init__(**kwargs)
        self.log_gamma = to_range(log_gamma)
        self._warn_if_noop(is_noop=self.log_gamma.is_constant(0.0), hint="log_gamma=(-0.3, 0.3)")

    @property
    def supports_per_instance_params(self) -> bool:
        return True

    @property
    def supports_per_instance_p(self) -> bool:
        return True

    def make_params(self, batch: SubjectsBatch) -> dict[str, Any]:
        n = self._resolve_n(batch)
        keep = self._keep_mask(batch, n)
        log_gamma = self._mask_identity(self.log_gamma.sample_1d(n), keep, identity=0.0)
        params = {"log_gamma": self._serialize_param(log_gamma)}
        self._tag_batched(params, batch, n, keep, ["log_gamma"])
        return params

    def apply_transform(self, batch: SubjectsBatch, params: dict[str, Any]) -> SubjectsBatch:
        _apply_gamma(self, batch, params["log_gamma"])
        return batch

    @property
    def invertible(self) -> bool:
        return True

    def inverse(self, params: dict[str, Any]) -> _GammaInverse:
        return _GammaInverse(log_gamma=params["log_gamma"], copy=False)


def _apply_gamma(transform, batch, log_gamma) -> None:
    for ib in transform._get_images(batch).values():
        data = ib.data
        gam = tables.gamma_values(log_gamma, data.shape[0])
        (gam_d,) = ops.upload(data.device, gam)
        ib.data = ops.gamma(_as_f32(data), gam_d)


class _GammaInverse(IntensityTransform):
    def __init__(self, *, log_gamma, **kwargs: Any) -> None:
        super().__init__(**kwargs)
        self._log_gamma = log_gamma

    def apply_transform(self, batch: SubjectsBatch, params: dict[str, Any]) -> SubjectsBatch:
        lg = self._log_gamma
        _apply_gamma(self, batch, [-v for v in lg] if isinstance(lg, list) else -lg)
        return batch
