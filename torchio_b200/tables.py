"""Host-side parameter tables: reference ``params`` dicts -> kernel inputs.

Everything here is tiny float64/fp32 host math (no voxels).  It follows the
reference's own host code so the device kernels start from bit-identical
numbers (TorchIO 2.0.0a2, paths relative to src/torchio/transforms/):
  output->input matrix   spatial/spatial.py:1582-1601
  Gaussian taps          intensity/blur.py:179-183 (shared), :292-328 (per element)
  coarse bias fields     intensity/bias_field.py:258-293, 316-329
  gamma = exp(log_gamma) intensity/gamma.py:103-120
"""

from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np
import torch

FLAG_PASSTHROUGH, FLAG_ELASTIC = 1, 2


# ---- spatial ---------------------------------------------------------------


def voxel_matrix(a_in: np.ndarray, a_out: np.ndarray, world) -> np.ndarray:
    """fp32 rows 0..2 of inv(A_in) @ inv(T) @ A_out (12 floats)."""
    inv_t = np.eye(4) if world is None else np.linalg.inv(np.asarray(world, dtype=np.float64))
    m = np.linalg.inv(a_in) @ inv_t @ a_out
    return m.astype(np.float32)[:3].reshape(12)


@dataclass
class SpatialTables:
    mat: np.ndarray            # (B, 12) fp32
    cp: np.ndarray | None      # (B, ni, nj, nk, 3) fp32
    flags: np.ndarray          # (B,) uint8
    passthrough: list[int]


def spatial_tables(affine_matrices, control_points, batch_size: int, a_in: np.ndarray,
                   a_out: np.ndarray, *, per_instance: bool, has_target: bool):
    """Pack per-element geometry.  ``affine_matrices``/``control_points`` are
    lists (per element) when ``per_instance`` else single values; entries may
    be None (identity).  Returns None for a complete no-op."""
    if not per_instance:
        affine_matrices = [affine_matrices] * batch_size
        control_points = [control_points] * batch_size
    if len(affine_matrices) != batch_size:
        raise RuntimeError(
            "Per-instance spatial parameters were recorded for"
            f" {len(affine_matrices)} elements but the batch has {batch_size}"
        )
    no_geometry = all(m is None for m in affine_matrices) and all(
        c is None for c in control_points
    )
    if no_geometry and not has_target:
        return None
    mat = np.empty((batch_size, 12), dtype=np.float32)
    flags = np.zeros(batch_size, dtype=np.uint8)
    rows = [b for b, w in enumerate(affine_matrices) if w is not None]
    if len(rows) < batch_size:
        identity = voxel_matrix(a_in, a_out, None)
        for b, w in enumerate(affine_matrices):
            if w is None:
                mat[b] = identity
    if rows:
        # stacked inv/matmul == per-element calls bit for bit (same LAPACK/BLAS kernels)
        worlds = np.stack([np.asarray(affine_matrices[b], dtype=np.float64) for b in rows])
        m = np.linalg.inv(a_in) @ np.linalg.inv(worlds) @ a_out
        mat[rows] = m.astype(np.float32)[:, :3].reshape(len(rows), 12)
    cp = None
    grids = [None if c is None else np.asarray(c, dtype=np.float32) for c in control_points]
    shapes = {g.shape for g in grids if g is not None}
    if shapes:
        if len(shapes) != 1:
            raise RuntimeError("control-point grids of one batch must share a shape")
        cp = np.zeros((batch_size, *shapes.pop()), dtype=np.float32)
        for b, g in enumerate(grids):
            if g is not None:
                cp[b] = g
                flags[b] |= FLAG_ELASTIC
    passthrough = []
    if per_instance and not has_target:
        for b in range(batch_size):
            if affine_matrices[b] is None and control_points[b] is None:
                flags[b] |= FLAG_PASSTHROUGH
                passthrough.append(b)
    return SpatialTables(mat, cp, flags, passthrough)


# ---- bias field --------------------------------------------------------------


def coarse_shape(spatial_shape, scale: float) -> list[int]:
    return [max(round(s * scale), 4) for s in spatial_shape]


def coarse_bias_fields(shape, std, seed, scale: float) -> torch.Tensor:
    """Host torch.normal draws from the recorded CPU-generator seeds."""
    b, c = int(shape[0]), int(shape[1])
    small = coarse_shape(shape[2:], scale)
    if isinstance(std, list):
        out = torch.empty((b, c, *small), dtype=torch.float32)
        for row, (s, sd) in enumerate(zip(std, seed, strict=True)):
            g = torch.Generator(device="cpu")
            g.manual_seed(int(sd))
            out[row] = torch.normal(mean=0.0, std=float(s), size=(1, c, *small), generator=g)[0]
        return out
    g = torch.Generator(device="cpu")
    g.manual_seed(int(seed))
    return torch.normal(mean=0.0, std=float(std), size=(b, c, *small), generator=g)


# ---- blur ---------------------------------------------------------------------


@dataclass
class BlurTables:
    taps: torch.Tensor       # (3, B, 2R+1) fp32
    radius: torch.Tensor     # (3, B) int32
    big_r: int
    axes_mask: int
    identity: torch.Tensor   # (B,) uint8


def _taps_shared(sigma: float) -> torch.Tensor:
    radius = max(int(np.ceil(3 * sigma)), 1)
    x = torch.arange(2 * radius + 1, dtype=torch.float32) - radius
    k = torch.exp(-0.5 * (x / sigma) ** 2)
    return k / k.sum()


def _taps_stacked(sigmas: np.ndarray) -> tuple[torch.Tensor, np.ndarray]:
    radii = np.zeros(len(sigmas), dtype=np.int64)
    pos = sigmas > 0
    radii[pos] = np.maximum(np.ceil(3 * sigmas[pos]).astype(np.int64), 1)
    rmax = int(radii.max())
    offs = (torch.arange(2 * rmax + 1, dtype=torch.float32) - rmax)[None]
    sig = torch.as_tensor(sigmas, dtype=torch.float32)[:, None]
    safe = torch.where(sig > 0, sig, torch.ones_like(sig))
    k = torch.exp(-0.5 * (offs / safe) ** 2)
    k = torch.where(offs.abs() <= torch.as_tensor(radii)[:, None], k, torch.zeros_like(k))
    delta = torch.zeros_like(k)
    delta[:, rmax] = 1.0
    k = torch.where(sig > 0, k, delta)
    return k / k.sum(dim=1, keepdim=True), radii


def blur_tables(sigmas_vox, batch_size: int) -> BlurTables | None:
    """Dispatch of _gaussian_smooth (blur.py:143-154): None = return input
    unchanged; 1-D or all-equal rows = shared taps; else per-element taps."""
    sig = np.asarray(sigmas_vox, dtype=np.float64)
    if np.all(sig <= 0):
        return None
    if sig.ndim == 2 and np.all(sig == sig[0]):
        sig = sig[0]
    rows: list[torch.Tensor | None] = []
    radius = torch.zeros((3, batch_size), dtype=torch.int32)
    if sig.ndim == 1:
        for axis in range(3):
            s = float(sig[axis])
            if s <= 0:
                rows.append(None)
                continue
            t = _taps_shared(s)
            radius[axis, :] = (t.numel() - 1) // 2
            rows.append(t[None].expand(batch_size, -1))
        identity = torch.zeros(batch_size, dtype=torch.uint8)
    else:
        if sig.shape[0] != batch_size:
            raise RuntimeError(
                f"Per-instance blur sigmas were recorded for {sig.shape[0]} elements"
                f" but the batch has {batch_size}"
            )
        for axis in range(3):
            col = sig[:, axis]
            if np.all(col <= 0):
                rows.append(None)
                continue
            t, radii = _taps_stacked(col)
            radius[axis] = torch.as_tensor(radii, dtype=torch.int32)
            rows.append(t)
        identity = torch.as_tensor(np.all(sig <= 0, axis=1)).to(torch.uint8)
    big_r = max((t.shape[1] - 1) // 2 for t in rows if t is not None)
    taps = torch.zeros((3, batch_size, 2 * big_r + 1), dtype=torch.float32)
    mask = 0
    for axis, t in enumerate(rows):
        if t is None:
            continue
        r = (t.shape[1] - 1) // 2
        taps[axis, :, big_r - r: big_r + r + 1] = t
        mask |= 1 << axis
    return BlurTables(taps, radius, big_r, mask, identity)


# ---- noise / gamma ---------------------------------------------------------


def per_element_vector(value, batch_size: int) -> np.ndarray:
    if isinstance(value, list):
        if len(value) != batch_size:
            raise RuntimeError(
                f"Per-instance parameters were recorded for {len(value)} elements"
                f" but the batch has {batch_size}"
            )
        return np.asarray(value, dtype=np.float32)
    return np.full(batch_size, value, dtype=np.float32)


def gamma_values(log_gamma, batch_size: int) -> np.ndarray:
    """exp(log_gamma): fp32 torch.exp per element, float64 math.exp when shared
    (gamma.py:117-120)."""
    if isinstance(log_gamma, list):
        if len(log_gamma) != batch_size:
            raise RuntimeError(
                f"Per-instance parameters were recorded for {len(log_gamma)} elements"
                f" but the batch has {batch_size}"
            )
        return torch.exp(torch.tensor(log_gamma, dtype=torch.float32)).numpy()
    return np.full(batch_size, math.exp(log_gamma), dtype=np.float32)
