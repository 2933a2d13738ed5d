"""CPU oracle, part 1: the reference's op sequence restated on torch CPU ops.

TEST INFRASTRUCTURE — NOT PRODUCT CODE.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import
this module.  ``torchio_b200`` never does.

Where the arithmetic lives: the reference (TorchIO 2.0.0a2 @ 2b019d2) delegates
every hot-path computation to PyTorch ATen ops (``pyproject.toml:48`` lists
``torch`` unpinned; this image has torch 2.11.0+cu128, build 70d99e9).  This
file restates *which* ops the reference calls, in which order, on which
layouts, as plain functions over tensors and the reference's ``params``
dictionaries (the schema recorded in ``AppliedTransform.params``).  It is
pinned against the golden vectors in ``tests/golden/*.npz`` (outputs of the
unmodified reference, see ``tests/golden/generate.py``): on the same torch
build the results are bit-identical.

Each function cites the reference lines it follows (paths relative to
``/root/reference/src/torchio``).
"""

from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F

_MODES = {"nearest": "nearest", "linear": "bilinear"}


# ----------------------------------------------------------------------------
# Spatial (Affine / ElasticDeformation / Spatial)
# ----------------------------------------------------------------------------


def spacing_of(affine: np.ndarray) -> tuple[float, float, float]:
    """Column norms of the 3x3 block (data/affine.py:104-109)."""
    rz = torch.as_tensor(np.asarray(affine, dtype=np.float64))[:3, :3]
    sp = torch.sqrt(torch.sum(rz**2, dim=0))
    return (float(sp[0]), float(sp[1]), float(sp[2]))


def output_to_input_matrix(a_in, a_out, world_affine) -> torch.Tensor:
    """inv(A_in) @ inv(T) @ A_out in float64, cast to fp32.

    transforms/spatial/spatial.py:1582-1601.
    """
    inv_in = np.linalg.inv(np.asarray(a_in, dtype=np.float64))
    if world_affine is None:
        inv_t = np.eye(4, dtype=np.float64)
    else:
        inv_t = np.linalg.inv(np.asarray(world_affine, dtype=np.float64))
    m = inv_in @ inv_t @ np.asarray(a_out, dtype=np.float64)
    return torch.as_tensor(m, dtype=torch.float32)


def voxel_coordinates(shape) -> torch.Tensor:
    """(I, J, K, 3) fp32 meshgrid of indices (spatial.py:1604-1613)."""
    axes = [torch.arange(n, dtype=torch.float32) for n in shape]
    gi, gj, gk = torch.meshgrid(*axes, indexing="ij")
    return torch.stack([gi, gj, gk], dim=-1)


def map_homogeneous(coords: torch.Tensor, m: torch.Tensor) -> torch.Tensor:
    """[c, 1] @ M^T then drop w (spatial.py:1616-1624).  CPU sgemm."""
    ones = torch.ones(*coords.shape[:-1], 1, dtype=coords.dtype)
    return (torch.cat([coords, ones], dim=-1) @ m.T)[..., :3]


def upsample_control_points(cp: torch.Tensor, shape) -> torch.Tensor:
    """Trilinear align_corners=True upsample of (ni,nj,nk,3) mm field.

    spatial.py:2171-2189.  The permuted view is what the reference passes.
    """
    field = cp.permute(3, 0, 1, 2)[None].float()
    dense = F.interpolate(
        field, size=list(shape), mode="trilinear", align_corners=True
    )
    return dense[0].permute(1, 2, 3, 0)


def sampling_grid(in_shape, a_in, out_shape, a_out, world_affine, cp, affine_first):
    """Input-voxel coordinates of every output voxel (spatial.py:1504-1579)."""
    m = output_to_input_matrix(a_in, a_out, world_affine)
    coords = voxel_coordinates(out_shape)
    if cp is None:
        return map_homogeneous(coords, m)
    disp = upsample_control_points(
        torch.as_tensor(cp, dtype=torch.float32), out_shape
    )
    sp_out = torch.as_tensor(spacing_of(a_out), dtype=torch.float32)
    sp_in = torch.as_tensor(spacing_of(a_in), dtype=torch.float32)
    if affine_first:
        return map_homogeneous(coords, m) + disp / sp_in
    return map_homogeneous(coords + disp / sp_out, m)


def normalise_grid(vox: torch.Tensor, in_shape) -> torch.Tensor:
    """[-1, 1] grid in (K, J, I) order (spatial.py:1627-1648,1806-1824)."""
    sizes = torch.tensor(
        [max(in_shape[0] - 1, 1), max(in_shape[1] - 1, 1), max(in_shape[2] - 1, 1)],
        dtype=torch.float32,
    )
    g = 2.0 * vox / sizes - 1.0
    if g.ndim == 4:
        g = g[None]
    return g.permute(0, 3, 2, 1, 4)  # (B|1, K, J, I, 3)


def grid_sample_with_fill(data, vox, in_shape, mode, fill):
    """grid_sample(zeros, align_corners) + ones-mask fill.

    spatial.py:1695-1731 (shared grid) and :1827-1857 (per-sample grid).
    ``fill``: python float (0.0 => no mask step, spatial.py:2072-2076) or a
    1-D tensor of per-channel fills.
    """
    b = data.shape[0]
    grid = normalise_grid(vox, in_shape)
    if grid.shape[0] == 1 and b > 1:
        grid = grid.expand(b, -1, -1, -1, -1)
    x = data.permute(0, 1, 4, 3, 2).float()
    out = F.grid_sample(
        x, grid, mode=_MODES[mode], padding_mode="zeros", align_corners=True
    )
    fill_t = None
    if isinstance(fill, torch.Tensor):
        fill_t = fill.to(torch.float32)
        if fill_t.ndim == 1:
            fill_t = fill_t.reshape(1, -1, 1, 1, 1)
    elif float(fill) != 0.0:
        fill_t = torch.as_tensor(float(fill), dtype=torch.float32)
    if fill_t is not None:
        mask = F.grid_sample(
            torch.ones_like(x), grid, padding_mode="zeros", align_corners=True
        )
        out = torch.where(mask > 0.5, out, fill_t)
    return out.permute(0, 1, 4, 3, 2).to(data.dtype)


def fill_value_for(data, kind, pad_value, pad_label):
    """spatial.py:2034-2060 ('minimum' and numeric; sample 0, per channel)."""
    if kind == "label":
        return float(pad_label)
    if isinstance(pad_value, (int, float)):
        return float(pad_value)
    if pad_value == "minimum":
        return torch.as_tensor(
            [float(ch.min().item()) for ch in data[0]], dtype=torch.float32
        )
    if pad_value not in ("mean", "otsu"):
        raise NotImplementedError(pad_value)
    return torch.as_tensor(
        [border_mean(ch, pad_value == "otsu") for ch in data[0]], dtype=torch.float32
    )


def border_mean(tensor, filter_otsu):
    """spatial.py:2105-2131: mean of the six boundary faces, optionally only of the voxels
    below their Otsu threshold."""
    borders = torch.cat([
        tensor[0, :, :].ravel(), tensor[-1, :, :].ravel(), tensor[:, 0, :].ravel(),
        tensor[:, -1, :].ravel(), tensor[:, :, 0].ravel(), tensor[:, :, -1].ravel(),
    ]).float()
    if not filter_otsu:
        return float(borders.mean().item())
    values = borders[borders < otsu_threshold(borders)]
    return float(values.mean().item()) if values.numel() > 0 else float(borders.mean().item())


def otsu_threshold(values):
    """spatial.py:2133-2168: sweep over the sorted values maximising the between-class variance
    (python floats, running sums)."""
    sorted_values, _ = values.sort()
    n = sorted_values.numel()
    if n == 0:
        return 0.0
    total = float(sorted_values.sum().item())
    best_threshold, best_variance, background = float(sorted_values[0].item()), 0.0, 0.0
    for count, item in enumerate(sorted_values[:-1].tolist(), start=1):
        background += item
        mean_b = background / count
        mean_f = (total - background) / (n - count)
        variance = (count / n) * ((n - count) / n) * (mean_b - mean_f) ** 2
        if variance > best_variance:
            best_variance, best_threshold = variance, item
    return best_threshold


def label_partial_volume(data, vox, in_shape, a_in, a_out, antialias_on, one_hot_mode, pad_label):
    """``label_interpolation="label"`` (spatial.py:1275-1389): one-hot per distinct value,
    optional anti-alias blur, per-channel sampling with zero padding, argmax, pad label where
    the channel sum is not > 0.5.  C > 1: the channels are sampled without re-encoding."""
    if data.shape[1] > 1:
        smoothed = data.float()
        if antialias_on:
            smoothed = antialias(smoothed, a_in, a_out)
        sampled = grid_sample_with_fill(smoothed, vox, in_shape, one_hot_mode, 0.0)
        return sampled.to(data.dtype) if data.dtype.is_floating_point else sampled
    labels = torch.unique(data)
    one_hot = (data[:, 0][:, None] == labels.reshape(1, -1, 1, 1, 1)).float()
    if antialias_on:
        one_hot = antialias(one_hot, a_in, a_out)
    sampled = grid_sample_with_fill(one_hot, vox, in_shape, one_hot_mode, 0.0)
    winners = sampled.argmax(dim=1)
    resampled = labels[winners]
    in_bounds = sampled.sum(dim=1) > 0.5
    resampled = torch.where(in_bounds, resampled, torch.full_like(resampled, pad_label))
    return resampled[:, None].to(data.dtype)


def spatial(images: dict, params: dict) -> None:
    """Spatial.apply_transform with target=None (spatial.py:560-610,1110-1272).

    ``images``: name -> {"kind": "scalar"|"label", "data": (B,C,I,J,K),
    "affines": list of 4x4}.  Mutated in place.
    """
    names = params.get("selected_images", [])
    if not names:
        return
    per_instance = "affine_matrix" in (params.get("_batched_keys") or [])
    first = images[names[0]]
    shape = tuple(first["data"].shape[-3:])
    a0 = np.asarray(first["affines"][0], dtype=np.float64)
    target = params["target"]  # {"shape", "affine"} or None (spatial.py:1140-1142)
    out_shape = shape if target is None else tuple(int(v) for v in target["shape"])
    a_out = a0 if target is None else np.asarray(target["affine"], dtype=np.float64)
    if per_instance:
        mats = params["affine_matrix"]
        cps = params["control_points"]
        if target is None and all(m is None for m in mats) and all(c is None for c in cps):
            return
        grids = [
            sampling_grid(shape, a0, out_shape, a_out, mats[b], cps[b], params["affine_first"])
            for b in range(len(mats))
        ]
        grid = torch.stack(grids, 0)
        passthrough = [] if target is not None else [  # spatial.py:1171-1175
            b for b in range(len(mats)) if mats[b] is None and cps[b] is None
        ]
    else:
        mat, cp = params["affine_matrix"], params["control_points"]
        if target is None and mat is None and cp is None:
            return
        grid = sampling_grid(shape, a0, out_shape, a_out, mat, cp, params["affine_first"])
        passthrough = []
    for name in names:
        img = images[name]
        data = img["data"]
        mode = (
            params["label_interpolation"]
            if img["kind"] == "label"
            else params["image_interpolation"]
        )
        if img["kind"] == "label" and mode == "label":
            out = label_partial_volume(
                data, grid, shape, a0, a_out, params.get("antialias", False),
                params.get("one_hot_label_interpolation", "linear"), float(params["default_pad_label"]))
            if passthrough:
                out = out.contiguous()
                for b in passthrough:
                    out[b] = data[b]
            img["data"] = out
            img["affines"] = [
                img["affines"][b] if b in passthrough else a_out.copy()
                for b in range(len(img["affines"]))
            ]
            continue
        fill = fill_value_for(
            data, img["kind"], params["default_pad_value"], params["default_pad_label"]
        )
        source = data
        if params.get("antialias", False) and img["kind"] != "label":  # spatial.py:1256-1257
            source = antialias(data, a0, a_out)
        out = grid_sample_with_fill(source, grid, shape, mode, fill)
        if passthrough:
            out = out.contiguous()
            for b in passthrough:  # spatial.py:1101-1106
                out[b] = data[b]
        img["data"] = out
        img["affines"] = [
            img["affines"][b] if b in passthrough else a_out.copy()
            for b in range(len(img["affines"]))
        ]


def antialias_sigmas(factors, spacing):
    """Cardoso et al. (MICCAI 2015) sigma in voxels per downsampled axis (spatial.py:1951-1978)."""
    sigmas = np.zeros(3, dtype=np.float64)
    for axis in range(3):
        k = factors[axis]
        if k <= 1.0:
            continue
        variance = (k**2 - 1) * (2 * np.sqrt(2 * np.log(2))) ** (-2)
        sigma_mm = spacing[axis] * np.sqrt(variance)
        sigmas[axis] = sigma_mm / spacing[axis]
    return sigmas


def antialias(data, a_in, a_out):
    """_antialias_batch + _gaussian_smooth_batch (spatial.py:1921-2031): replicate pad + conv3d per
    downsampled axis with one shared kernel."""
    sp_in = np.asarray(spacing_of(a_in), dtype=np.float64)
    sigmas = antialias_sigmas(np.asarray(spacing_of(a_out), dtype=np.float64) / sp_in, sp_in)
    if np.all(sigmas == 0):
        return data
    result = data.float()
    b, c = result.shape[:2]
    for axis in range(3):
        sigma = float(sigmas[axis])
        if sigma <= 0:
            continue
        radius = max(int(np.ceil(3 * sigma)), 1)
        x = torch.arange(2 * radius + 1, dtype=torch.float32, device=data.device) - radius
        k = torch.exp(-0.5 * (x / sigma) ** 2)
        k = k / k.sum()
        k_shape = [1, 1, 1]
        k_shape[axis] = 2 * radius + 1
        pad = [0] * 6
        pad[2 * (2 - axis)] = radius
        pad[2 * (2 - axis) + 1] = radius
        padded = F.pad(result, pad, mode="replicate")
        result = F.conv3d(padded.reshape(b * c, 1, *padded.shape[2:]), k.reshape(1, 1, *k_shape), padding=0)
        result = result.reshape(b, c, *result.shape[2:])
    return result.to(data.dtype)


# ----------------------------------------------------------------------------
# BiasField
# ----------------------------------------------------------------------------


def coarse_shape(spatial_shape, scale):
    """intensity/bias_field.py:281,319 (python banker's round)."""
    return [max(round(s * scale), 4) for s in spatial_shape]


def coarse_bias_fields(shape, std, seed, scale):
    """Host draws of the coarse N(0, std) field(s).

    Per-instance: one CPU generator per element (bias_field.py:258-293);
    shared: one generator, (B, C, *small) in one draw (:316-329).
    """
    b, c = shape[0], shape[1]
    small = coarse_shape(shape[2:], scale)
    if isinstance(std, list):
        fields = []
        for s, sd in zip(std, seed, strict=True):
            g = torch.Generator(device="cpu")
            g.manual_seed(sd)
            fields.append(
                torch.normal(mean=0.0, std=s, size=(1, c, *small), generator=g, device="cpu")
            )
        return torch.cat(fields, 0)
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    return torch.normal(mean=0.0, std=std, size=(b, c, *small), generator=g, device="cpu")


def bias_field(images: dict, params: dict, divide: bool = False) -> None:
    """BiasField.apply_transform (bias_field.py:99-132,201-255,296-341)."""
    std, seed, scale = params["std"], params["seed"], params["scale"]
    per_instance = "_batched_keys" in params
    if not per_instance and std == 0:
        return
    for img in images.values():
        if img["kind"] != "scalar":
            continue
        data = img["data"]
        if per_instance:
            identity = [s == 0 for s in std]
            if all(identity):
                continue
            coarse = coarse_bias_fields(data.shape, std, seed, scale).to(data.device)
            field = torch.exp(
                F.interpolate(
                    coarse, size=list(data.shape[2:]), mode="trilinear",
                    align_corners=True,
                )
            )
            out = (data / field if divide else data * field).to(data.dtype)
            if any(identity):
                mask = torch.tensor(identity, dtype=torch.bool)
                out[mask] = data[mask]
        else:
            coarse = coarse_bias_fields(data.shape, std, seed, scale).to(data.device)
            field = torch.exp(
                F.interpolate(
                    coarse, size=list(data.shape[2:]), mode="trilinear",
                    align_corners=True,
                )
            )
            out = data / field if divide else data * field
        img["data"] = out


# ----------------------------------------------------------------------------
# Blur
# ----------------------------------------------------------------------------


def gaussian_taps_shared(sigma: float) -> tuple[torch.Tensor, int]:
    """intensity/blur.py:179-183."""
    radius = max(int(np.ceil(3 * sigma)), 1)
    x = torch.arange(2 * radius + 1, dtype=torch.float32) - radius
    k = torch.exp(-0.5 * (x / sigma) ** 2)
    return k / k.sum(), radius


def gaussian_taps_stacked(sigmas: np.ndarray) -> tuple[torch.Tensor, int]:
    """Per-element taps padded to the batch-max radius (blur.py:273-328)."""
    radii = np.zeros_like(sigmas, dtype=np.int64)
    pos = sigmas > 0
    radii[pos] = np.maximum(np.ceil(3 * sigmas[pos]).astype(np.int64), 1)
    rmax = int(radii.max())
    offs = (torch.arange(2 * rmax + 1, dtype=torch.float32) - rmax)[None]
    sig = torch.as_tensor(sigmas, dtype=torch.float32)[:, None]
    rad = torch.as_tensor(radii)[:, None]
    safe = torch.where(sig > 0, sig, torch.ones_like(sig))
    k = torch.exp(-0.5 * (offs / safe) ** 2)
    k = torch.where(offs.abs() <= rad, k, torch.zeros_like(k))
    delta = torch.zeros_like(k)
    delta[:, rmax] = 1.0
    k = torch.where(sig > 0, k, delta)
    return k / k.sum(dim=1, keepdim=True), rmax


def _conv_axis(x, taps, radius, axis, groups):
    pad = [0] * 6
    pad[2 * (2 - axis)] = radius
    pad[2 * (2 - axis) + 1] = radius
    shape = [1, 1, 1]
    shape[axis] = taps.shape[-1]
    w = taps.reshape(-1, 1, *shape)
    return F.conv3d(F.pad(x, pad, mode="replicate"), w, padding=0, groups=groups)


def gaussian_smooth(data: torch.Tensor, sigmas) -> torch.Tensor:
    """blur.py:129-252 (dispatch + both paths)."""
    sig = np.asarray(sigmas, dtype=np.float64)
    if np.all(sig <= 0):
        return data
    if sig.ndim == 2 and np.all(sig == sig[0]):
        sig = sig[0]
    b, c = data.shape[:2]
    out = data.float()
    if sig.ndim == 1:
        for axis in range(3):
            s = float(sig[axis])
            if s <= 0:
                continue
            taps, radius = gaussian_taps_shared(s)
            y = _conv_axis(out.reshape(b * c, 1, *out.shape[2:]), taps, radius, axis, 1)
            out = y.reshape(b, c, *y.shape[2:])
        return out.to(data.dtype)
    no_blur = np.all(sig <= 0, axis=1)
    for axis in range(3):
        col = sig[:, axis]
        if np.all(col <= 0):
            continue
        taps, radius = gaussian_taps_stacked(col)
        taps = taps.repeat_interleave(c, dim=0)
        y = _conv_axis(out.reshape(1, b * c, *out.shape[2:]), taps, radius, axis, b * c)
        out = y.reshape(b, c, *y.shape[2:])
    out = out.to(data.dtype)
    if no_blur.any():
        mask = torch.as_tensor(no_blur)
        out[mask] = data[mask]
    return out


def blur(images: dict, params: dict) -> None:
    """Blur.apply_transform (blur.py:75-126)."""
    per_instance = "_batched_keys" in params
    for img in images.values():
        if img["kind"] != "scalar":
            continue
        if per_instance:
            mm = np.asarray(params["std"], dtype=np.float64)
            sp = np.asarray([spacing_of(a) for a in img["affines"]], dtype=np.float64)
            vox = np.divide(mm, sp, out=np.zeros_like(mm), where=sp > 0)
        else:
            sp = np.asarray(spacing_of(img["affines"][0]), dtype=np.float64)
            vox = [s / q if q > 0 else 0.0 for s, q in zip(params["std"], sp)]
        img["data"] = gaussian_smooth(img["data"], vox)


# ----------------------------------------------------------------------------
# Noise / Gamma
# ----------------------------------------------------------------------------


def _per_element(value, ndim=5):
    if isinstance(value, list):
        return torch.tensor(value, dtype=torch.float32).reshape(-1, *([1] * (ndim - 1)))
    return value


def noise(images: dict, params: dict) -> None:
    """Noise.apply_transform (intensity/noise.py:98-123,166-178)."""
    g = torch.Generator(device="cpu")
    g.manual_seed(params["seed"])
    keep = params.get("_keep")
    mean, std = _per_element(params["mean"]), _per_element(params["std"])
    for img in images.values():
        if img["kind"] != "scalar":
            continue
        x = img["data"]
        # the reference draws on the CPU generator whatever the data's device (noise.py:166-178)
        n1 = mean + std * torch.randn(x.shape, generator=g, device="cpu").to(x.device)
        if params.get("rician", False):
            n2 = mean + std * torch.randn(x.shape, generator=g, device="cpu").to(x.device)
            y = torch.sqrt((x + n1) ** 2 + n2**2)
        else:
            y = x + n1
        if keep is not None:
            km = torch.tensor(keep, dtype=torch.bool).reshape(-1, 1, 1, 1, 1)
            y = torch.where(km, y, x)
        img["data"] = y


def gamma(images: dict, params: dict, invert: bool = False) -> None:
    """Gamma.apply_transform (intensity/gamma.py:80-120)."""
    lg = params["log_gamma"]
    if invert:
        lg = [-v for v in lg] if isinstance(lg, list) else -lg
    for img in images.values():
        if img["kind"] != "scalar":
            continue
        x = img["data"]
        if isinstance(lg, list):
            gam = torch.exp(torch.tensor(lg, dtype=torch.float32)).reshape(-1, 1, 1, 1, 1)
        else:
            gam = math.exp(lg)
        img["data"] = x.sign() * x.abs().pow(gam)


def standardize(images: dict, params: dict) -> None:
    """Standardize.apply_transform (intensity/standardize.py:81-94): the recorded (mean, std) of
    the first sample applied to every element."""
    for name, img in images.items():
        if img["kind"] != "scalar" or name not in params["stats"]:
            continue
        mean, std = params["stats"][name]
        if std == 0:
            raise RuntimeError(f'Standard deviation is zero for masked values in "{name}". Cannot standardize.')
        img["data"] = (img["data"].float() - mean) / std


def normalize(images: dict, params: dict) -> None:
    """Normalize.apply_transform (intensity/normalize.py:153-183): clamp to the recorded input
    range, map it onto [out_min, out_max] (per element when the params were sampled per instance)."""
    for name, img in images.items():
        if img["kind"] != "scalar":
            continue
        if "in_min" in params:
            in_min, in_max = params["in_min"], params["in_max"]
        else:
            if name not in params.get("in_ranges", {}):
                continue
            in_min, in_max = params["in_ranges"][name]
        in_range = in_max - in_min
        if in_range == 0:
            continue
        data = img["data"].float()
        if isinstance(params["out_min"], list):  # normalize.py:318-326
            lo = torch.tensor(params["out_min"], dtype=torch.float32, device=data.device).reshape(-1, 1, 1, 1, 1)
            hi = torch.tensor(params["out_max"], dtype=torch.float32, device=data.device).reshape(-1, 1, 1, 1, 1)
            out_min, out_range = lo, hi - lo
        else:
            out_min, out_range = params["out_min"], params["out_max"] - params["out_min"]
        data = data.clamp(in_min, in_max)
        img["data"] = (data - in_min) / in_range * out_range + out_min


def flip(images: dict, params: dict) -> None:
    """Flip.apply_transform (spatial/flip.py:186-263): data reversed along the sampled
    axes, per element when params are per-instance; affines untouched."""
    axes = params["axes"]
    per_instance = "_batched_keys" in params
    for img in images.values():
        data = img["data"]
        if per_instance:
            out = data.clone()
            for b, ax in enumerate(axes):
                if ax:
                    out[b] = torch.flip(data[b], [a - 3 for a in ax])
            img["data"] = out
        elif axes:
            img["data"] = torch.flip(data, [a - 3 for a in axes])


def _shift_origin(img: dict, voxels) -> None:
    shift = np.asarray(voxels, dtype=np.float64)
    for a in img["affines"]:
        a[:3, 3] += a[:3, :3] @ shift


def crop(images: dict, params: dict) -> None:
    """Crop.apply_transform (spatial/crop.py:77-101)."""
    i0, i1, j0, j1, k0, k1 = params["cropping"]
    for img in images.values():
        d = img["data"]
        si, sj, sk = d.shape[-3:]
        img["data"] = d[..., i0:si - i1 or None, j0:sj - j1 or None, k0:sk - k1 or None].contiguous()
        _shift_origin(img, (i0, j0, k0))


def _quantile(values, q):
    """compute_quantile (transforms/_statistics.py:11-45): kthvalue twice + lerp."""
    import math

    index = q * (values.numel() - 1)
    lower = math.floor(index)
    lower_value = torch.kthvalue(values, lower + 1).values
    if index == lower:
        return lower_value
    upper_value = torch.kthvalue(values, lower + 2).values
    return lower_value.lerp(upper_value, index - lower)


def pad(images: dict, params: dict) -> None:
    """Pad.apply_transform (spatial/pad.py:88-110, _padding.py:41-110): F.pad's modes, or a
    constant pad with one whole-volume statistic per batch element."""
    i0, i1, j0, j1, k0, k1 = params["padding"]
    mode = params["padding_mode"]
    pad_arg = (k0, k1, j0, j1, i0, i1)
    for img in images.values():
        data = img["data"]
        if mode not in ("mean", "median", "minimum"):
            img["data"] = torch.nn.functional.pad(data, pad_arg, mode=mode, value=params["fill"])
        else:
            flat = data.flatten(start_dim=1)
            if mode == "minimum":
                statistic = flat.amin(dim=1)
            else:
                float_flat = flat if data.dtype in (torch.float32, torch.float64) else flat.float()
                if mode == "mean":
                    statistic = float_flat.mean(dim=1)
                else:
                    statistic = torch.stack([_quantile(values, 0.5) for values in float_flat])
                statistic = statistic.to(data.dtype)
            padded = torch.nn.functional.pad(data, pad_arg)
            interior = torch.nn.functional.pad(
                torch.ones((1, 1, *data.shape[-3:]), dtype=torch.bool), pad_arg)
            img["data"] = torch.where(interior, padded, statistic.reshape(-1, 1, 1, 1, 1))
        _shift_origin(img, (-i0, -j0, -k0))


_APPLY = {
    "Flip": flip,
    "Crop": crop,
    "Pad": pad,
    "Spatial": spatial,
    "Resample": spatial,
    "Affine": spatial,
    "ElasticDeformation": spatial,
    "BiasField": bias_field,
    "Blur": blur,
    "Noise": noise,
    "Gamma": gamma,
    "Standardize": standardize,
    "Normalize": normalize,
}


def replay(images: dict, history: list[dict]) -> dict:
    """Apply recorded ``[{name, params}, ...]`` in order; returns ``images``."""
    for step in history:
        _APPLY[step["name"]](images, step["params"])
    return images
