"""Helpers shared by the parity tests (test infrastructure)."""

from __future__ import annotations

import json
from pathlib import Path

import numpy as np
import torch

from golden_cases import CASES_BY_NAME, build_inputs

GOLDEN = Path(__file__).resolve().parent / "golden"


def load_golden(name):
    """Return (case, images-in-oracle-format, history, expected-outputs)."""
    case = CASES_BY_NAME[name]
    inputs = build_inputs(case)
    names = list(inputs["subjects"][0].keys())
    images = {}
    for n in names:
        kind = inputs["subjects"][0][n][0]
        data = torch.stack([s[n][1] for s in inputs["subjects"]])
        affines = [np.array(s[n][2], dtype=np.float64) for s in inputs["subjects"]]
        images[n] = {"kind": kind, "data": data, "affines": affines}
    z = np.load(GOLDEN / f"{name}.npz")
    history = json.loads(bytes(z["history"]).decode())
    expected = {n: torch.from_numpy(z[f"out_{n}"]) for n in names}
    expected_aff = {n: z[f"aff_{n}"] for n in names}
    return case, images, history, expected, expected_aff


def report(actual: torch.Tensor, expected: torch.Tensor) -> dict:
    """Max-abs / dynamic range and mismatch statistics."""
    a = actual.double()
    e = expected.double()
    rng = float(e.max() - e.min()) or 1.0
    diff = (a - e).abs()
    return {
        "max_abs": float(diff.max()),
        "max_abs_over_range": float(diff.max()) / rng,
        "frac_gt_1e-4_range": float((diff > 1e-4 * rng).double().mean()),
        "n_mismatch": int((a != e).sum()),
        "n": a.numel(),
    }


# ---- product-side helpers -----------------------------------------------------


def product_batch(images, device=None):
    """Oracle-format images dict -> torchio_b200.SubjectsBatch."""
    import torchio_b200 as tio

    batches = {}
    for name, img in images.items():
        cls = tio.LabelMap if img["kind"] == "label" else tio.ScalarImage
        data = img["data"].clone()
        if device is not None:
            data = data.to(device)
        affines = [tio.AffineMatrix(a) for a in img["affines"]]
        batches[name] = tio.ImagesBatch(data, affines, image_class=cls)
    return tio.SubjectsBatch(batches)


def make_product_transform(spec):
    import warnings

    import torchio_b200 as tio

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        if isinstance(spec, list):
            return tio.Compose([make_product_transform(s) for s in spec], copy=False)
        name, kwargs = spec
        return getattr(tio, name)(**kwargs)


def blank_transform(name):
    """A transform instance whose apply_transform can replay recorded params."""
    import warnings

    import torchio_b200 as tio

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        required = {"Crop": {"cropping": 0}, "Pad": {"padding": 0}}.get(name, {})
        return getattr(tio, name)(**required)


def product_replay(batch, history):
    """Apply recorded params through the product's apply_transform (CUDA)."""
    for step in history:
        batch = blank_transform(step["name"]).apply_transform(batch, step["params"])
    return batch
