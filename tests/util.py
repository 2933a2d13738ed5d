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
