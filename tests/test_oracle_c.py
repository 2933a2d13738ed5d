"""The plain-C restatement must reproduce the reference's golden outputs.

CPU-only.  Discrete outputs (nearest-neighbour label maps, fill decisions)
must be bit-exact; fp32 images within 2e-6 of the dynamic range (libm vs the
vectorised ATen exp/pow/log/sin/cos differ by an ulp; conv3d summation order
is unspecified).
"""

import copy

import pytest
import torch

from golden_cases import CASES, NEIGHBOUR_CASES
from oracle import c_port
from util import load_golden, report

TOL = {"Noise": 5e-6}


@pytest.mark.parametrize("name", [c["name"] for c in CASES + NEIGHBOUR_CASES])
def test_c_oracle_matches_reference_golden(name):
    _, images, history, expected, _ = load_golden(name)
    out = c_port.replay(copy.deepcopy(images), history)
    tol = max([TOL.get(h["name"], 2e-6) for h in history] + [2e-6])
    for n, exp in expected.items():
        got = out[n]["data"]
        assert got.dtype == exp.dtype and got.shape == exp.shape
        r = report(got, exp)
        if images[n]["kind"] == "label":
            assert r["n_mismatch"] == 0, r
        else:
            assert r["max_abs_over_range"] <= tol, r


def test_c_mt19937_randn_matches_torch():
    for seed, n in ((0, 16), (1234, 4096), (2**31 - 1, 16 * 1000 + 5), (7, 37)):
        g = torch.Generator().manual_seed(seed)
        ref = torch.randn(n, generator=g)
        z, used = c_port.randn_mt19937(seed, 0, n)
        assert used == n + (16 if n % 16 else 0)
        assert (z - ref).abs().max() <= 4e-6
        # stream continuation: a second draw from the same generator
        ref2 = torch.randn(64, generator=g)
        z2, _ = c_port.randn_mt19937(seed, used, 64)
        assert (z2 - ref2).abs().max() <= 4e-6


def test_c_crop_patches_equals_slicing():
    g = torch.Generator().manual_seed(2)
    for dtype in (torch.uint8, torch.int16, torch.float32, torch.int64):
        vol = (torch.rand((2, 9, 10, 11), generator=g) * 50).to(dtype)
        corners = [[0, 0, 0], [2, 3, 4], [5, 4, 3]]
        got = c_port.crop_patches(vol, corners, (4, 6, 7))
        for row, (i, j, k) in enumerate(corners):
            assert torch.equal(got[row], vol[:, i:i + 4, j:j + 6, k:k + 7])
