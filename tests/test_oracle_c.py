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


def test_c_remap_equals_torch_pad_flip_crop_on_random_cases():
    """orc_remap against F.pad / torch.flip / slicing for random offsets, modes, dtypes."""
    import numpy as np

    rng = np.random.default_rng(4)
    g = torch.Generator().manual_seed(4)
    for trial in range(40):
        dtype = [torch.float32, torch.int16, torch.uint8, torch.float64][trial % 4]
        shape = tuple(int(v) for v in rng.integers(3, 9, 3))
        x = (torch.rand((2, 2, *shape), generator=g) * 90).to(dtype)
        mode = ["constant", "replicate", "reflect", "circular"][int(rng.integers(0, 4))]
        limit = [min(s - 1, 3) if mode in ("reflect", "circular") else 3 for s in shape]
        pad = [int(rng.integers(0, limit[a] + 1)) for a in range(3) for _ in range(2)]
        if mode != "constant" and dtype not in (torch.float32, torch.float64):
            continue  # F.pad's non-constant modes are float-only on the CPU; covered for floats
        kw = {"value": 7} if mode == "constant" else {}
        want = torch.nn.functional.pad(x, (pad[4], pad[5], pad[2], pad[3], pad[0], pad[1]), mode=mode, **kw)
        out_shape = tuple(shape[a] + pad[2 * a] + pad[2 * a + 1] for a in range(3))
        got = c_port.remap(x, out_shape, (pad[0], pad[2], pad[4]), mode=mode, fill=7)
        assert torch.equal(got, want), (trial, mode, pad)
        # crop + per-element flips in one remap: the reversal acts on the source volume,
        # i.e. out = flip(source)[window]
        bits = [int(rng.integers(0, 8)) for _ in range(2)]
        back = c_port.remap(got, shape, (-pad[0], -pad[2], -pad[4]), flip_bits=bits)
        for b in range(2):
            dims = [a - 3 for a in range(3) if bits[b] >> a & 1]
            flipped = torch.flip(got[b], dims) if dims else got[b]
            ref = flipped[:, pad[0]:pad[0] + shape[0], pad[2]:pad[2] + shape[1], pad[4]:pad[4] + shape[2]]
            assert torch.equal(back[b], ref), (trial, bits)
