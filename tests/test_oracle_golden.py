"""The torch-port oracle must reproduce the reference's golden outputs.

CPU-only.  On the torch build that generated the fixtures the match is
bit-exact; on another CPU/torch build ATen kernels may differ in the last ulp,
so the assertion is 2e-6 of range for images and exact for labels.
"""

import copy

import pytest
import torch

from golden_cases import CASES, NEIGHBOUR_CASES, RESAMPLE_CASES, STAT_CASES
from oracle import torch_port
from util import load_golden, report


@pytest.mark.parametrize("name", [c["name"] for c in CASES + NEIGHBOUR_CASES + STAT_CASES + RESAMPLE_CASES])
def test_torch_port_matches_reference_golden(name):
    _, images, history, expected, expected_aff = load_golden(name)
    out = torch_port.replay(copy.deepcopy(images), history)
    for n, exp in expected.items():
        got = out[n]["data"]
        assert got.dtype == exp.dtype
        assert got.shape == exp.shape
        if images[n]["kind"] == "label":
            assert torch.equal(got, exp), report(got, exp)
        else:
            r = report(got, exp)
            assert r["max_abs_over_range"] <= 2e-6, r
        for b, a in enumerate(out[n]["affines"]):
            assert abs(a - expected_aff[n][b]).max() < 1e-12
