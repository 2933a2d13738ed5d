"""GPU parity: the CUDA path (through the C-ABI) vs the reference's golden
outputs and vs the C oracle, on every golden case.

Bars: nearest-neighbour label maps bit-exact; fp32 images within 1e-4 of the
dynamic range of the reference output (north-star tolerance) — in practice the
resample kernel is bit-exact and the transcendental kernels are ~1e-7.
"""

import copy

import pytest
import torch

from golden_cases import CALL_CASES, CASES, NEIGHBOUR_CASES, RESAMPLE_CASES
from util import load_golden, product_batch, product_replay, report

pytestmark = pytest.mark.gpu

TOL_RANGE = 1e-4
SPATIAL = {"Affine", "ElasticDeformation", "Spatial", "Resample"}


# fast coordinates: the reference's own coordinate noise (<= ~2e-5 voxel) times the local gradient
TOL_FAST_VS_ORACLE = 1e-4


@pytest.mark.parametrize("name", [c["name"] for c in CASES + RESAMPLE_CASES])
def test_cuda_matches_reference_golden(name, coords):
    from oracle import c_port

    _, images, history, expected, expected_aff = load_golden(name)
    batch = product_batch(images, device="cuda")
    out = product_replay(batch, history)
    torch.cuda.synchronize()
    oracle = c_port.replay(copy.deepcopy(images), history)
    # (the anti-alias pre-filter is a blur: conv summation order, not a pure gather)
    only_spatial = all(h["name"] in SPATIAL and not h["params"].get("antialias") for h in history)
    for n, exp in expected.items():
        got = out.images[n].data.cpu()
        assert got.dtype == exp.dtype and got.shape == exp.shape
        r = report(got, exp)
        if images[n]["kind"] == "label":
            assert r["n_mismatch"] == 0, (n, r)
        else:
            assert r["max_abs_over_range"] <= TOL_RANGE, (n, r)
        ro = report(got, oracle[n]["data"])
        if only_spatial and images[n]["kind"] == "label":
            assert ro["n_mismatch"] == 0, (n, ro)  # same coordinates, same rounding
        elif only_spatial and coords == "exact":  # same coordinates; taps blended with FMA lerps (<= 1 ulp)
            assert ro["max_abs_over_range"] <= 3e-7, (n, ro)
        elif coords == "exact":
            assert ro["max_abs_over_range"] <= 2e-6, (n, ro)
        else:
            assert ro["max_abs_over_range"] <= TOL_FAST_VS_ORACLE and ro["frac_gt_1e-4_range"] == 0.0, (n, ro)
        assert r["frac_gt_1e-4_range"] == 0.0 or images[n]["kind"] == "label", (n, r)
        for b, a in enumerate(out.images[n].affines):
            assert abs(a.numpy() - expected_aff[n][b]).max() < 1e-12


@pytest.mark.parametrize("name", [c["name"] for c in NEIGHBOUR_CASES])
def test_cuda_neighbours_match_reference_golden(name):
    """Flip / Crop / Pad: pure index moves, bit-exact with the reference (the composed
    case contains an Affine: labels exact, images within the resample tolerance);
    affines follow the reference's origin shifts; the inverse restores the input."""
    _, images, history, expected, expected_aff = load_golden(name)
    batch = product_batch(images, device="cuda")
    out = product_replay(batch, history)
    exact = all(h["name"] in ("Flip", "Crop", "Pad") for h in history)
    # padding_mode="mean": the fill is a whole-volume fp32 mean (torch: fp32 cascade sum; here:
    # fp64 sums rounded once) -- equal to rounding, not to the bit
    mean_fill = any(h["params"].get("padding_mode") == "mean" for h in history)
    for n, exp in expected.items():
        got = out.images[n].data.cpu()
        assert got.dtype == exp.dtype and got.shape == exp.shape
        if mean_fill and images[n]["kind"] != "label":
            assert report(got, exp)["max_abs_over_range"] <= 1e-6, (n, report(got, exp))
        elif exact or images[n]["kind"] == "label":
            assert torch.equal(got, exp), (n, report(got, exp))
        else:
            assert report(got, exp)["max_abs_over_range"] <= TOL_RANGE
        for b, a in enumerate(out.images[n].affines):
            assert abs(a.numpy() - expected_aff[n][b]).max() < 1e-12


@pytest.mark.parametrize("name", [c["name"] for c in CALL_CASES])
def test_croporpad_public_call_matches_reference(name):
    """CropOrPad through the public call with the reference's seed: bit-exact data, the
    reference's affines, and the same three history records (Pad, Crop, CropOrPad)."""
    import json
    import warnings

    from golden_cases import CASES_BY_NAME
    from util import make_product_transform

    case = CASES_BY_NAME[name]
    _, images, history, expected, expected_aff = load_golden(name)
    batch = product_batch(images, device="cuda")
    torch.manual_seed(case["seed"])
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        out = make_product_transform(case["transform"])(batch)
    for n, exp in expected.items():
        got = out.images[n].data.cpu()
        assert got.dtype == exp.dtype and got.shape == exp.shape
        assert torch.equal(got, exp), (n, report(got, exp))
        for b, a in enumerate(out.images[n].affines):
            assert abs(a.numpy() - expected_aff[n][b]).max() < 1e-12
    mine = [{"name": t.name, "params": t.params} for t in out.applied_transforms]
    assert json.loads(json.dumps(mine)) == history


@pytest.mark.parametrize("name", ["flip_b4_per_instance", "crop_aniso", "pad_constant"])
def test_neighbour_inverse_round_trip(name):
    import warnings

    from golden_cases import CASES_BY_NAME
    from util import make_product_transform

    case, images, _, _, _ = load_golden(name)
    batch = product_batch(images, device="cuda")
    before = {n: ib.data.clone() for n, ib in batch.images.items()}
    torch.manual_seed(case["seed"])
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        out = make_product_transform(CASES_BY_NAME[name]["transform"])(batch)
        back = out.apply_inverse_transform()
    for n, ref in before.items():
        got = back.images[n].data
        if name == "crop_aniso":  # cropped voxels come back as padding zeros
            i0, i1, j0, j1, k0, k1 = case["transform"][1]["cropping"]
            sl = (..., slice(i0, ref.shape[-3] - i1), slice(j0, ref.shape[-2] - j1), slice(k0, ref.shape[-1] - k1))
            assert torch.equal(got[sl], ref[sl])
        else:
            assert torch.equal(got, ref)


@pytest.mark.parametrize("name", ["compose_full_b2", "affine_gated", "noise_rician_gated",
                                  "compose_flip_pad_affine_crop", "flip_gated_two_axes"])
def test_public_call_path_matches_golden(name):
    """Same check through Compose.__call__ with the reference's seed: sampling,
    gating and kernels together."""
    from golden_cases import CASES_BY_NAME
    from util import make_product_transform

    case = CASES_BY_NAME[name]
    _, images, history, expected, _ = load_golden(name)
    transform = make_product_transform(case["transform"])
    batch = product_batch(images, device="cuda")
    torch.manual_seed(case["seed"])
    import warnings

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        out = transform(batch)
    for n, exp in expected.items():
        got = out.images[n].data.cpu()
        r = report(got, exp)
        if images[n]["kind"] == "label":
            assert r["n_mismatch"] == 0, (n, r)
        else:
            assert r["max_abs_over_range"] <= TOL_RANGE, (n, r)


def test_cpu_resident_input_round_trips_through_the_gpu():
    """A CPU batch is staged to the GPU and comes back on the CPU."""
    from golden_cases import CASES_BY_NAME
    from util import make_product_transform

    case = CASES_BY_NAME["compose_config2_b2"]
    _, images, history, expected, _ = load_golden(case["name"])
    transform = make_product_transform(case["transform"])
    batch = product_batch(images)
    torch.manual_seed(case["seed"])
    import warnings

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        out = transform(batch)
    for n, exp in expected.items():
        got = out.images[n].data
        assert got.device.type == "cpu"
        r = report(got, exp)
        assert r["max_abs_over_range"] <= TOL_RANGE, (n, r)
