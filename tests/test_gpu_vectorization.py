"""Ports of the reference's equivalence tests for the hot path, on the GPU:

* `assert_vectorized` (reference tests/conftest.py:16-75, used by tests/test_vectorization.py:33-68):
  a per-instance transform applied to a batch equals the same per-element params applied to every
  element alone, and gated-out elements are bit-for-bit no-ops;
* "inverse restores geometry" (reference tests/test_spatial.py:295-312) through `_SpatialInverse`,
  plus target spaces: `Resample` by spacing, by image name, and back.
"""

import copy
import warnings

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _tio():
    import torchio_b200 as tio

    return tio


def _batch(batch_size=4, shape=(12, 12, 12), seg=False):
    tio = _tio()
    torch.manual_seed(99)
    data = torch.rand(1, *shape)
    subjects = []
    for index in range(batch_size):
        kwargs = {"t1": tio.ScalarImage((data.clone() + index).cuda())}
        if seg:
            kwargs["seg"] = tio.LabelMap((torch.rand(1, *shape) * 4).to(torch.int16).cuda())
        subjects.append(tio.Subject(**kwargs))
    return tio.SubjectsBatch.from_subjects(subjects)


def assert_vectorized(transform, batch, *, rtol=1e-5, atol=1e-6):
    from torchio_b200.params import slice_params

    tio = _tio()
    original = copy.deepcopy(batch)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        result = transform(batch)
    params = result.applied_transforms[-1].params
    assert "_batched_keys" in params, "per-instance path was not active"
    keep = params.get("_keep")
    names = list(transform._get_images(result))
    subjects = original.unbatch()
    for index in range(original.batch_size):
        single = tio.SubjectsBatch.from_subjects([subjects[index]])
        single_input = {n: im.data.clone() for n, im in transform._get_images(single).items()}
        single = transform.apply_transform(single, slice_params(params, index, index + 1))
        gated_out = keep is not None and not keep[index]
        for name in names:
            row = transform._get_images(result)[name].data[index:index + 1]
            torch.testing.assert_close(row, transform._get_images(single)[name].data, rtol=rtol, atol=atol)
            if gated_out:
                torch.testing.assert_close(row, single_input[name], rtol=0, atol=0)


def _make(name, **kw):
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return getattr(_tio(), name)(**kw)


@pytest.mark.parametrize("spec", [
    ("Blur", {"std": (0.5, 2.0)}), ("BiasField", {"std": (0.3, 0.8)}),
    ("Flip", {"axes": (0, 1, 2), "flip_probability": 0.5}), ("Gamma", {"log_gamma": (-0.3, 0.3)}),
    # (a numeric pad value: "minimum" is defined by sample 0 of whatever batch the transform sees)
    ("Affine", {"scales": (0.9, 1.1), "degrees": (-10, 10), "default_pad_value": 0.5}),
    ("ElasticDeformation", {"max_displacement": 2.0, "default_pad_value": 0.5}),
], ids=lambda s: s[0])
def test_vectorized_matches_per_element(spec):
    torch.manual_seed(0)
    assert_vectorized(_make(spec[0], **spec[1]), _batch(seg=spec[0] in ("Affine", "ElasticDeformation", "Flip")))


@pytest.mark.parametrize("spec", [
    ("Blur", {"std": 1.5, "p": 0.5}), ("BiasField", {"std": 0.5, "p": 0.5}),
    ("Flip", {"axes": (0, 1, 2), "flip_probability": 1.0, "p": 0.5}), ("Gamma", {"log_gamma": 0.3, "p": 0.5}),
    ("Affine", {"degrees": (-10, 10), "p": 0.5, "default_pad_value": 0.5}),
], ids=lambda s: s[0])
def test_vectorized_matches_per_element_with_gating(spec):
    torch.manual_seed(0)
    assert_vectorized(_make(spec[0], **spec[1]), _batch(batch_size=6))


def test_inverse_restores_geometry_and_excluded_images():
    tio = _tio()
    g = torch.Generator().manual_seed(1)
    t1 = torch.rand((1, 20, 24, 16), generator=g)
    seg = (torch.rand((1, 20, 24, 16), generator=g) * 4).to(torch.int16)
    affine = np.diag([1.2, 0.9, 1.5, 1.0])
    affine[:3, 3] = [5.0, -3.0, 2.0]
    subject = tio.Subject(t1=tio.ScalarImage(t1.cuda(), affine=affine.copy()),
                          seg=tio.LabelMap(seg.cuda(), affine=affine.copy()))
    transform = _make("Affine", scales=(1.1, 0.9, 1.0), degrees=(0.0, 0.0, 20.0), translation=(1.0, -2.0, 0.5),
                      center="image", default_pad_value=0.0, default_pad_label=0.0)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        moved = transform(subject)
        restored = moved.apply_inverse_transform()
    assert not torch.allclose(moved.t1.data, subject.t1.data)
    np.testing.assert_allclose(moved.t1.affine.numpy(), subject.t1.affine.numpy())
    assert tuple(restored.t1.spatial_shape) == tuple(subject.t1.spatial_shape)
    np.testing.assert_allclose(restored.t1.affine.numpy(), subject.t1.affine.numpy())
    # the round trip is two interpolations of white noise: only the smooth part comes back, but the
    # geometry does — a smooth image returns to itself up to interpolation error
    i, j, k = torch.meshgrid(torch.arange(20.), torch.arange(24.), torch.arange(16.), indexing="ij")
    smooth = (torch.sin(i / 6) + torch.cos(j / 7) + k / 16)[None]
    s2 = tio.Subject(t1=tio.ScalarImage(smooth.cuda(), affine=affine.copy()))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        back = transform(s2).apply_inverse_transform()
    inner = (slice(None), slice(5, 15), slice(6, 18), slice(4, 12))
    assert float((back.t1.data.cpu()[inner] - smooth[inner]).abs().max()) < 2e-2
    # excluded images stay bit-for-bit identical through forward + inverse
    only_t1 = _make("Affine", degrees=(0.0, 0.0, 20.0), include=["t1"])
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        round_trip = only_t1(subject).apply_inverse_transform()
    assert torch.equal(round_trip.seg.data, subject.seg.data)


def test_resample_targets_spacing_name_and_back():
    """`Resample(spacing)` moves every image onto the new grid (shape, affine, physical centre kept);
    `Resample("name")` onto another image's grid; the inverse returns to the original grid."""
    tio = _tio()
    g = torch.Generator().manual_seed(2)
    i, j, k = torch.meshgrid(torch.arange(24.), torch.arange(20.), torch.arange(16.), indexing="ij")
    smooth = (torch.sin(i / 5) + torch.cos(j / 6) + k / 10)[None]
    affine = np.diag([1.0, 1.0, 2.0, 1.0])
    subject = tio.Subject(t1=tio.ScalarImage(smooth.cuda(), affine=affine.copy()),
                          seg=tio.LabelMap((torch.rand((1, 24, 20, 16), generator=g) * 3).to(torch.uint8).cuda(),
                                           affine=affine.copy()))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        down = tio.Resample(2, antialias=True)(subject)
    assert tuple(down.t1.spatial_shape) == (12, 10, 16) and tuple(down.seg.spatial_shape) == (12, 10, 16)
    assert np.allclose(down.t1.affine.spacing, (2.0, 2.0, 2.0))
    old_centre = affine[:3, 3] + affine[:3, :3] @ ((np.array([24, 20, 16]) - 1) / 2)
    a = down.t1.affine.numpy()
    assert np.allclose(a[:3, 3] + a[:3, :3] @ ((np.array([12, 10, 16]) - 1) / 2), old_centre)
    assert down.seg.data.dtype == torch.uint8 and set(torch.unique(down.seg.data).tolist()) <= {0, 1, 2}
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        back = down.apply_inverse_transform()
    assert tuple(back.t1.spatial_shape) == (24, 20, 16)
    np.testing.assert_allclose(back.t1.affine.numpy(), affine)
    inner = (slice(None), slice(4, 20), slice(4, 16), slice(2, 14))
    assert float((back.t1.data.cpu()[inner] - smooth[inner]).abs().max()) < 0.15
    # by name: t1 onto the grid of a coarser image of the same subject
    coarse = tio.ScalarImage(torch.zeros(1, 8, 10, 4).cuda(), affine=np.diag([3.0, 2.0, 8.0, 1.0]))
    both = tio.Subject(t1=subject.t1, ref=coarse)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        named = tio.Resample("ref", include=["t1"])(both)
        explicit = tio.Resample(((8, 10, 4), np.diag([3.0, 2.0, 8.0, 1.0])), include=["t1"])(both)
    assert tuple(named.t1.spatial_shape) == (8, 10, 4)
    assert torch.equal(named.t1.data, explicit.t1.data)
    with pytest.raises(ValueError):
        tio.Resample("missing")(both)
    with pytest.raises(ValueError):
        tio.Resample((1.0, -1.0, 1.0))(both)
