"""Input dtypes other than fp32 and autograd inputs (SURVEY §8 b, "dtype rules"): the kernels
compute in fp32 and the transforms hand back what the reference hands back — the input dtype after
Spatial / BiasField / Blur, torch's type promotion after Noise / Gamma — within the precision of
that dtype; tensors that require grad are refused (forward-only kernels), not silently detached."""

import copy
import warnings

import pytest
import torch

pytestmark = pytest.mark.gpu


def _pipeline(tio):
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return tio.Compose([
            tio.Affine(scales=(0.9, 1.1), degrees=(-10, 10)),
            tio.ElasticDeformation(max_displacement=3.0),
            tio.BiasField(), tio.Blur(std=(0, 2)), tio.Noise(std=(0, 0.25)),
            tio.Gamma(log_gamma=(-0.3, 0.3))], copy=False)


# relative precision of one rounding to the dtype, with head-room for the chain of six casts
@pytest.mark.parametrize("dtype,tol", [(torch.float16, 6e-3), (torch.bfloat16, 5e-2), (torch.float64, 1e-4)])
def test_half_and_double_inputs_follow_the_reference_dtype_rules(dtype, tol):
    import torchio_b200 as tio
    from oracle import torch_port

    g = torch.Generator().manual_seed(3)
    data = (torch.rand((2, 1, 24, 20, 18), generator=g) + 0.2).to(dtype)
    reference_input = {"t1": {"kind": "scalar", "data": data.clone(),
                              "affines": [tio.AffineMatrix().numpy().copy() for _ in range(2)]}}
    batch = tio.SubjectsBatch({"t1": tio.ImagesBatch(data.cuda(), [tio.AffineMatrix() for _ in range(2)])})
    torch.manual_seed(11)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        out = _pipeline(tio)(batch)
    history = [{"name": t.name, "params": t.params} for t in out.applied_transforms]
    expected = torch_port.replay(copy.deepcopy(reference_input), history)["t1"]["data"]
    got = out.images["t1"].data.cpu()
    assert got.dtype == expected.dtype, (got.dtype, expected.dtype)
    rng = float(expected.double().max() - expected.double().min())
    err = float((got.double() - expected.double()).abs().max()) / rng
    assert err <= tol, err


def test_spatial_alone_returns_the_input_dtype():
    import torchio_b200 as tio

    for dtype in (torch.float16, torch.bfloat16, torch.float64):
        data = torch.rand((2, 1, 16, 16, 16)).to(dtype)
        batch = tio.SubjectsBatch({"t1": tio.ImagesBatch(data.cuda(), [tio.AffineMatrix() for _ in range(2)])})
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            out = tio.Affine(degrees=(-10, 10), copy=False)(batch)
        assert out.images["t1"].data.dtype == dtype


def test_inputs_that_require_grad_are_refused():
    import torchio_b200 as tio

    data = torch.rand((1, 1, 16, 16, 16), device="cuda", requires_grad=True)
    batch = tio.SubjectsBatch({"t1": tio.ImagesBatch(data, [tio.AffineMatrix()])})
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        with pytest.raises(NotImplementedError, match="forward-only"):
            tio.Affine(degrees=(-10, 10), copy=False)(batch)
        with pytest.raises(NotImplementedError, match="forward-only"):
            tio.Gamma(log_gamma=(-0.3, 0.3), copy=False)(batch)
