"""Host logic: with the same torch.manual_seed the product samples exactly the
params the reference recorded (RNG call order, schema, values).  CPU-only."""

import json
import warnings

import pytest
import torch

from golden_cases import CALL_CASES, CASES, NEIGHBOUR_CASES, RESAMPLE_CASES
from util import load_golden, make_product_transform, product_batch


def _sample_only(transform, batch, monkeypatch, fuse):
    """Run the gate + make_params + history path with the kernels stubbed out."""
    import torchio_b200 as tio
    from torchio_b200.transforms import intensity

    leaves = transform.transforms if isinstance(transform, tio.Compose) else [transform]
    for leaf in leaves:
        leaf.apply_transform = lambda b, p: b
    monkeypatch.setattr(intensity, "run_stages", lambda images, builders: None)
    transform.fuse = fuse
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        out = transform._forward_batch(batch)
    return [{"name": t.name, "params": t.params} for t in out.applied_transforms]


@pytest.mark.parametrize("fuse", [False, True])
# (kernels are stubbed here, so a pipeline whose later params depend on an earlier shape
# change is checked through the real call on the GPU instead: test_gpu_golden.py)
@pytest.mark.parametrize("name", [c["name"] for c in CASES + NEIGHBOUR_CASES + RESAMPLE_CASES
                                  if c["name"] != "compose_flip_pad_affine_crop"])
def test_params_match_reference(name, fuse, monkeypatch):
    """Sequential and Compose-fused execution draw the same RNG stream."""
    case, images, history, _, _ = load_golden(name)
    batch = product_batch(images)
    transform = make_product_transform(case["transform"])
    torch.manual_seed(case["seed"])
    mine = _sample_only(transform, batch, monkeypatch, fuse)
    # JSON round trip == what the reference stores in history
    assert json.loads(json.dumps(mine)) == history


SPECS = [
    ("Affine", {"scales": (0.9, 1.1), "degrees": (-10, 10), "translation": (-3, 3)}),
    ("Affine", {"scales": (0.8, 1.2), "isotropic": True, "degrees": (0, 0, -20, 20, 0, 0), "p": 0.6}),
    ("ElasticDeformation", {}),
    ("ElasticDeformation", {"max_displacement": (1.0, 6.0), "num_control_points": (5, 6, 7),
                            "locked_borders": 1, "p": 0.5}),
    ("Spatial", {"scales": (0.9, 1.1, 1.0, 1.0, 0.95, 1.05), "degrees": 5.0,
                 "max_displacement": (0.0, 4.0, 0.0, 0.0, 2.0, 2.0), "affine_first": False}),
]


@pytest.mark.parametrize("spec", SPECS, ids=[f"{n}-{i}" for i, (n, _) in enumerate(SPECS)])
def test_vectorised_sampler_equals_per_element_loop(spec):
    """The one-draw fast sampler consumes the RNG stream exactly like the
    reference-shaped per-element loop (values, order, and what is left over)."""
    import torchio_b200 as tio

    name, kwargs = spec
    images = {"t1": {"kind": "scalar", "data": torch.zeros(6, 1, 12, 10, 8),
                     "affines": [__import__("numpy").eye(4)] * 6}}
    results = []
    for force_slow in (False, True):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            t = getattr(tio, name)(**kwargs)
        if force_slow:
            t._draw_plan = lambda: None
        torch.manual_seed(77)
        params = t.make_params(product_batch(images))
        tail = torch.rand(3).tolist()  # the stream must be left in the same place
        results.append((json.dumps(params, sort_keys=True), tail))
    assert results[0] == results[1]


def test_lazy_params_materialise_on_every_read_path():
    import copy
    import pickle

    from torchio_b200.params import LazyParams

    calls = []

    def thunk():
        calls.append(1)
        return [[1.0, 2.0], None]

    p = LazyParams({"a": 1})
    p.set_lazy("m", thunk)
    assert "m" in p and not calls and p["a"] == 1 and not calls
    assert p["m"] == [[1.0, 2.0], None] and len(calls) == 1
    for make in (lambda q: dict(q.items()), lambda q: json.loads(json.dumps(q)),
                 lambda q: pickle.loads(pickle.dumps(q)), copy.deepcopy, lambda q: q.copy()):
        q = LazyParams({"a": 1})
        q.set_lazy("m", thunk)
        assert make(q) == {"a": 1, "m": [[1.0, 2.0], None]}
    q = LazyParams({"a": 1})
    q.set_lazy("m", thunk)
    assert q == {"a": 1, "m": [[1.0, 2.0], None]}


def test_slice_params_keeps_shared_entries_and_slices_per_instance_ones():
    """params.slice_params: the per-element rule of data/batch.py:365-399 over a range."""
    import torch

    import torchio_b200 as tio
    from torchio_b200.params import LazyParams, slice_params

    shared = {"std": 0.5, "seed": 7}
    assert slice_params(shared, 1, 3) is shared
    p = LazyParams({"std": [0.1, 0.2, 0.3, 0.4], "seed": 11, "_batch_size": 4,
                    "_batched_keys": ["std", "big"], "_keep": [True, False, True, True]})
    forced = []
    p.set_lazy("big", lambda: forced.append(1) or [[0], [1], [2], [3]])
    q = slice_params(p, 1, 3)
    assert not forced  # still lazy
    assert q["std"] == [0.2, 0.3] and q["seed"] == 11 and q["_batch_size"] == 2
    assert q["_keep"] == [False, True] and q["big"] == [[1], [2]] and forced
    # spatial params carry their numpy geometry along
    torch.manual_seed(3)
    x = torch.zeros((4, 1, 8, 8, 8))
    batch = tio.SubjectsBatch({"t1": tio.ImagesBatch(x, [tio.AffineMatrix() for _ in range(4)])})
    t = tio.Affine(scales=(0.9, 1.1), degrees=(-10, 10))
    params = t.make_params(batch)
    part = slice_params(params, 2, 4)
    assert len(part._packed[0]) == 2 and part._packed[2] is True
    assert part["affine_matrix"] == params["affine_matrix"][2:4]


@pytest.mark.parametrize("name", [c["name"] for c in CALL_CASES])
def test_croporpad_params_match_reference(name):
    """CropOrPad.make_params (shape arithmetic, units, the torch.randint draws of a random
    crop) equals the reference's recorded params; the derived Pad / Crop records are
    checked with the data on the GPU (tests/test_gpu_golden.py)."""
    case, images, history, _, _ = load_golden(name)
    batch = product_batch(images)
    transform = make_product_transform(case["transform"])
    torch.manual_seed(case["seed"])
    assert torch.rand(1).item() < 2  # the gate draw of Transform._forward_batch
    params = transform.make_params(batch)
    assert json.loads(json.dumps(params)) == history[-1]["params"]
    assert history[-1]["name"] == "CropOrPad"


@pytest.mark.parametrize("name", ["compose_config2_b2", "compose_full_b2", "compose_full_b1_48"])
def test_streaming_plan_samples_like_sequential_application(name):
    """Compose._plan (used when a host batch is streamed in slices) draws gates and params
    for the whole batch up front: same RNG order, same params as the reference's history."""
    case, images, history, _, _ = load_golden(name)
    batch = product_batch(images)
    transform = make_product_transform(case["transform"])
    torch.manual_seed(case["seed"])
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        plan = transform._plan(batch)
    mine = [{"name": type(t).__name__, "params": p} for _, applied in plan for t, p in applied]
    assert json.loads(json.dumps(mine)) == history
    # and slicing the planned params row-wise is what unbatching the history would give
    from torchio_b200.params import slice_params

    b = batch.batch_size
    for _, applied in plan:
        for _, params in applied:
            part = slice_params(params, 0, 1)
            if "_batched_keys" in params:
                assert part["_batch_size"] == 1
                for key in params["_batched_keys"]:
                    assert part[key] == params[key][0:1]
            else:
                assert part is params
    assert b >= 1
