"""Host logic: with the same torch.manual_seed the product samples exactly the
params the reference recorded (RNG call order, schema, values).  CPU-only."""

import json
import warnings

import pytest
import torch

from golden_cases import CASES
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
@pytest.mark.parametrize("name", [c["name"] for c in CASES])
def test_params_match_reference(name, fuse, monkeypatch):
    """Sequential and Compose-fused execution draw the same RNG stream."""
    case, images, history, _, _ = load_golden(name)
    batch = product_batch(images)
    transform = make_product_transform(case["transform"])
    torch.manual_seed(case["seed"])
    mine = _sample_only(transform, batch, monkeypatch, fuse)
    # JSON round trip == what the reference stores in history
    assert json.loads(json.dumps(mine)) == history
