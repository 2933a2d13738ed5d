"""Regression tests for the round-1 review findings (host logic on CPU, streamed path on the GPU)."""

import copy
import json
import pickle
import warnings

import numpy as np
import pytest
import torch


def test_lazy_params_are_forced_by_every_copy_path():
    """`dict(params)`, `{**params}`, `|`, `update`, `pop`, `setdefault` must see the materialised
    matrices, not the placeholder of a lazy entry (an inverse built from such a copy was an identity)."""
    from torchio_b200.params import LazyParams

    def make():
        p = LazyParams(a=1)
        p.set_lazy("affine_matrix", lambda: [[1.0, 2.0]])
        return p

    want = {"a": 1, "affine_matrix": [[1.0, 2.0]]}
    assert dict(make()) == want
    assert {**make()} == want
    assert ({"z": 0} | make()) == {"z": 0, **want} and (make() | {"z": 0}) == {**want, "z": 0}
    d = {}
    d.update(make())
    assert d == want
    assert make().pop("affine_matrix") == [[1.0, 2.0]]
    assert make().setdefault("affine_matrix", None) == [[1.0, 2.0]]
    assert copy.deepcopy(make()) == want and pickle.loads(pickle.dumps(dict(make()))) == want
    assert json.loads(json.dumps(make())) == want


def test_spatial_history_copy_still_inverts():
    import torchio_b200 as tio

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        t = tio.Affine(degrees=(-10, 10))
    batch = tio.SubjectsBatch({"t1": tio.ImagesBatch(torch.rand(2, 1, 8, 8, 8), [tio.AffineMatrix()] * 2)})
    torch.manual_seed(3)
    params = t.make_params(batch)
    copied = dict(params)
    assert copied["affine_matrix"] is not None and all(m is not None for m in copied["affine_matrix"])


def test_inverse_plan_reports_skipped_records():
    import torchio_b200 as tio
    from torchio_b200.transforms.inverse import plan_inverse

    history = [tio.AppliedTransform("Gamma", {"log_gamma": 0.2}), tio.AppliedTransform("Noise", {"seed": 1}),
               tio.AppliedTransform("NoSuchTransform", {}), tio.AppliedTransform("Flip", {"axes": [0]})]
    plan = plan_inverse(history)
    assert [type(s).__name__ for s in plan.steps][0] == "Flip" and len(plan.steps) == 2
    assert ("Noise", "one-way") in plan.skipped and ("NoSuchTransform", "unknown") in plan.skipped
    with pytest.warns(UserWarning):
        tio.get_inverse_transform(history)
    assert len(tio.get_inverse_transform(history, warn=False, ignore_intensity=True).transforms) == 1


@pytest.mark.gpu
def test_streamed_blur_uses_whole_batch_tables_and_compose_stays_picklable():
    """Heterogeneous spacings + a shared-sigma Blur + chunk of one element: the slices must use batch
    element 0's spacing and the whole batch's shared/stacked decision, i.e. equal the one-shot rows.
    The Compose must survive pickle / deepcopy after a streamed call (no streams in its __dict__)."""
    import torchio_b200 as tio

    g = torch.Generator().manual_seed(2)
    x = torch.rand((3, 1, 24, 32, 32), generator=g)
    spacings = [(1.0, 1.0, 1.0), (0.5, 2.0, 1.0), (2.0, 0.7, 1.5)]

    def batch():
        return tio.SubjectsBatch({"t1": tio.ImagesBatch(
            x.clone().pin_memory(), [tio.AffineMatrix(np.diag([*s, 1.0])) for s in spacings])})

    for per_instance in (False, True):
        outs = []
        for chunk in (0, 1):
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                pipe = tio.Compose([tio.Blur(std=(0.5, 2.0), per_instance=per_instance),
                                    tio.Gamma(log_gamma=(-0.2, 0.2))], copy=False)
            pipe.chunk_size = chunk
            torch.manual_seed(8)
            outs.append(pipe(batch()).images["t1"].data)
            pickle.loads(pickle.dumps(pipe))
            copy.deepcopy(pipe)
        assert float((outs[0] - outs[1]).abs().max()) <= 1e-6, per_instance
    # a deep copy of a pinned batch stays pinned (asynchronous staging keeps working with copy=True)
    assert copy.deepcopy(batch()).images["t1"].data.is_pinned()
