"""Host logic: the product's kernel tables equal the oracle's independently
built ones for every golden case.  CPU-only."""

import numpy as np
import pytest
import torch

from golden_cases import CASES
from oracle import c_port, torch_port
from torchio_b200 import tables
from torchio_b200.transforms.spatial import _unpack_geometry
from util import load_golden


@pytest.mark.parametrize("name", [c["name"] for c in CASES])
def test_tables_match_oracle(name):
    _, images, history, _, _ = load_golden(name)
    first = next(iter(images.values()))
    b = first["data"].shape[0]
    shape = tuple(first["data"].shape[-3:])
    a0 = first["affines"][0]
    for step in history:
        p = step["params"]
        if step["name"] in ("Affine", "ElasticDeformation", "Spatial"):
            want = c_port.spatial_tables(p, b, shape, a0)
            mats, cps, per_instance = _unpack_geometry(p)
            got = tables.spatial_tables(
                mats if per_instance else mats[0], cps if per_instance else cps[0], b,
                a0, a0, per_instance=per_instance, has_target=False,
            )
            if want is None:
                assert got is None
                continue
            mat, cp, flags, _ = want
            assert np.array_equal(got.mat, mat.numpy())
            assert np.array_equal(got.flags, flags.numpy())
            assert (got.cp is None) == (cp is None)
            if cp is not None:
                assert np.array_equal(got.cp, cp.numpy())
        elif step["name"] == "Blur":
            per_instance = "_batched_keys" in p
            if per_instance:
                mm = np.asarray(p["std"], dtype=np.float64)
                sp = np.asarray([torch_port.spacing_of(a) for a in first["affines"]])
                vox = np.divide(mm, sp, out=np.zeros_like(mm), where=sp > 0)
            else:
                sp = np.asarray(torch_port.spacing_of(a0))
                vox = [s / q for s, q in zip(p["std"], sp)]
            want = c_port.blur_tables(vox, b)
            got = tables.blur_tables(vox, b)
            assert (want is None) == (got is None)
            if want is not None:
                taps, radius, big_r, identity = want
                assert torch.equal(got.taps, taps) and torch.equal(got.radius, radius)
                assert got.big_r == big_r and torch.equal(got.identity, identity)
                assert got.axes_mask == sum(
                    1 << a for a in range(3) if int(radius[a].max()) > 0
                )
        elif step["name"] == "BiasField":
            data_shape = first["data"].shape
            want = torch_port.coarse_bias_fields(data_shape, p["std"], p["seed"], p["scale"])
            got = tables.coarse_bias_fields(data_shape, p["std"], p["seed"], p["scale"])
            assert torch.equal(want, got)
