"""Full-size (256^3, BASELINE.json's volume size) checks through size-independent
properties — the oracle is too slow at this size, so correctness is pinned by
invariants of the domain plus spot comparisons of sub-blocks against the oracle."""

import json
import warnings

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

S = 256


def _tio():
    import torchio_b200 as tio

    return tio


def _quiet(fn, *a, **k):
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return fn(*a, **k)


def _volumes(b, seed=0):
    g = torch.Generator().manual_seed(seed)
    return torch.rand((b, 1, S, S, S), generator=g)


def _labels(b):
    i = torch.arange(S)
    lab = ((i[:, None, None] // 37 + i[None, :, None] // 53 + i[None, None, :] // 29) % 5).to(torch.int16)
    return lab[None, None].repeat(b, 1, 1, 1, 1).contiguous()


def _batch(images, labels=None):
    tio = _tio()
    b = images.shape[0]
    d = {"t1": tio.ImagesBatch(images.cuda(), [tio.AffineMatrix() for _ in range(b)])}
    if labels is not None:
        d["seg"] = tio.ImagesBatch(labels.cuda(), [tio.AffineMatrix() for _ in range(b)],
                                   image_class=tio.LabelMap)
    return tio.SubjectsBatch(d)


def test_full_pipeline_gated_rows_are_bit_exact_and_labels_stay_in_set():
    tio = _tio()
    x, lab = _volumes(4), _labels(4)
    pipe = _quiet(lambda: tio.Compose([
        tio.Affine(scales=(0.9, 1.1), degrees=(-10, 10), p=0.5),
        tio.ElasticDeformation(p=0.5), tio.BiasField(p=0.5), tio.Blur(std=(0, 2), p=0.5),
        tio.Noise(std=(0, 0.25), p=0.5), tio.Gamma(log_gamma=(-0.3, 0.3), p=0.5)]))
    torch.manual_seed(21)
    out = _quiet(pipe, _batch(x, lab))
    keeps = {t.name: t.params.get("_keep") for t in out.applied_transforms}
    assert any(k is not None and not all(k) for k in keeps.values())
    y = out.images["t1"].data.cpu()
    seg = out.images["seg"].data.cpu()
    for b in range(4):
        untouched = all(k is None or not k[b] for k in keeps.values()) and all(
            k is not None for k in keeps.values())
        if untouched:
            assert torch.equal(y[b], x[b])
        spatial_off = all(keeps[n] is not None and not keeps[n][b] for n in ("Affine", "ElasticDeformation")
                          if n in keeps)
        if spatial_off:
            assert torch.equal(seg[b], lab[b])
    assert set(torch.unique(seg).tolist()) <= {0, 1, 2, 3, 4}
    assert torch.isfinite(y).all()


def test_trilinear_resample_is_linear_and_tile_path_equals_general_path(coords):
    from torchio_b200 import ops

    rng = np.random.default_rng(5)
    x, y = _volumes(1, 1).cuda(), _volumes(1, 2).cuda()
    ang = rng.uniform(-0.17, 0.17, 3)
    cx, sx, cy, sy, cz, sz = np.cos(ang[0]), np.sin(ang[0]), np.cos(ang[1]), np.sin(ang[1]), np.cos(ang[2]), np.sin(ang[2])
    r = (np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]]) @ np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
         @ np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]])) * 1.07
    c = np.full(3, (S - 1) / 2)
    m = np.eye(4); m[:3, :3] = r; m[:3, 3] = c - r @ c + np.array([1.5, -2.25, 0.75])
    mat = torch.tensor(m.astype(np.float32)[:3].reshape(1, 12)).cuda()
    cp = torch.tensor(rng.uniform(-6, 6, (1, 7, 7, 7, 3)).astype(np.float32)).cuda()
    flags = torch.tensor([2], dtype=torch.uint8).cuda()
    kw = dict(affine_first=True, mode=ops.LINEAR, fill=None)
    one = (1.0, 1.0, 1.0)
    rx = ops.resample(x, mat, cp, flags, one, one, **kw)
    ry = ops.resample(y, mat, cp, flags, one, one, **kw)
    rxy = ops.resample(0.3 * x - 1.7 * y, mat, cp, flags, one, one, **kw)
    assert float((rxy - (0.3 * rx - 1.7 * ry)).abs().max()) <= 5e-6
    general = ops.resample(x, mat, cp, flags, one, one, box_hint=-1, **kw)
    tol = 1e-6 if coords == "exact" else 1e-4  # fast: the reference's coordinate noise x white-noise gradients (measured 6.2e-5)
    assert float((rx - general).abs().max()) <= tol
    fill = torch.tensor([-3.0]).cuda()
    f_fast = ops.resample(x, mat, cp, flags, one, one, affine_first=True, mode=ops.LINEAR, fill=fill)
    f_gen = ops.resample(x, mat, cp, flags, one, one, affine_first=True, mode=ops.LINEAR, fill=fill, box_hint=-1)
    assert torch.equal(f_fast == -3.0, f_gen == -3.0)  # identical fill decisions
    assert float((f_fast - f_gen).abs().max()) <= tol


def test_resample_block_matches_c_oracle_at_full_size():
    """A 48x40x64 output block of a 256^3 resample vs the C oracle run on the
    sub-volume that contains its pre-image (coordinates shifted accordingly)."""
    import ctypes

    from oracle import c_port
    from torchio_b200 import ops

    rng = np.random.default_rng(8)
    x = _volumes(1, 3)
    # near-identity affine so the pre-image of the block stays inside a known crop
    m = np.eye(4); m[:3, :3] += rng.uniform(-0.02, 0.02, (3, 3)); m[:3, 3] = rng.uniform(-1, 1, 3)
    mat = m.astype(np.float32)[:3].reshape(1, 12)
    mat_t = torch.tensor(mat)  # kept alive: the oracle receives its raw pointer
    lab = _labels(1)
    for data, mode in ((x, 1), (lab, 0)):
        got = ops.resample(data.cuda(), mat_t.cuda(), None, None, (1, 1, 1), (1, 1, 1),
                           affine_first=True, mode=mode, fill=None, box_hint=-1).cpu()
        want = torch.empty_like(data)
        p = c_port._p
        sp = torch.ones(3)
        c_port.lib().orc_resample(p(data), p(want), c_port._DTYPES[data.dtype], 1, 1, S, S, S, S, S, S,
                                  p(mat_t), None, None, 0, 0, 0, p(sp), p(sp), 1, mode, None)
        assert torch.equal(got, want)


def test_intensity_identities_inverses_and_seed_replay():
    tio = _tio()
    x = _volumes(2, 4) - 0.25
    batch = _batch(x)
    # zero-parameter identities are exact
    for t in (_quiet(lambda: tio.BiasField(std=0.0)), _quiet(lambda: tio.Blur(std=0.0)),
              _quiet(lambda: tio.Gamma(log_gamma=0.0)), _quiet(lambda: tio.Noise(std=0.0))):
        out = _quiet(t, _batch(x))
        assert torch.equal(out.images["t1"].data.cpu(), x), type(t).__name__
    # blur keeps constants (taps sum to one) and bias/gamma invert
    const = torch.full((1, 1, S, S, S), 0.625)
    out = _quiet(_quiet(lambda: tio.Blur(std=(0.5, 2.0))), _batch(const))
    assert float((out.images["t1"].data.cpu() - 0.625).abs().max()) <= 2e-6
    for make, atol in ((lambda: tio.Gamma(log_gamma=(-0.3, 0.3)), 2e-4), (lambda: tio.BiasField(), 2e-5)):
        t = _quiet(make)
        torch.manual_seed(3)
        fwd = _quiet(t, _batch(x))
        back = _quiet(fwd.apply_inverse_transform)
        assert float((back.images["t1"].data.cpu() - x).abs().max()) <= atol
    # same params + seed => identical output (reference tests/test_noise.py:109-131)
    noise = _quiet(lambda: tio.Noise(std=(0.05, 0.25)))
    torch.manual_seed(9)
    a = _quiet(noise, _batch(x))
    params = a.applied_transforms[0].params
    b = noise.apply_transform(_batch(x), json.loads(json.dumps(params)))
    assert torch.equal(a.images["t1"].data, b.images["t1"].data)
    del batch


def test_device_mt_stream_window_at_full_batch_positions():
    """Normals for the last volume of a 32 x 256^3 batch (stream offset 31 * 2^24)
    equal torch's CPU stream there."""
    from torchio_b200 import ops

    seed, offset, n = 424242, 31 * 2**24, 2**16
    z = ops.randn_mt19937(seed, offset, n, "cuda").cpu()
    g = torch.Generator().manual_seed(seed)
    torch.randn(offset, generator=g)
    want = torch.randn(n, generator=g)
    assert float((z - want).abs().max()) <= 4e-6


def test_fused_equals_sequential_at_full_size():
    tio = _tio()
    x = _volumes(2, 6)
    outs = []
    for fuse in (True, False):
        pipe = _quiet(lambda: tio.Compose([tio.BiasField(), tio.Blur(std=(0, 2)), tio.Noise(std=(0, 0.25)),
                                           tio.Gamma(log_gamma=(-0.3, 0.3))]))
        pipe.fuse = fuse
        torch.manual_seed(13)
        outs.append(_quiet(pipe, _batch(x)).images["t1"].data)
    rng = float(outs[1].max() - outs[1].min())
    assert float((outs[0] - outs[1]).abs().max()) <= 3e-6 * rng
