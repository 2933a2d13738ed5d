"""default_pad_value="otsu": the product's vectorised host sweep (cumsum) against the oracle's
restatement of the reference loop (spatial.py:2105-2168); the reference itself pins the oracle
through tests/golden/affine_fill_otsu.npz / affine_fill_mean.npz."""

import torch

from oracle import torch_port as tp
from torchio_b200.transforms.spatial import _border_faces, _otsu_border_mean


def test_otsu_border_mean_equals_the_reference_sweep():
    for seed in range(40):
        g = torch.Generator().manual_seed(seed)
        shape = [(9, 8, 7), (16, 14, 12), (5, 5, 5), (20, 3, 11)][seed % 4]
        x = torch.rand((1, 1, *shape), generator=g)
        if seed % 3 == 0:  # bimodal: background + foreground
            x = (x > 0.6).float() * torch.rand(x.shape, generator=g) + 0.05 * torch.rand(x.shape, generator=g)
        if seed % 5 == 0:  # many ties
            x = torch.round(x * 4) / 4
        if seed == 7:  # constant borders: no split improves the variance
            x = torch.full_like(x, 0.5)
        want = tp.border_mean(x[0, 0], True)
        got = _otsu_border_mean(_border_faces(x)[0].float().numpy())
        assert want == got, (seed, want, got)
