"""Throughput of the neighbour kernels at the bench's size (B = 32, 1x256^3): tio_rescale
(Standardize / Normalize epilogue), tio_moments / tio_quantiles (sample 0), the label paths of K1
(nearest tile kernel, fused partial-volume mode, materialised one-hot path).
GPU box: python tools/aux_bench.py"""
import os
import sys
import warnings

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torchio_b200 as tio  # noqa: E402
from torchio_b200 import ops  # noqa: E402


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


B, S = 32, 256
x = torch.rand((B, 1, S, S, S), device="cuda")
gb = 8.0 * x.numel() / 1e9
ms = timeit(lambda: ops.rescale(x, lo=0.1, hi=0.9, sub=0.1, div=0.8, mul=2.0, add=-1.0))
print(f"tio_rescale (clamp, sub, div, mul, add)   {ms:7.3f} ms  {gb / ms * 1e3:6.0f} GB/s of 8 B/voxel", flush=True)
ms = timeit(lambda: ops.moments(x[0]))
print(f"tio_moments (sample 0, 64 MiB, incl. D2H) {ms:7.3f} ms", flush=True)
ms = timeit(lambda: ops.quantile_neighbours(x[0], [0.005, 0.995]))
print(f"tio_quantiles (sample 0, 2 quantiles)     {ms:7.3f} ms", flush=True)

idx = torch.arange(S, device="cuda")
ring = torch.minimum(idx, S - 1 - idx)
depth = torch.minimum(torch.minimum(ring[:, None, None], ring[None, :, None]), ring[None, None, :])
seg = (depth * 5 // (S // 2)).clamp_(0, 4).to(torch.int16)[None, None].expand(B, 1, -1, -1, -1).contiguous()
affines = [tio.AffineMatrix() for _ in range(B)]
for label_interpolation in ("nearest", "label"):
    for name, kw in (("Affine", dict(scales=(0.9, 1.1), degrees=(-10, 10))), ("ElasticDeformation", {})):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            t = getattr(tio, name)(label_interpolation=label_interpolation, copy=False, **kw)

            def step():
                batch = tio.SubjectsBatch({"seg": tio.ImagesBatch(seg, list(affines), image_class=tio.LabelMap)})
                return t(batch)

            torch.manual_seed(1)
            ms = timeit(step, 5)
        print(f"{name:20s} int16 LabelMap label_interpolation={label_interpolation:8s} {ms:7.3f} ms per call"
              f" (host sampling included)  {4.0 * seg.numel() / 1e9 / ms * 1e3:6.0f} GB/s of 4 B/voxel", flush=True)
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    t = tio.Resample(2, label_interpolation="label", antialias=True, copy=False)
    small = seg[:8].contiguous()
    ms = timeit(lambda: t(tio.SubjectsBatch({"seg": tio.ImagesBatch(small, list(affines[:8]), image_class=tio.LabelMap)})), 5)
print(f"Resample(2, antialias, label) on 8 x 256^3 (one-hot -> blur -> K1 -> argmax) {ms:7.3f} ms", flush=True)
