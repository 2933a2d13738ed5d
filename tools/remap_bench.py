"""HBM throughput of the index-move kernels at the bench's volume size (B = 32, 1x256^3):
tio_remap (Flip, Crop, Pad) and tio_crop_patches (8 patches of 128^3 per volume).
GPU box: python tools/remap_bench.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
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
for dtype in (torch.float32, torch.int16, torch.uint8):
    x = (torch.rand((B, 1, S, S, S), device="cuda") * 100).to(dtype)
    es = x.element_size()
    flip = torch.tensor([5] * B, dtype=torch.uint8, device="cuda")  # axes I and K
    cases = {
        "flip I+K": lambda: ops.remap(x, (S, S, S), (0, 0, 0), flip=flip),
        "flip I": lambda: ops.remap(x, (S, S, S), (0, 0, 0), flip=torch.ones(B, dtype=torch.uint8, device="cuda")),
        "crop 16 (aligned)": lambda: ops.remap(x, (S - 32, S - 32, S - 32), (-16, -16, -16)),
        "crop 15/17 (unaligned)": lambda: ops.remap(x, (S - 32, S - 32, S - 32), (-15, -15, -15)),
        "pad 16 constant": lambda: ops.remap(x, (S + 32, S + 32, S + 32), (16, 16, 16), fill=1),
        "pad 16 reflect": lambda: ops.remap(x, (S + 32, S + 32, S + 32), (16, 16, 16), mode="reflect"),
    }
    for name, fn in cases.items():
        out = fn()
        ms = timeit(fn)
        gb = 2.0 * out.numel() * es / 1e9
        print(f"REMAP {str(dtype):14s} {name:24s} {ms:7.3f} ms  {gb / ms * 1e3:7.0f} GB/s (2 x output bytes)", flush=True)
    vol = x[0]
    rng = np.random.default_rng(0)
    corners = rng.integers(0, S - 128, (8, 3))
    fn = lambda: ops.crop_patches(vol, corners, (128, 128, 128))
    ms = timeit(fn, 20)
    gb = 2.0 * 8 * 128**3 * es / 1e9
    print(f"CROP  {str(dtype):14s} 8 x 128^3 of one volume  {ms:7.3f} ms  {gb / ms * 1e3:7.0f} GB/s", flush=True)
    del x
