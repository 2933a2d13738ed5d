"""End-to-end (pinned host in -> device -> pinned host out) step time of the bench pipeline for
slice sizes x {plain call, Compose.stream(depth)}.  GPU box: python tools/e2e_sweep.py [B]"""
import os
import sys
import time
import warnings

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torchio_b200 as tio  # noqa: E402
from torchio_b200 import parallel  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
S = 256
parallel.bind_to_gpu_numa(0)
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    pipe = tio.Compose([
        tio.Affine(scales=(0.9, 1.1), degrees=(-10, 10)), tio.ElasticDeformation(), tio.BiasField(),
        tio.Blur(std=(0, 2)), tio.Noise(std=(0, 0.25)), tio.Gamma(log_gamma=(-0.3, 0.3))], copy=False)
g = torch.Generator().manual_seed(1)
host = torch.empty((B, 1, S, S, S), pin_memory=True)
for b in range(B):
    host[b].copy_(torch.rand((1, S, S, S), generator=g))
affines = [tio.AffineMatrix() for _ in range(B)]


def batches(n):
    for _ in range(n):
        yield tio.SubjectsBatch({"t1": tio.ImagesBatch(host, list(affines))})


def run(chunk_mb, depth, steps=10):
    pipe.chunk_bytes = chunk_mb << 20
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        torch.manual_seed(3)
        for _ in pipe.stream(batches(4), depth=max(depth, 0)):
            pass
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        if depth < 0:
            for b in batches(steps):
                out = pipe(b)
        else:
            for out in pipe.stream(batches(steps), depth=depth):
                pass
        torch.cuda.synchronize()
    return (time.perf_counter() - t0) * 1e3 / steps


for chunk_mb in (64, 128, 256, 512):
    row = [f"{run(chunk_mb, d):7.2f}" for d in (-1, 1, 2)]
    print(f"chunk {chunk_mb:4d} MB   plain {row[0]}  stream(1) {row[1]}  stream(2) {row[2]}  ms/step", flush=True)
