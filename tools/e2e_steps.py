"""Per-step wall time of `Compose.stream` on pinned host batches (is a slow run a few long stalls
or uniformly slow?) plus pinned/device allocator counters.  GPU box: python tools/e2e_steps.py"""
import os
import sys
import time
import warnings

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torchio_b200 as tio  # noqa: E402
from torchio_b200 import parallel  # noqa: E402

B, S = 32, 256
parallel.bind_to_gpu_numa(0)
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    pipe = tio.Compose([
        tio.Affine(scales=(0.9, 1.1), degrees=(-10, 10)), tio.ElasticDeformation(), tio.BiasField(),
        tio.Blur(std=(0, 2)), tio.Noise(std=(0, 0.25)), tio.Gamma(log_gamma=(-0.3, 0.3))], copy=False)
host = torch.rand((B, 1, S, S, S)).pin_memory()
affines = [tio.AffineMatrix() for _ in range(B)]
# a resident copy + a few resident steps first, like bench.py does before its e2e leg
resident = host.cuda()
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    for _ in range(5):
        pipe(tio.SubjectsBatch({"t1": tio.ImagesBatch(resident, list(affines))}))
torch.cuda.synchronize()


def batches(n):
    for _ in range(n):
        yield tio.SubjectsBatch({"t1": tio.ImagesBatch(host, list(affines))})


def stats():
    d = torch.cuda.memory_stats()
    h = torch.cuda.host_memory_stats() if hasattr(torch.cuda, "host_memory_stats") else {}
    return (d.get("num_device_alloc", -1), d.get("num_device_free", -1), d.get("reserved_bytes.all.current", 0) >> 20,
            h.get("num_host_alloc", -1), h.get("num_host_free", -1))


for depth in (1, 1, 0, 1):
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        times = []
        s0 = stats()
        t = time.perf_counter()
        for out in pipe.stream(batches(24), depth=depth):
            now = time.perf_counter()
            times.append((now - t) * 1e3)
            t = now
        torch.cuda.synchronize()
    print(f"depth {depth}: mean {sum(times[4:]) / len(times[4:]):.1f} ms  steps:", " ".join(f"{x:.0f}" for x in times),
          " alloc stats before/after (dev alloc, dev free, reserved MiB, host alloc, host free):", s0, stats(), flush=True)
