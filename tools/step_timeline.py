"""Diagnostic: host issue time vs GPU time per Compose step (not a benchmark)."""
import os, sys, time, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torchio_b200 as tio
os.environ.setdefault("TIO_B200_NOISE", "philox")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
dev = torch.device("cuda")
x = torch.rand((B, 1, 256, 256, 256), device=dev)
affs = [tio.AffineMatrix() for _ in range(B)]
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    names = [("Affine", dict(scales=(0.9, 1.1), degrees=(-10, 10))), ("ElasticDeformation", {}),
             ("BiasField", {}), ("Blur", dict(std=(0, 2))), ("Noise", dict(std=(0, 0.25))),
             ("Gamma", dict(log_gamma=(-0.3, 0.3)))]
    ts = [getattr(tio, n)(**kw) for n, kw in names]
    pipe = tio.Compose(ts, copy=False)
def step():
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return pipe(tio.SubjectsBatch({"t1": tio.ImagesBatch(x, list(affs))}))
for _ in range(3):
    step()
torch.cuda.synchronize()
for it in range(4):
    t0 = time.perf_counter(); out = step(); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"step {it}: host issue {1e3*(t1-t0):7.2f} ms   issue+drain {1e3*(t2-t0):7.2f} ms")
# per-transform host/GPU split
batch = tio.SubjectsBatch({"t1": tio.ImagesBatch(x, list(affs))})
for t in ts:
    torch.cuda.synchronize(); t0 = time.perf_counter()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        batch = t._forward_batch(batch)
    t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"{type(t).__name__:20s} host {1e3*(t1-t0):7.2f} ms   total {1e3*(t2-t0):7.2f} ms")
