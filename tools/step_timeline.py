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
# free-running: does the host stall when it runs ahead of the GPU?
torch.cuda.synchronize()
ts_ = []
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
t_start = time.perf_counter()
for it in range(12):
    t0 = time.perf_counter(); out = step(); ts_.append(time.perf_counter() - t0)
e1.record(); torch.cuda.synchronize()
print("free-running host issue ms per step:", " ".join(f"{1e3*t:.1f}" for t in ts_))
print(f"free-running GPU ms/step {e0.elapsed_time(e1)/12:.2f}  wall {1e3*(time.perf_counter()-t_start)/12:.2f}")
print("allocated GiB", torch.cuda.memory_allocated()/2**30, "reserved GiB", torch.cuda.memory_reserved()/2**30)
# per-op GPU time inside the free-running pipeline (CUDA events on the launch stream)
from torchio_b200 import ops
import torchio_b200.transforms.spatial as _sp, torchio_b200.transforms.intensity as _it
evs = {}
def wrap(mod, name):
    raw = getattr(ops, name)
    def f(*a, **k):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); r = raw(*a, **k); e.record(); evs.setdefault(name, []).append((s, e)); return r
    setattr(ops, name, f)
for nm in ("resample", "min_sample0", "intensity_fused", "upload"):
    wrap(ops, nm)
torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
for it in range(10):
    out = step()
e1.record(); torch.cuda.synchronize()
tot = e0.elapsed_time(e1) / 10
acc = 0
for nm, lst in evs.items():
    ms = sum(s.elapsed_time(e) for s, e in lst) / 10
    acc += ms
    print(f"  {nm:18s} {ms:7.3f} ms/step over {len(lst)//10} calls")
print(f"  sum {acc:.3f} ms/step   step {tot:.3f} ms   unaccounted {tot-acc:.3f} ms")
