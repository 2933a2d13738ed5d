"""Diagnostic: which host-side call blocks while the pipeline free-runs?"""
import os, sys, time, warnings, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torchio_b200 as tio
from torchio_b200 import ops, _native
os.environ.setdefault("TIO_B200_NOISE", "philox")
B = 32
dev = torch.device("cuda")
x = torch.rand((B, 1, 256, 256, 256), device=dev)
affs = [tio.AffineMatrix() for _ in range(B)]
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    pipe = tio.Compose([tio.Affine(scales=(0.9, 1.1), degrees=(-10, 10)), tio.ElasticDeformation(),
                        tio.BiasField(), tio.Blur(std=(0, 2)), tio.Noise(std=(0, 0.25)),
                        tio.Gamma(log_gamma=(-0.3, 0.3))], copy=False)
log = []
def wrap(obj, name, label):
    raw = getattr(obj, name)
    def f(*a, **k):
        t0 = time.perf_counter(); r = raw(*a, **k); log.append((label, time.perf_counter() - t0)); return r
    setattr(obj, name, f)
wrap(_native, "call", "native.call")
wrap(ops, "upload", "ops.upload")
wrap(torch, "empty", "torch.empty")
wrap(torch, "empty_like", "torch.empty_like")
wrap(torch, "full", "torch.full")
def step():
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return pipe(tio.SubjectsBatch({"t1": tio.ImagesBatch(x, list(affs))}))
for _ in range(3):
    step()
torch.cuda.synchronize()
for it in range(14):
    log.clear()
    t0 = time.perf_counter(); out = step(); dt = time.perf_counter() - t0
    agg = collections.OrderedDict()
    for label, d in log:
        agg[label] = agg.get(label, 0) + d
    worst = max(log, key=lambda z: z[1])
    print(f"step {it:2d} host {1e3*dt:6.1f} ms | " + " ".join(f"{k}={1e3*v:.1f}" for k, v in agg.items())
          + f" | worst {worst[0]} {1e3*worst[1]:.1f}")
torch.cuda.synchronize()
