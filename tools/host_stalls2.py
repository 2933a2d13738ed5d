"""Diagnostic: find the occasional tens-of-ms host stall (GC? allocator?) in free-running steps."""
import gc, os, sys, time, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torchio_b200 as tio
B = 32
dev = torch.device("cuda")
x = torch.rand((B, 1, 256, 256, 256), device=dev)
affs = [tio.AffineMatrix() for _ in range(B)]
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    pipe = tio.Compose([tio.Affine(scales=(0.9, 1.1), degrees=(-10, 10)), tio.ElasticDeformation(),
                        tio.BiasField(), tio.Blur(std=(0, 2)), tio.Noise(std=(0, 0.25)),
                        tio.Gamma(log_gamma=(-0.3, 0.3))], copy=False)
gc_log = []
_t = [0.0]
def cb(phase, info):
    if phase == "start":
        _t[0] = time.perf_counter()
    else:
        gc_log.append((info["generation"], time.perf_counter() - _t[0], info["collected"]))
gc.callbacks.append(cb)
def step():
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return pipe(tio.SubjectsBatch({"t1": tio.ImagesBatch(x, list(affs))}))
for mode in ("default", "gc-disabled"):
    if mode == "gc-disabled":
        gc.collect(); gc.disable()
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    for it in range(24):
        gc_log.clear()
        s0 = torch.cuda.memory_stats()
        t0 = time.perf_counter(); out = step(); dt = time.perf_counter() - t0
        s1 = torch.cuda.memory_stats()
        d_alloc = s1["num_device_alloc"] - s0["num_device_alloc"]
        d_free = s1["num_device_free"] - s0["num_device_free"]
        if dt > 6e-3 or gc_log or d_alloc or d_free:
            print(f"{mode} step {it:2d} host {1e3*dt:6.1f} ms  gc={[(g, round(1e3*d,1), c) for g,d,c in gc_log]} cudaMalloc={d_alloc} cudaFree={d_free}")
    torch.cuda.synchronize()
    print(mode, "done")
