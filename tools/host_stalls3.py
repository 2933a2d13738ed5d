"""Diagnostic: cProfile the steps that stall on the host."""
import cProfile, io, os, pstats, sys, time, warnings
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
def step():
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return pipe(tio.SubjectsBatch({"t1": tio.ImagesBatch(x, list(affs))}))
for _ in range(3):
    step()
torch.cuda.synchronize()
shown = 0
for it in range(40):
    pr = cProfile.Profile()
    t0 = time.perf_counter()
    pr.enable(); out = step(); pr.disable()
    dt = time.perf_counter() - t0
    if dt > 30e-3 and shown < 2:
        shown += 1
        s = io.StringIO()
        pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(8)
        print(f"step {it} host {1e3*dt:.1f} ms"); print(s.getvalue()[:2500])
torch.cuda.synchronize()
