"""Diagnostic: GPU timeline (kernels + memcpys) of free-running Compose steps via torch.profiler."""
import os, sys, json, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
import torchio_b200 as tio
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
def step():
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return pipe(tio.SubjectsBatch({"t1": tio.ImagesBatch(x, list(affs))}))
for _ in range(4):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(int(os.environ.get("TRACE_STEPS", "4"))):
        out = step()
    torch.cuda.synchronize()
os.makedirs("gpurun_out", exist_ok=True)
prof.export_chrome_trace("gpurun_out/trace.json")
ev = json.load(open("gpurun_out/trace.json"))["traceEvents"]
gpu = [e for e in ev if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset") and "ts" in e]
gpu.sort(key=lambda e: e["ts"])
t0 = gpu[0]["ts"]
prev_end = None
busy = sum(e["dur"] for e in gpu)
span = gpu[-1]["ts"] + gpu[-1]["dur"] - t0
gaps = []
pe = None
for e in gpu:
    if pe is not None and e["ts"] - pe > 100:
        gaps.append(((e["ts"] - t0) / 1e3, (e["ts"] - pe) / 1e3, e["name"][:40]))
    pe = e["ts"] + e["dur"]
print(f"span {span/1e3:.2f} ms  busy {busy/1e3:.2f} ms  idle {(span-busy)/1e3:.2f} ms; gaps > 0.1 ms: {gaps}")
if os.environ.get("TRACE_SUMMARY"):
    sys.exit(0)
for e in gpu:
    gap = 0 if prev_end is None else e["ts"] - prev_end
    name = e["name"][:60]
    print(f"{(e['ts']-t0)/1e3:9.3f} ms  dur {e['dur']/1e3:8.3f} ms  gap {gap/1e3:7.3f} ms  {e['cat']:10s} {name}")
    prev_end = e["ts"] + e["dur"]
