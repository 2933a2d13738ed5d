"""Diagnostic: timeline of host-resident (end-to-end) Compose steps: copies vs kernels."""
import os, sys, json, time, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
import torchio_b200 as tio
B = int(os.environ.get("B", "32"))
host = torch.empty((B, 1, 256, 256, 256), pin_memory=True).uniform_()
affs = [tio.AffineMatrix() for _ in range(B)]
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    pipe = tio.Compose([tio.Affine(scales=(0.9, 1.1), degrees=(-10, 10)), tio.ElasticDeformation(),
                        tio.BiasField(), tio.Blur(std=(0, 2)), tio.Noise(std=(0, 0.25)),
                        tio.Gamma(log_gamma=(-0.3, 0.3))], copy=False)
if "CHUNK" in os.environ:
    pipe.chunk_size = int(os.environ["CHUNK"])
def step():
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return pipe(tio.SubjectsBatch({"t1": tio.ImagesBatch(host, list(affs))}))
for _ in range(3):
    out = step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(3):
    out = step()
torch.cuda.synchronize()
print(f"wall per step {1e3*(time.perf_counter()-t0)/3:.1f} ms")
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    out = step()
    torch.cuda.synchronize()
prof.export_chrome_trace("gpurun_out/trace_e2e.json")
ev = json.load(open("gpurun_out/trace_e2e.json"))["traceEvents"]
gpu = [e for e in ev if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset") and "ts" in e]
gpu.sort(key=lambda e: e["ts"])
t0 = gpu[0]["ts"]
span = max(e["ts"] + e["dur"] for e in gpu) - t0
def busy(pred):
    return sum(e["dur"] for e in gpu if pred(e)) / 1e3
print(f"span {span/1e3:.1f} ms  HtoD {busy(lambda e: 'HtoD' in e['name']):.1f}  DtoH {busy(lambda e: 'DtoH' in e['name']):.1f}  kernels {busy(lambda e: e['cat']=='kernel'):.1f}")
for e in gpu:
    if e["cat"] == "gpu_memcpy" and e["dur"] > 500:
        print(f"  {(e['ts']-t0)/1e3:8.2f} ms  dur {e['dur']/1e3:7.2f}  {e['name'][:40]}  {e.get('args',{}).get('bytes', '')} B  {e.get('args',{}).get('memory bandwidth (GB/s)', '')} GB/s")
