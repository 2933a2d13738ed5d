#!/usr/bin/env python
"""bench.py — augmented voxels/s of the 256^3 fp32 Compose pipeline on B200.

    python bench.py --gpus N --steps K --warmup W            # our arm (CUDA kernels)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path

A "step" is one pass of the hot path over one batch of synthetic volumes:
``Compose([Affine, ElasticDeformation, BiasField, Blur, Noise, Gamma])`` on
``(B, 1, 256, 256, 256)`` fp32 (BASELINE.json configs[2]; ``--workload config2``
runs configs[1] = the first two transforms).  One process per GPU; every rank
augments its own batch (weak scaling, no data-path collective).

Prints ONE JSON line on rank 0 (keys documented in DESIGN.md §Measurement).
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
import warnings
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

VOL = 256
ALGO_BYTES_PER_VOXEL_RESAMPLE = 8  # one fp32 read + one fp32 write


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="full", choices=["full", "config2"])
    ap.add_argument("--batch", type=int, default=32, help="volumes per GPU per step")
    ap.add_argument("--size", type=int, default=VOL)
    ap.add_argument("--noise", default=os.environ.get("TIO_B200_NOISE", "exact"),
                    choices=["exact", "philox"])
    ap.add_argument("--cpu-sample-batch", type=int, default=2)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-clocks", action="store_true", help="diagnostic: skip the nvidia-smi sampler")
    ap.add_argument("--labels", action="store_true",
                    help="configs[3] shape: add an int16 LabelMap (nearest-neighbour resample) to every volume")
    return ap.parse_args()


def pipeline_spec(workload):
    spec = [
        ("Affine", {"scales": (0.9, 1.1), "degrees": (-10, 10)}),
        ("ElasticDeformation", {}),
    ]
    if workload == "full":
        spec += [
            ("BiasField", {}),
            ("Blur", {"std": (0, 2)}),
            ("Noise", {"std": (0, 0.25)}),
            ("Gamma", {"log_gamma": (-0.3, 0.3)}),
        ]
    return spec


def synth_volumes(batch, size, pin):
    """(B,1,S,S,S) fp32 in [0,1): torch.rand(seed 1000+b) per element, on the host."""
    out = torch.empty((batch, 1, size, size, size), dtype=torch.float32, pin_memory=pin)
    for b in range(batch):
        g = torch.Generator().manual_seed(1000 + b)
        torch.rand((1, size, size, size), generator=g, out=out[b])
    return out


# ----------------------------------------------------------------------------
# clocks
# ----------------------------------------------------------------------------


class ClockSampler:
    """nvidia-smi sampled every 200 ms during the timed region."""

    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                 "-lms", "200", "-i", str(self.index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def wait_first_sample(self, timeout=10.0):
        """nvidia-smi takes a while to start: do not let its start-up overlap the timed region."""
        t0 = time.perf_counter()
        while self.proc is not None and not self.lines and time.perf_counter() - t0 < timeout:
            time.sleep(0.05)

    def stop(self, t_begin=None, t_end=None):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        window = [l for t, l in self.lines
                  if (t_begin is None or t >= t_begin) and (t_end is None or t <= t_end + 0.25)]
        if not window:  # region shorter than the sampling period: take the nearest samples
            window = [l for _, l in self.lines[-2:]]
        for line in window:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {
            "sm_mhz": sm[len(sm) // 2] if sm else None,
            "sm_max_mhz": max(smax) if smax else None,
            "reasons": sorted(reasons),
            "samples": len(sm),
        }


# ----------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------


def run_b200(args, rank, world, local_rank):
    import torch.distributed as dist

    import torchio_b200 as tio
    from torchio_b200 import ops

    os.environ["TIO_B200_NOISE"] = args.noise
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        pipeline = tio.Compose(
            [getattr(tio, n)(**kw) for n, kw in pipeline_spec(args.workload)], copy=False)
    host = synth_volumes(args.batch, args.size, pin=True)
    resident = host.to(dev)
    affines = [tio.AffineMatrix() for _ in range(args.batch)]
    voxels = args.batch * args.size**3

    labels_host = labels_dev = None
    if args.labels:  # concentric boxes, values 0..4 (SURVEY.md §8d synthetic label)
        idx = torch.arange(args.size)
        ring = torch.minimum(idx, args.size - 1 - idx)
        depth = torch.minimum(torch.minimum(ring[:, None, None], ring[None, :, None]), ring[None, None, :])
        one = (depth * 5 // max(args.size // 2, 1)).clamp_(0, 4).to(torch.int16)
        labels_host = one[None, None].expand(args.batch, 1, -1, -1, -1).contiguous().pin_memory()
        labels_dev = labels_host.to(dev)

    def make_batch(tensor):
        images = {"t1": tio.ImagesBatch(tensor, list(affines))}
        if args.labels:
            images["seg"] = tio.ImagesBatch(labels_dev if tensor.is_cuda else labels_host, list(affines),
                                            image_class=tio.LabelMap)
        return tio.SubjectsBatch(images)

    # event hooks around the dominant kernel (K1) inside the real step
    k1_events = []
    raw_resample = ops.resample

    def timed_resample(*a, **kw):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        out = raw_resample(*a, **kw)
        e.record()
        k1_events.append((s, e))
        return out

    ops.resample = timed_resample

    def step(tensor):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            return pipeline(make_batch(tensor))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank) if (rank == 0 and not args.no_clocks) else None
    if sampler:
        sampler.start()
    torch.manual_seed(1234 + rank)
    for _ in range(args.warmup):
        out = step(resident)
    if sampler:
        sampler.wait_first_sample()
    k1_events.clear()
    barrier()
    launches0 = ops.launches()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    wall_begin = time.perf_counter()
    t0.record()
    for _ in range(args.steps):
        out = step(resident)
    t1.record()
    host_issue_ms = (time.perf_counter() - wall_begin) * 1e3 / args.steps
    barrier()
    wall_end = time.perf_counter()
    ms = t0.elapsed_time(t1)
    launches = ops.launches() - launches0
    clocks = sampler.stop(wall_begin, wall_end) if sampler else None
    k1_ms = [s.elapsed_time(e) for s, e in k1_events]
    del out

    # end to end through the public call with HOST buffers: pinned input ->
    # H2D -> kernels -> D2H into pinned output, every step
    e2e = None
    if not args.no_e2e:
        torch.manual_seed(4321 + rank)
        for _ in range(max(1, min(args.warmup, 2))):
            res = step(host)
        barrier()
        w0 = time.perf_counter()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            res = step(host)
        e1.record()
        barrier()
        e2e_ms = max(e0.elapsed_time(e1), (time.perf_counter() - w0) * 1e3)
        assert res.images["t1"].data.device.type == "cpu"
        moved = host.numel() * 4 + (labels_host.numel() * 2 if args.labels else 0)
        e2e = {"ms": e2e_ms, "bytes_in": moved, "bytes_out": moved}
        del res
    ops.resample = raw_resample

    if world > 1:
        t = torch.tensor([ms, e2e["ms"] if e2e else 0.0], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t[0])
        if e2e:
            e2e["ms"] = float(t[1])
    if rank != 0:
        return None

    peaks = {}
    peaks_path = ROOT / "MEASURED_PEAKS.json"
    if peaks_path.exists():
        peaks = json.loads(peaks_path.read_text())
    peak = float(peaks.get("hbm_gbs", 6650.0))
    k1_avg_ms = sum(k1_ms) / len(k1_ms) if k1_ms else float("nan")
    achieved = ALGO_BYTES_PER_VOXEL_RESAMPLE * voxels / (k1_avg_ms * 1e-3) / 1e9
    value = world * voxels * args.steps / (ms * 1e-3)
    line = {
        "metric": "augmented voxels/sec on 256^3 fp32 Compose pipeline",
        "value": value,
        "unit": "voxels/s",
        "n_gpus": world,
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": ms / args.steps,
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "f32",
        "data": "synthetic",
        "config": {
            "workload": ("configs[2]: batch %d of 1x%d^3 fp32, Compose(Affine, ElasticDeformation,"
                         " BiasField, Blur, Noise, Gamma) per GPU" % (args.batch, args.size))
            + (" + int16 LabelMap (nearest)" if args.labels else "")
            if args.workload == "full" else
            ("configs[1]: batch %d of 1x%d^3 fp32, Compose(Affine, ElasticDeformation) per GPU"
             % (args.batch, args.size)),
            "global_batch": world * args.batch,
            "parallelism": f"dp{world} (independent volumes, no data-path collective)",
            "noise_normals": args.noise,
            "l2_policy": "inputs (%.1f GiB/GPU) larger than L2 (126 MB)" % (voxels * 4 / 2**30),
            "includes": "host param sampling + table upload + all kernels of the step",
        },
        "gpu_launches": launches,
        "host_issue_ms_per_step": host_issue_ms,
        "roofline": {
            "kernel": "resample_kernel (K1, %d launches in the timed region)" % len(k1_ms),
            "bound": "hbm",
            "achieved": achieved,
            "peak": peak,
            "unit": "GB/s",
            "frac": achieved / peak,
            "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s",
            "algorithmic_bytes_per_launch": ALGO_BYTES_PER_VOXEL_RESAMPLE * voxels,
            "avg_launch_ms": k1_avg_ms,
            "share_of_step": sum(k1_ms) / ms if k1_ms else None,
            "traffic": k1_traffic(args),
        },
        "clocks": clocks,
    }
    if e2e:
        line["e2e"] = {
            "value": world * voxels * args.steps / (e2e["ms"] * 1e-3),
            "unit": "voxels/s",
            "h2d_bytes_per_step": e2e["bytes_in"],
            "d2h_bytes_per_step": e2e["bytes_out"],
            "ms_per_step": e2e["ms"] / args.steps,
        }
    if not args.no_cpu_baseline and world == 1:
        line["cpu_baseline"] = cpu_reference(args, steps=1, warmup=0)
    return line


# ----------------------------------------------------------------------------
# reference arm / CPU baseline: the oracle's torch-op port = the op sequence
# the reference executes on the host (oracle/torch_port.py)
# ----------------------------------------------------------------------------


def k1_traffic(args):
    """DRAM bytes per K1 launch from the committed `ncu --set full` capture of this
    workload (profiles/r1_k1_traffic.json), or None when the run differs from it."""
    path = ROOT / "profiles" / "r1_k1_traffic.json"
    if not path.exists() or args.batch != 32 or args.size != VOL:
        return None
    return json.loads(path.read_text()).get("bytes_per_launch")


def cpu_reference(args, steps, warmup):
    import numpy as np

    import torchio_b200 as tio
    from oracle import torch_port

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    # bounded sample: ~13 s per step at batch 2 on the GPU box's 128 host threads; keep the
    # whole --steps run within a few minutes by dropping to one volume for long runs
    b = args.cpu_sample_batch if steps * args.cpu_sample_batch <= 24 else 1
    size = args.size
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        transforms = [getattr(tio, n)(**kw) for n, kw in pipeline_spec(args.workload)]
    data = synth_volumes(b, size, pin=False)
    batch = tio.SubjectsBatch(
        {"t1": tio.ImagesBatch(data, [tio.AffineMatrix() for _ in range(b)])})

    def one_step():
        # sample params with the product's host code (identical to the
        # reference's sampling, tests/test_host_params.py), replay on CPU ops
        history = []
        for t in transforms:
            torch.rand(1)
            history.append({"name": type(t).__name__, "params": t.make_params(batch)})
        images = {"t1": {"kind": "scalar", "data": data,
                         "affines": [np.eye(4) for _ in range(b)]}}
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            torch_port.replay(images, history)

    torch.manual_seed(99)
    for _ in range(warmup):
        one_step()
    t0 = time.perf_counter()
    for _ in range(steps):
        one_step()
    dt = time.perf_counter() - t0
    return {
        "value": b * size**3 * steps / dt,
        "unit": "voxels/s",
        "cores": cores,
        "kind": "port",
        "sample": f"{steps} step(s) of batch {b} x 1x{size}^3 fp32, same Compose, "
                  f"torch {torch.__version__} CPU ops, {torch.get_num_threads()} threads",
        "seconds": dt,
    }


def run_reference(args, rank, world):
    if rank != 0:
        return None
    base = cpu_reference(args, steps=args.steps, warmup=min(args.warmup, 1))
    return {
        "impl": "reference",
        "metric": "augmented voxels/sec on 256^3 fp32 Compose pipeline",
        "value": base["value"],
        "unit": "voxels/s",
        "n_gpus": world,
        "steps": args.steps,
        "warmup": min(args.warmup, 1),
        "ms_per_step": base["seconds"] / args.steps * 1e3,
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "f32",
        "data": "synthetic",
        "config": {
            "workload": "bounded sample of the same Compose: batch %d of 1x%d^3 per step on the"
                        " host cores (rank 0 only)" % (args.cpu_sample_batch, args.size),
            "parallelism": "host threads",
        },
        "cpu_baseline": {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": base["value"], "unit": "voxels/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference":
        line = run_reference(args, rank, world)
        if line is not None:
            print(json.dumps(line), flush=True)
        return
    if world > 1:
        import torch.distributed as dist

        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        line = run_b200(args, rank, world, local_rank)
        if line is not None:
            print(json.dumps(line), flush=True)
    finally:
        if world > 1:
            import torch.distributed as dist

            dist.destroy_process_group()


if __name__ == "__main__":
    main()
