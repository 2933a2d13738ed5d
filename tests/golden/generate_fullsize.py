"""Golden vectors at BASELINE.json's own volume size (256^3) from the UNMODIFIED reference.

TEST INFRASTRUCTURE.  Build container only (reads /root/reference; see generate.py for the
shim).  ~10 minutes and ~25 GB of host memory on 8 cores:

    python tests/golden/generate_fullsize.py

The full outputs stay out of the repository (64 MiB per volume): per case the fixture holds the
sampled params (history), a strided lattice of every output, a dense corner block (padding and
fill decisions) and a dense centre block (`golden_cases.full_views`), the output affines, and
SHA-256 digests of the full tensors (labels are bit-exact, so their digest is a full check).
"""

from __future__ import annotations

import hashlib
import json
import sys
import time
import warnings
from pathlib import Path

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE / "_shim"))
sys.path.insert(1, "/root/reference/src")
sys.path.insert(2, str(HERE.parent))

import torchio as tio  # noqa: E402  (the reference)

from generate import _make_transform, _to_reference_batch  # noqa: E402
from golden_cases import FULL_CASES, build_inputs, full_views  # noqa: E402


def run_case(case):
    batch = _to_reference_batch(build_inputs(case))
    transform = _make_transform(case["transform"])
    torch.manual_seed(case["seed"])
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        out = transform(batch)
    history = [{"name": t.name, "params": t.params} for t in out.applied_transforms]
    arrays = {"history": np.frombuffer(json.dumps(history).encode(), dtype=np.uint8)}
    digests = {}
    for name, img_batch in out.images.items():
        data = img_batch.data.contiguous()
        for view, t in full_views(data).items():
            arrays[f"{view}_{name}"] = t.contiguous().numpy()
        arrays[f"aff_{name}"] = np.stack([a.numpy() for a in img_batch.affines])
        arrays[f"minmax_{name}"] = np.array([float(data.min()), float(data.max())])
        digests[name] = hashlib.sha256(data.numpy().tobytes()).hexdigest()
    arrays["sha256"] = np.frombuffer(json.dumps(digests).encode(), dtype=np.uint8)
    return arrays


def main():
    torch.set_num_threads(8)
    for case in FULL_CASES:
        t0 = time.time()
        arrays = run_case(case)
        path = HERE / f"{case['name']}.npz"
        np.savez_compressed(path, **arrays)
        print(f"{case['name']:28s} {path.stat().st_size / 1024:8.1f} KiB  {time.time() - t0:6.1f} s", flush=True)


if __name__ == "__main__":
    main()
