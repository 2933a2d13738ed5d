"""Golden vectors for the patch path: the UNMODIFIED reference's
UniformSampler / Queue / SubjectsLoader run on CPU.

TEST INFRASTRUCTURE.  Build container only (needs /root/reference + ``_shim/``):

    python tests/golden/generate_patches.py

Records, per case, the order in which the reference yields patches (subject id,
corner index) and a checksum + a few full patches, plus the collated batch shapes.
Inputs are regenerated from seeds by ``tests/golden_cases.py::patch_subject_data``.
"""

from __future__ import annotations

import json
import random
import sys
from pathlib import Path

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE / "_shim"))
sys.path.insert(1, "/root/reference/src")
sys.path.insert(2, str(HERE.parent))

import torchio as tio  # noqa: E402  (the reference)

from golden_cases import PATCH_CASES, patch_subject_data  # noqa: E402


def build_subjects(case):
    subjects = []
    for sid in range(case["num_subjects"]):
        t1, seg, affine = patch_subject_data(case, sid)
        subjects.append(tio.Subject(t1=tio.ScalarImage(t1, affine=affine.copy()),
                                    seg=tio.LabelMap(seg, affine=affine.copy()), sid=sid))
    return subjects


def run_case(case):
    subjects = build_subjects(case)
    sampler = tio.UniformSampler(subjects[0], patch_size=case["patch_size"])
    queue = tio.Queue(subjects, sampler, max_length=case["max_length"],
                      patches_per_volume=case["patches_per_volume"], num_workers=0,
                      shuffle_subjects=case["shuffle_subjects"], shuffle_patches=case["shuffle_patches"])
    torch.manual_seed(case["seed"])
    random.seed(case["seed"])
    order, sums, origins, first = [], [], [], []
    for patch in queue:
        loc = patch.patch_location
        order.append([int(patch.sid), *[int(v) for v in loc.index]])
        sums.append([float(patch.t1.data.double().sum()), float(patch.seg.data.double().sum())])
        origins.append([float(v) for v in patch.t1.affine.data[:3, 3]])
        if len(first) < 3:
            first.append((patch.t1.data.numpy().copy(), patch.seg.data.numpy().copy()))
    # the collated view of the same epoch (fresh seeds): batch shapes and locations
    torch.manual_seed(case["seed"])
    random.seed(case["seed"])
    loader = tio.SubjectsLoader(queue, batch_size=case["batch_size"])
    batch_shapes, batch_locs = [], []
    for batch in loader:
        batch_shapes.append(list(batch.t1.data.shape))
        batch_locs.append([[int(v) for v in loc.index] for loc in batch.metadata["patch_location"]])
    return {
        "meta": json.dumps({"order": order, "sums": sums, "origins": origins,
                            "batch_shapes": batch_shapes, "batch_locs": batch_locs,
                            "patches_per_epoch": queue.patches_per_epoch, "max_memory": queue.max_memory}),
        **{f"t1_{i}": a for i, (a, _) in enumerate(first)},
        **{f"seg_{i}": b for i, (_, b) in enumerate(first)},
    }


def main():
    for case in PATCH_CASES:
        arrays = run_case(case)
        path = HERE / f"patches_{case['name']}.npz"
        np.savez_compressed(path, **arrays)
        print(path.name, path.stat().st_size, "bytes")


if __name__ == "__main__":
    main()
