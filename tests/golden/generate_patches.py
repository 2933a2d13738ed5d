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


def run_samplers():
    """Weighted / label / grid samplers on one subject (subject 0 of the first case + a
    probability map): corners drawn, patch sums, grid locations with and without padding."""
    case = PATCH_CASES[0]
    t1, seg, affine = patch_subject_data(case, 0)
    g = torch.Generator().manual_seed(77)
    prob = torch.rand((1, *case["shape"]), generator=g) ** 4
    subject = tio.Subject(t1=tio.ScalarImage(t1, affine=affine.copy()), seg=tio.LabelMap(seg, affine=affine.copy()),
                          prob=tio.ScalarImage(prob, affine=affine.copy()))
    out = {}
    torch.manual_seed(5)
    ws = tio.WeightedSampler(subject, patch_size=case["patch_size"], probability_map="prob")
    out["weighted"] = [[list(p.patch_location.index), float(p.t1.data.double().sum())] for p in ws(subject, 6)]
    torch.manual_seed(6)
    ls = tio.LabelSampler(subject, patch_size=case["patch_size"], label_name="seg", label_probabilities={1: 1.0, 3: 2.0})
    out["label_probs"] = [[list(p.patch_location.index), float(p.seg.data.double().sum())] for p in ls(subject, 6)]
    torch.manual_seed(7)
    ls2 = tio.LabelSampler(subject, patch_size=case["patch_size"], label_name="seg")
    out["label_default"] = [list(p.patch_location.index) for p in ls2(subject, 4)]
    gs = tio.GridSampler(subject, patch_size=case["patch_size"], patch_overlap=(2, 4, 4))
    out["grid"] = [[list(loc.index), float(gs[i].t1.data.double().sum())] for i, loc in enumerate(gs.locations)]
    gp = tio.GridSampler(subject, patch_size=case["patch_size"], patch_overlap=(2, 4, 4), padding_mode="reflect")
    out["grid_padded_shape"] = list(gp.subject.spatial_shape)
    out["grid_padded"] = [[list(loc.index), float(gp[i].t1.data.double().sum())] for i, loc in enumerate(gp.locations)]
    (HERE / "patches_samplers.json").write_text(json.dumps(out))
    print("patches_samplers.json", len(json.dumps(out)), "bytes")


def main():
    run_samplers()
    for case in PATCH_CASES:
        arrays = run_case(case)
        path = HERE / f"patches_{case['name']}.npz"
        np.savez_compressed(path, **arrays)
        print(path.name, path.stat().st_size, "bytes")


if __name__ == "__main__":
    main()
