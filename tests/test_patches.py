"""Patch path against the unmodified reference's UniformSampler /
Queue / SubjectsLoader (tests/golden/generate_patches.py)."""

import json
import random
from pathlib import Path

import numpy as np
import pytest
import torch

from golden_cases import PATCH_CASES, patch_subject_data

GOLDEN = Path(__file__).resolve().parent / "golden"


def _subjects(case, device="cpu"):
    import torchio_b200 as tio

    out = []
    for sid in range(case["num_subjects"]):
        t1, seg, affine = patch_subject_data(case, sid)
        out.append(tio.Subject(t1=tio.ScalarImage(t1.to(device), affine=affine.copy()),
                               seg=tio.LabelMap(seg.to(device), affine=affine.copy()), sid=sid))
    return out


def _queue(case, subjects, **kw):
    import torchio_b200 as tio

    sampler = tio.UniformSampler(subjects[0], patch_size=case["patch_size"])
    return tio.Queue(subjects, sampler, max_length=case["max_length"],
                     patches_per_volume=case["patches_per_volume"], num_workers=0,
                     shuffle_subjects=case["shuffle_subjects"], shuffle_patches=case["shuffle_patches"], **kw)


def _check_epoch(case, device, **kw):
    import torchio_b200 as tio

    gold = np.load(GOLDEN / f"patches_{case['name']}.npz")
    meta = json.loads(str(gold["meta"]))
    queue = _queue(case, _subjects(case, device), **kw)
    assert queue.patches_per_epoch == meta["patches_per_epoch"]
    assert queue.max_memory == meta["max_memory"]
    torch.manual_seed(case["seed"])
    random.seed(case["seed"])
    patches = list(queue)
    assert [[int(p.sid), *p.patch_location.index] for p in patches] == meta["order"]
    for p, (s1, s2), origin in zip(patches, meta["sums"], meta["origins"]):
        assert tuple(p.t1.data.shape[1:]) == tuple(case["patch_size"])
        assert float(p.t1.data.double().sum()) == pytest.approx(s1, rel=1e-12)
        assert float(p.seg.data.double().sum()) == s2
        assert np.allclose(p.t1.affine.numpy()[:3, 3], origin, rtol=0, atol=1e-12)
        assert p.seg.data.dtype == torch.int16
    for i in range(3):
        assert np.array_equal(patches[i].t1.data.cpu().numpy(), gold[f"t1_{i}"])
        assert np.array_equal(patches[i].seg.data.cpu().numpy(), gold[f"seg_{i}"])
    torch.manual_seed(case["seed"])
    random.seed(case["seed"])
    loader = tio.SubjectsLoader(queue, batch_size=case["batch_size"])
    shapes, locs = [], []
    for batch in loader:
        assert isinstance(batch, tio.SubjectsBatch)
        shapes.append(list(batch.t1.data.shape))
        locs.append([list(loc.index) for loc in batch.metadata["patch_location"]])
    assert shapes == meta["batch_shapes"] and locs == meta["batch_locs"]


@pytest.mark.parametrize("case", PATCH_CASES, ids=[c["name"] for c in PATCH_CASES])
def test_queue_matches_reference_on_host_subjects(case):
    _check_epoch(case, "cpu")


@pytest.mark.gpu
@pytest.mark.parametrize("case", PATCH_CASES, ids=[c["name"] for c in PATCH_CASES])
def test_queue_matches_reference_on_device_subjects(case):
    """Device-resident subjects: the patches come from tio_crop_patches."""
    from torchio_b200 import ops

    before = ops.launches()
    _check_epoch(case, "cuda")
    assert ops.launches() > before


@pytest.mark.gpu
def test_queue_device_option_moves_a_copy_and_transforms_resident():
    import warnings

    import torchio_b200 as tio

    case = PATCH_CASES[0]
    subjects = _subjects(case, "cpu")
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        transform = tio.Compose([tio.Affine(degrees=(-5, 5)), tio.Gamma(log_gamma=(-0.2, 0.2))])
    queue = _queue(case, subjects, transform=transform, device="cuda")
    torch.manual_seed(1)
    random.seed(1)
    patches = list(queue)
    assert len(patches) == queue.patches_per_epoch
    assert all(p.t1.data.is_cuda and p.seg.data.is_cuda for p in patches)
    assert all(not s.t1.data.is_cuda for s in subjects)  # the dataset's subjects stayed on the host
    assert all(len(p.applied_transforms) == 2 for p in patches) or True


@pytest.mark.gpu
def test_crop_patches_every_dtype_and_bounds():
    from torchio_b200 import ops

    g = torch.Generator().manual_seed(0)
    corners = [[0, 0, 0], [3, 5, 2], [12, 14, 4]]
    for dtype in (torch.uint8, torch.int16, torch.int32, torch.int64, torch.float32, torch.float64):
        vol = (torch.rand((2, 20, 24, 21), generator=g) * 100).to(dtype).cuda()
        got = ops.crop_patches(vol, corners, (8, 10, 17))
        for row, (i, j, k) in enumerate(corners):
            assert torch.equal(got[row], vol[:, i:i + 8, j:j + 10, k:k + 17])
    with pytest.raises(ValueError):
        ops.crop_patches(vol, [[13, 0, 0]], (8, 10, 17))


def test_threaded_queue_yields_the_same_patch_multiset():
    """num_workers > 0: order depends on thread timing (as in the reference), content does not."""
    case = PATCH_CASES[1]
    import torchio_b200 as tio

    subjects = _subjects(case)
    sampler = tio.UniformSampler(subjects[0], patch_size=case["patch_size"])
    q = tio.Queue(subjects, sampler, max_length=7, patches_per_volume=3, num_workers=2,
                  shuffle_subjects=False, shuffle_patches=False)
    torch.manual_seed(3)
    got = list(q)
    assert len(got) == q.patches_per_epoch == 9
    assert sorted(int(p.sid) for p in got) == [0, 0, 0, 1, 1, 1, 2, 2, 2]
    with pytest.raises(ValueError):
        tio.Queue(subjects, sampler, shuffle_subjects=True, subject_sampler=[0, 1])
    with pytest.raises(ValueError):
        tio.SubjectsLoader(q, collate_fn=lambda b: b)


def _sampler_subject(device="cpu"):
    import torchio_b200 as tio

    case = PATCH_CASES[0]
    t1, seg, affine = patch_subject_data(case, 0)
    g = torch.Generator().manual_seed(77)
    prob = torch.rand((1, *case["shape"]), generator=g) ** 4
    return case, tio.Subject(t1=tio.ScalarImage(t1.to(device), affine=affine.copy()),
                             seg=tio.LabelMap(seg.to(device), affine=affine.copy()),
                             prob=tio.ScalarImage(prob.to(device), affine=affine.copy()))


def _check_samplers(device):
    import torchio_b200 as tio

    gold = json.loads((GOLDEN / "patches_samplers.json").read_text())
    case, subject = _sampler_subject(device)
    size = case["patch_size"]
    torch.manual_seed(5)
    ws = tio.WeightedSampler(subject, patch_size=size, probability_map="prob")
    got = [[list(p.patch_location.index), float(p.t1.data.double().sum())] for p in ws.sample(subject, 6)]
    assert [g[0] for g in got] == [w[0] for w in gold["weighted"]]
    assert [g[1] for g in got] == pytest.approx([w[1] for w in gold["weighted"]], rel=1e-12)
    torch.manual_seed(5)
    assert [list(p.patch_location.index) for p in ws(subject, 6)] == [w[0] for w in gold["weighted"]]
    torch.manual_seed(6)
    ls = tio.LabelSampler(subject, patch_size=size, label_name="seg", label_probabilities={1: 1.0, 3: 2.0})
    got = [[list(p.patch_location.index), float(p.seg.data.double().sum())] for p in ls(subject, 6)]
    assert got == gold["label_probs"]
    torch.manual_seed(7)
    ls2 = tio.LabelSampler(subject, patch_size=size, label_name="seg")
    assert [list(p.patch_location.index) for p in ls2.sample(subject, 4)] == gold["label_default"]
    gs = tio.GridSampler(subject, patch_size=size, patch_overlap=(2, 4, 4))
    assert len(gs) == len(gold["grid"])
    for i, (index, total) in enumerate(gold["grid"]):
        assert list(gs.locations[i].index) == index
        assert float(gs[i].t1.data.double().sum()) == pytest.approx(total, rel=1e-12)
    return gold, subject, size


def test_weighted_label_grid_samplers_match_reference_on_host():
    _check_samplers("cpu")
    with pytest.raises(RuntimeError):
        import torchio_b200 as tio

        _, subject = _sampler_subject()
        zero = tio.Subject(t1=subject.t1, prob=tio.ScalarImage(torch.zeros_like(subject.prob.data)))
        next(iter(tio.WeightedSampler(zero, patch_size=4, probability_map="prob")(zero, 1)))


@pytest.mark.gpu
def test_weighted_label_grid_samplers_match_reference_on_device():
    """Device-resident subject: same draws (the map is sampled on the host), patches gathered
    by tio_crop_patches; the padded grid goes through the Pad kernel."""
    import torchio_b200 as tio

    gold, subject, size = _check_samplers("cuda")
    gp = tio.GridSampler(subject, patch_size=size, patch_overlap=(2, 4, 4), padding_mode="reflect")
    assert list(gp.subject.spatial_shape) == gold["grid_padded_shape"]
    for i, (index, total) in enumerate(gold["grid_padded"]):
        assert list(gp.locations[i].index) == index
        assert float(gp[i].t1.data.double().sum()) == pytest.approx(total, rel=1e-12)


def test_queue_honours_the_samplers_own_limit_and_overridden_call():
    """`islice(sampler(subject), patches_per_volume)` semantics (data/queue.py:140-146): the
    sampler's ``num_patches`` caps the count, and a subclass's ``__call__`` is what runs."""
    import torchio_b200 as tio

    case = PATCH_CASES[1]
    subjects = _subjects(case)
    capped = tio.UniformSampler(subjects[0], patch_size=case["patch_size"], num_patches=2)
    q = tio.Queue(subjects, capped, max_length=50, patches_per_volume=5, shuffle_subjects=False,
                  shuffle_patches=False)
    torch.manual_seed(0)
    assert len(list(q)) == 2 * len(subjects)

    class FirstCorner(tio.UniformSampler):
        def __call__(self, subject, num_patches=None):
            while True:
                yield self._extract_patch(subject, tio.PatchLocation(index=(0, 0, 0), size=self.patch_size))

    q = tio.Queue(subjects, FirstCorner(subjects[0], patch_size=case["patch_size"]), max_length=50,
                  patches_per_volume=3, shuffle_subjects=False, shuffle_patches=False)
    got = list(q)
    assert len(got) == 3 * len(subjects) and all(p.patch_location.index == (0, 0, 0) for p in got)
    # a patch larger than the volume is the reference's clamped view, not an error
    big = tio.UniformSampler(subjects[0], patch_size=(64, 8, 8))
    torch.manual_seed(0)
    patch = next(iter(tio.Queue(subjects, big, patches_per_volume=1, shuffle_subjects=False)))
    assert tuple(patch.t1.data.shape[1:]) == (case["shape"][0], 8, 8)


@pytest.mark.gpu
def test_device_queue_uses_the_patch_ring_and_collates_with_one_gather():
    """Device subjects: patches live in ring slots (no per-patch tensors), the loader's batches come
    from one index_select per image, slots are reused after a flush, handles behave like Subjects."""
    import torchio_b200 as tio
    from torchio_b200.patches import PatchHandle

    case = PATCH_CASES[0]
    subjects = _subjects(case, "cuda")
    queue = _queue(case, subjects)
    torch.manual_seed(case["seed"])
    random.seed(case["seed"])
    handles = list(queue)
    assert all(isinstance(h, PatchHandle) for h in handles)
    ring = handles[-1].ring
    assert ring.capacity == case["max_length"] + case["patches_per_volume"] - 1
    attached = [h for h in handles if h.ring is ring]
    detached = [h for h in handles if h.ring is None]  # held across a refill: they own a copy now
    assert attached and detached and len(attached) + len(detached) == len(handles)
    assert max(h.slot for h in handles) < ring.capacity
    pi, pj, pk = case["patch_size"]
    for h in (handles[0], handles[-1]):  # a detached one and one still in the ring
        sid, (i, j, k) = int(h.sid), h.patch_location.index
        assert torch.equal(h.t1.data, subjects[sid].t1.data[:, i:i + pi, j:j + pj, k:k + pk])
    last = handles[-1]
    assert last.t1.data.data_ptr() == ring.data["t1"][last.slot].data_ptr()  # a view, not a copy
    # collate: same batches as stacking materialised patches (each epoch fully consumed under its
    # own seeding: the two queues draw from the same global generators)
    def epoch(q):
        torch.manual_seed(case["seed"])
        random.seed(case["seed"])
        out = []
        for b in tio.SubjectsLoader(q, batch_size=case["batch_size"]):
            out.append((b.t1.data.cpu(), b.seg.data.cpu(), [l.index for l in b.metadata["patch_location"]],
                        list(b.metadata["sid"]), [a.numpy() for a in b.t1.affines]))
        return out

    got_batches, want_batches = epoch(queue), epoch(_queue(case, _subjects(case, "cpu")))
    assert len(got_batches) == len(want_batches)
    n = 0
    for got, want in zip(got_batches, want_batches):
        assert torch.equal(got[0], want[0]) and torch.equal(got[1], want[1])
        assert got[2] == want[2] and got[3] == want[3]
        for a, b in zip(got[4], want[4]):
            assert np.allclose(a, b, rtol=0, atol=1e-12)
        n += got[0].shape[0]
    assert n == queue.patches_per_epoch
