"""Patch path: samplers, Queue and loaders (host-side mirror of
data/patch.py, data/sampler.py, data/queue.py and loader.py, TorchIO 2.0.0a2).

Same classes, constructor arguments, iteration order and RNG draws as the
reference (patch corners come from the global torch CPU generator, three
``torch.randint`` calls per patch in i, j, k order — data/sampler.py:218-223;
subject and buffer shuffles use Python's ``random`` — data/queue.py:167,177).

Difference by design: when a subject's tensors live on the GPU, the patches a
`Queue` asks for (``patches_per_volume`` at a time) are gathered by one
`ops.crop_patches` launch per image into a dense block and handed out as views
of that block, instead of one strided view per patch that ``torch.stack``
copies again at collation.  Host-resident subjects keep the reference's
zero-copy views.  ``Queue(device=...)`` (extension) moves each loaded subject to
that device before the transform, so augmentation and patch extraction run
resident.
"""

from __future__ import annotations

import random as _random
from collections import deque
from collections.abc import Iterator, Sequence, Sized
from concurrent.futures import Future, ThreadPoolExecutor
from dataclasses import dataclass
from itertools import islice
from typing import Any

import numpy as np
import torch
from torch.utils.data import DataLoader, Dataset, IterableDataset, Sampler

from . import ops
from .data import ImagesBatch, Subject, SubjectsBatch


@dataclass(frozen=True)
class PatchLocation:
    """Corner index and size of a patch inside its volume (data/patch.py:11-63)."""

    index: tuple[int, int, int]
    size: tuple[int, int, int]
    subject_index: int | None = None

    @property
    def index_ini(self) -> tuple[int, int, int]:
        return self.index

    @property
    def index_fin(self) -> tuple[int, int, int]:
        return (self.index[0] + self.size[0], self.index[1] + self.size[1],
                self.index[2] + self.size[2])

    def to_slices(self) -> tuple[slice, slice, slice]:
        ini, fin = self.index_ini, self.index_fin
        return (slice(ini[0], fin[0]), slice(ini[1], fin[1]), slice(ini[2], fin[2]))

    def scaled(self, factor: tuple[float, float, float]) -> PatchLocation:
        return PatchLocation(
            index=tuple(round(self.index[a] * factor[a]) for a in range(3)),
            size=tuple(round(self.size[a] * factor[a]) for a in range(3)),
            subject_index=self.subject_index,
        )


class PatchSampler:
    """Base class of patch samplers (data/sampler.py:25-67)."""

    def __init__(self, patch_size) -> None:
        if isinstance(patch_size, int):
            patch_size = (patch_size, patch_size, patch_size)
        self.patch_size = tuple(int(v) for v in patch_size)

    def __call__(self, subject: Subject, num_patches: int | None = None) -> Iterator[Subject]:
        raise NotImplementedError(f"{type(self).__name__} must implement __call__")

    def _extract_patch(self, subject: Subject, location: PatchLocation) -> Subject:
        si, sj, sk = location.to_slices()
        kwargs: dict[str, Any] = {name: image[:, si, sj, sk] for name, image in subject.images.items()}
        kwargs.update(subject.metadata)
        kwargs["patch_location"] = location
        return Subject(**kwargs)

    def _extract_patches(self, subject: Subject, locations: list[PatchLocation]) -> list[Subject]:
        """All ``locations`` of one subject.  Device-resident images: one gather launch
        per image; host-resident: the reference's views."""
        if not locations:
            return []
        images = subject.images
        on_device = all(img.data.is_cuda for img in images.values())
        same_size = all(loc.size == locations[0].size for loc in locations)
        if not (on_device and same_size):
            return [self._extract_patch(subject, loc) for loc in locations]
        corners = np.asarray([loc.index for loc in locations], dtype=np.int32)
        blocks = {name: ops.crop_patches(img.data, corners, locations[0].size) for name, img in images.items()}
        patches = []
        for row, loc in enumerate(locations):
            kwargs: dict[str, Any] = {}
            for name, img in images.items():
                matrix = img.affine.numpy().copy()
                matrix[:3, 3] = matrix[:3, 3] + matrix[:3, :3] @ np.asarray(loc.index, dtype=np.float64)
                kwargs[name] = type(img)(blocks[name][row], affine=matrix, **img.metadata)
            kwargs.update(subject.metadata)
            kwargs["patch_location"] = loc
            patches.append(Subject(**kwargs))
        return patches


class UniformSampler(PatchSampler, IterableDataset):
    """Uniformly random patches (data/sampler.py:165-223)."""

    def __init__(self, subject: Subject, patch_size, num_patches: int | None = None) -> None:
        super().__init__(patch_size)
        self.subject = subject
        self.num_patches = num_patches

    def __call__(self, subject: Subject, num_patches: int | None = None) -> Iterator[Subject]:
        limit = num_patches or self.num_patches
        count = 0
        while limit is None or count < limit:
            loc = PatchLocation(index=self._random_index(subject.spatial_shape), size=self.patch_size)
            yield self._extract_patch(subject, loc)
            count += 1

    def __iter__(self) -> Iterator[Subject]:
        return self(self.subject, self.num_patches)

    def sample(self, subject: Subject, num_patches: int) -> list[Subject]:
        """``list(islice(self(subject), num_patches))`` with the same RNG draws, the
        patches of a device-resident subject gathered in one launch per image."""
        shape = subject.spatial_shape
        locations = [PatchLocation(index=self._random_index(shape), size=self.patch_size)
                     for _ in range(num_patches)]
        return self._extract_patches(subject, locations)

    def _random_index(self, spatial_shape) -> tuple[int, int, int]:
        def _rand(d: int) -> int:
            hi = max(spatial_shape[d] - self.patch_size[d], 0) + 1
            return int(torch.randint(0, hi, (1,)).item())

        return (_rand(0), _rand(1), _rand(2))


class GridSampler(PatchSampler, Dataset):
    """Patches on a regular grid for dense inference (data/sampler.py:70-162)."""

    def __init__(self, subject: Subject, patch_size, patch_overlap=0, padding_mode: str | None = None,
                 fill: float = 0) -> None:
        super().__init__(patch_size)
        if isinstance(patch_overlap, int):
            patch_overlap = (patch_overlap, patch_overlap, patch_overlap)
        self.patch_overlap = tuple(int(v) for v in patch_overlap)
        self.padding_mode = padding_mode
        self.fill = fill
        self.subject = self._maybe_pad(subject)
        self.locations = self._compute_locations(self.subject.spatial_shape)

    def __len__(self) -> int:
        return len(self.locations)

    def __getitem__(self, index: int) -> Subject:
        return self._extract_patch(self.subject, self.locations[index])

    def _maybe_pad(self, subject: Subject) -> Subject:
        if self.padding_mode is None:
            return subject
        from .transforms.neighbours import Pad

        border = tuple(v // 2 for v in self.patch_overlap)
        padding = (border[0], border[0], border[1], border[1], border[2], border[2])
        return Pad(padding=padding, padding_mode=self.padding_mode, fill=self.fill, copy=False)(subject)

    def _compute_locations(self, spatial_shape) -> list[PatchLocation]:
        per_axis: list[list[int]] = []
        for dim in range(3):
            size, patch, overlap = spatial_shape[dim], self.patch_size[dim], self.patch_overlap[dim]
            step = max(patch - overlap, 1)
            indices = list(range(0, size - patch + 1, step))
            if not indices or indices[-1] != size - patch:
                indices.append(max(size - patch, 0))
            per_axis.append(indices)
        return [PatchLocation(index=(i, j, k), size=self.patch_size)
                for i in per_axis[0] for j in per_axis[1] for k in per_axis[2]]


def _mask_borders(prob: torch.Tensor, spatial_shape, patch_size) -> torch.Tensor:
    """Zero probability where a patch centre cannot be placed (data/sampler.py:340-360)."""
    prob = prob.clone()
    for d in range(3):
        half = patch_size[d] // 2
        if half > 0:
            lo: list[slice] = [slice(None)] * 3
            lo[d] = slice(0, half)
            prob[tuple(lo)] = 0
        tail = spatial_shape[d] - half
        if tail < spatial_shape[d]:
            hi: list[slice] = [slice(None)] * 3
            hi[d] = slice(tail, None)
            prob[tuple(hi)] = 0
    return prob


def _center_to_corner(center, spatial_shape, patch_size) -> tuple[int, int, int]:
    """Centre voxel -> patch corner, clamped into the volume (data/sampler.py:363-375)."""
    result = []
    for d in range(3):
        corner = max(0, center[d] - patch_size[d] // 2)
        result.append(min(corner, spatial_shape[d] - patch_size[d]))
    return (result[0], result[1], result[2])


class WeightedSampler(PatchSampler, IterableDataset):
    """Random patches weighted by a probability map (data/sampler.py:226-283).

    The centre voxels are drawn with ``torch.multinomial`` on the flattened map.  For a
    device-resident subject the map is brought to the host once per subject and drawn there,
    so the draws are those of the reference (global CPU generator), whatever the device."""

    def __init__(self, subject: Subject, patch_size, probability_map: str, num_patches: int | None = None) -> None:
        super().__init__(patch_size)
        self.subject = subject
        self.probability_map = probability_map
        self.num_patches = num_patches

    def _flat_map(self, subject: Subject):
        prob = self._build_probability_map_for(subject)
        flat = prob.flatten()
        if flat.sum() == 0:
            raise RuntimeError(f"Probability map '{self.probability_map}' is all zeros")
        return flat.cpu(), tuple(prob.shape)

    def _draw(self, flat, shape, subject: Subject) -> PatchLocation:
        idx_flat = torch.multinomial(flat, 1).item()
        center = tuple(int(x) for x in np.unravel_index(int(idx_flat), shape))
        return PatchLocation(index=_center_to_corner(center, subject.spatial_shape, self.patch_size),
                             size=self.patch_size)

    def __call__(self, subject: Subject, num_patches: int | None = None) -> Iterator[Subject]:
        flat, shape = self._flat_map(subject)
        limit = num_patches or self.num_patches
        count = 0
        while limit is None or count < limit:
            yield self._extract_patch(subject, self._draw(flat, shape, subject))
            count += 1

    def __iter__(self) -> Iterator[Subject]:
        return self(self.subject, self.num_patches)

    def sample(self, subject: Subject, num_patches: int) -> list[Subject]:
        """``list(islice(self(subject), num_patches))``, same draws, one gather launch per
        image for a device-resident subject."""
        flat, shape = self._flat_map(subject)
        return self._extract_patches(subject, [self._draw(flat, shape, subject) for _ in range(num_patches)])

    def _build_probability_map_for(self, subject: Subject) -> torch.Tensor:
        prob_data = subject.images[self.probability_map].data[0].float()
        return _mask_borders(prob_data, subject.spatial_shape, self.patch_size)

    def _build_probability_map(self) -> torch.Tensor:
        return self._build_probability_map_for(self.subject)


class LabelSampler(WeightedSampler):
    """Random patches centred on labelled voxels (data/sampler.py:286-333)."""

    def __init__(self, subject: Subject, patch_size, label_name: str,
                 label_probabilities: dict[int, float] | None = None, num_patches: int | None = None) -> None:
        super().__init__(subject, patch_size, probability_map=label_name, num_patches=num_patches)
        self.label_name = label_name
        self.label_probabilities = label_probabilities

    def _build_probability_map_for(self, subject: Subject) -> torch.Tensor:
        label_data = subject.images[self.label_name].data[0]
        if self.label_probabilities is not None:
            prob = torch.zeros_like(label_data, dtype=torch.float32)
            for label, weight in self.label_probabilities.items():
                prob[label_data == label] = weight
        else:
            prob = (label_data > 0).float()
        return _mask_borders(prob, subject.spatial_shape, self.patch_size)


class Queue(IterableDataset):
    """Patch buffer for stochastic patch-based training (data/queue.py:21-208)."""

    def __init__(self, subjects: Sequence[Subject], patch_sampler: PatchSampler, max_length: int = 300,
                 patches_per_volume: int = 10, num_workers: int = 0, shuffle_subjects: bool = True,
                 shuffle_patches: bool = True, transform: Any | None = None,
                 subject_sampler: Sampler | None = None, device: str | torch.device | None = None) -> None:
        if subject_sampler is not None and shuffle_subjects:
            raise ValueError(
                "shuffle_subjects must be False when subject_sampler"
                " is provided (the sampler controls the order)"
            )
        self.subjects = subjects
        self.patch_sampler = patch_sampler
        self.max_length = max_length
        self.patches_per_volume = patches_per_volume
        self.num_workers = num_workers
        self.shuffle_subjects = shuffle_subjects
        self.shuffle_patches = shuffle_patches
        self.transform = transform
        self.subject_sampler = subject_sampler
        self.device = None if device is None else torch.device(device)

    def __iter__(self) -> Iterator[Subject]:
        buffer: list[Subject] = []
        subject_iter = self._make_subject_iter()
        if self.num_workers > 0:
            yield from self._iter_threaded(subject_iter, buffer)
        else:
            yield from self._iter_sync(subject_iter, buffer)

    def _iter_sync(self, subject_iter, buffer) -> Iterator[Subject]:
        for raw in subject_iter:
            prepared = self._prepare(raw)
            buffer.extend(self._sample_patches(prepared))
            yield from self._drain_if_full(buffer)
        yield from self._flush(buffer)

    def _iter_threaded(self, subject_iter, buffer) -> Iterator[Subject]:
        with ThreadPoolExecutor(max_workers=self.num_workers) as pool:
            futures: deque[Future[Subject]] = deque()
            for raw in subject_iter:
                futures.append(pool.submit(self._prepare, raw))
                yield from self._collect_ready(futures, buffer)
                yield from self._drain_if_full(buffer)
            for future in futures:
                buffer.extend(self._sample_patches(future.result()))
        yield from self._flush(buffer)

    def _collect_ready(self, futures, buffer) -> Iterator[Subject]:
        while futures and futures[0].done():
            buffer.extend(self._sample_patches(futures.popleft().result()))
        return iter(())

    def _drain_if_full(self, buffer) -> Iterator[Subject]:
        if len(buffer) >= self.max_length:
            yield from self._flush(buffer)

    def _flush(self, buffer) -> Iterator[Subject]:
        if self.shuffle_patches:
            _random.shuffle(buffer)
        while buffer:
            yield buffer.pop()

    def _prepare(self, subject: Subject) -> Subject:
        subject.load()
        if self.device is not None:  # a moved copy: the dataset's own subject stays where it is
            kwargs: dict[str, Any] = {
                name: img.new_like(data=img.data.to(self.device, non_blocking=True))
                for name, img in subject.images.items()}
            kwargs.update(subject.metadata)
            moved = Subject(**kwargs)
            moved.applied_transforms = list(subject.applied_transforms)
            subject = moved
        if self.transform is not None:
            subject = self.transform(subject)
        return subject

    def _sample_patches(self, subject: Subject) -> list[Subject]:
        batched = getattr(self.patch_sampler, "sample", None)
        if batched is not None:
            return batched(subject, self.patches_per_volume)
        return list(islice(iter(self.patch_sampler(subject)), self.patches_per_volume))

    def _make_subject_iter(self) -> Iterator[Subject]:
        if self.subject_sampler is not None:
            return (self.subjects[i] for i in list(self.subject_sampler))
        subjects = list(self.subjects)
        if self.shuffle_subjects:
            _random.shuffle(subjects)
        return iter(subjects)

    @property
    def num_subjects(self) -> int:
        sampler = self.subject_sampler
        if sampler is not None:
            if not isinstance(sampler, Sized):
                raise TypeError("subject_sampler must have a __len__ method")
            return len(sampler)
        return len(self.subjects)

    @property
    def patches_per_epoch(self) -> int:
        return self.num_subjects * self.patches_per_volume

    @property
    def max_memory(self) -> int:
        sample = self.subjects[0]
        channels = sum(img.num_channels for img in sample.images.values())
        voxels = 1
        for s in self.patch_sampler.patch_size:
            voxels *= s
        return 4 * channels * voxels * self.max_length

    @property
    def max_memory_pretty(self) -> str:
        value = float(self.max_memory)
        for unit in ("Bytes", "KiB", "MiB", "GiB", "TiB"):
            if value < 1024 or unit == "TiB":
                return f"{value:.0f} {unit}" if unit == "Bytes" else f"{value:.1f} {unit}"
            value /= 1024
        return f"{value:.1f} TiB"


def collate_subjects(batch: Sequence[Any]) -> SubjectsBatch:
    """List of Subjects -> SubjectsBatch with stacked 5-D tensors (loader.py:15-24)."""
    return SubjectsBatch.from_subjects(list(batch))


def collate_images(batch: Sequence[Any]) -> ImagesBatch:
    return ImagesBatch.from_images(list(batch))


class SubjectsLoader(DataLoader):
    """DataLoader that returns `SubjectsBatch` instances (loader.py:38-65)."""

    def __init__(self, dataset: Dataset, **kwargs: Any) -> None:
        if "collate_fn" in kwargs:
            raise ValueError(
                "SubjectsLoader sets collate_fn automatically; "
                "pass a plain DataLoader if you need a custom collate_fn"
            )
        super().__init__(dataset, collate_fn=collate_subjects, **kwargs)


class ImagesLoader(DataLoader):
    """DataLoader that returns `ImagesBatch` instances (loader.py:68-90)."""

    def __init__(self, dataset: Dataset, **kwargs: Any) -> None:
        if "collate_fn" in kwargs:
            raise ValueError(
                "ImagesLoader sets collate_fn automatically; "
                "pass a plain DataLoader if you need a custom collate_fn"
            )
        super().__init__(dataset, collate_fn=collate_images, **kwargs)


StudiesLoader = SubjectsLoader
collate_studies = collate_subjects
