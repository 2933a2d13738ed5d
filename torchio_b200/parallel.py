"""Multi-GPU plumbing: one process per GPU, batch-dimension sharding.

The augmentation path shards by independent volumes: every
rank transforms its own contiguous block of subjects, and nothing crosses
GPUs on the data path.  The only exchange the north-star names is the final
gather of augmented volumes/patches to rank 0, done with `torch.distributed`
(NCCL on GPUs, gloo in the CPU tests).

Seeding: rank r draws its parameters from ``base_seed + r`` so ranks are
independent and a run is reproducible for a fixed world size.  (The
reference has no process-group code; `Queue(subject_sampler=
DistributedSampler(...))` is its only hook, data/queue.py:49,169-176.)
"""

from __future__ import annotations

import torch
import torch.distributed as dist

from .data import SubjectsBatch


def shard_range(total: int, rank: int, world: int) -> range:
    """Contiguous block of ``total`` items owned by ``rank`` (sizes differ by <= 1)."""
    if not 0 <= rank < world:
        raise ValueError(f"rank {rank} outside world of size {world}")
    base, extra = divmod(total, world)
    start = rank * base + min(rank, extra)
    return range(start, start + base + (1 if rank < extra else 0))


def shard_subjects(subjects: list, rank: int | None = None, world: int | None = None) -> list:
    rank = dist.get_rank() if rank is None else rank
    world = dist.get_world_size() if world is None else world
    block = shard_range(len(subjects), rank, world)
    return [subjects[i] for i in block]


def seed_for_rank(base_seed: int, rank: int | None = None) -> int:
    rank = dist.get_rank() if rank is None else rank
    seed = int(base_seed) + int(rank)
    torch.manual_seed(seed)
    return seed


def gather_batch_to_root(batch: SubjectsBatch, *, root: int = 0) -> dict[str, torch.Tensor] | None:
    """Gather every image tensor of the rank-local batches to ``root``.

    Returns ``{name: (sum_B, C, I, J, K) tensor}`` on root (rank order), None
    elsewhere.  Ranks may hold different batch sizes.  With NCCL the tensors
    stay on their GPUs and move over NVLink; with gloo they are CPU tensors."""
    world, rank = dist.get_world_size(), dist.get_rank()
    backend = dist.get_backend()
    out: dict[str, torch.Tensor] = {}
    for name, ib in batch.images.items():
        data = ib.data.contiguous()
        sizes = [torch.zeros(1, dtype=torch.int64, device=data.device) for _ in range(world)]
        dist.all_gather(sizes, torch.tensor([data.shape[0]], dtype=torch.int64, device=data.device))
        counts = [int(s.item()) for s in sizes]
        if backend == "nccl":
            # point-to-point: each rank sends once, root receives in rank order
            if rank == root:
                parts = []
                for src in range(world):
                    if src == root:
                        parts.append(data)
                        continue
                    buf = torch.empty((counts[src], *data.shape[1:]), dtype=data.dtype,
                                      device=data.device)
                    dist.recv(buf, src=src)
                    parts.append(buf)
                out[name] = torch.cat(parts, dim=0)
            else:
                dist.send(data, dst=root)
        else:
            gathered = None
            if rank == root:
                gathered = [torch.empty((counts[r], *data.shape[1:]), dtype=data.dtype)
                            for r in range(world)]
            if len(set(counts)) == 1:
                dist.gather(data, gathered, dst=root)
            else:  # ragged: gloo gather needs equal sizes, fall back to send/recv
                if rank == root:
                    for src in range(world):
                        if src == root:
                            gathered[src] = data
                        else:
                            dist.recv(gathered[src], src=src)
                else:
                    dist.send(data, dst=root)
            if rank == root:
                out[name] = torch.cat(gathered, dim=0)
    return out if rank == root else None
