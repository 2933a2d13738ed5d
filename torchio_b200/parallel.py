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


def gather_buffers(batch: SubjectsBatch, counts: list[int]) -> dict[str, torch.Tensor]:
    """Root-side destination of `gather_batch_to_root`: ``{name: (sum(counts), C, I, J, K)}`` on
    the batch's device, allocated once and reused across gathers."""
    total = int(sum(counts))
    return {name: torch.empty((total, *ib.data.shape[1:]), dtype=ib.data.dtype, device=ib.data.device)
            for name, ib in batch.images.items()}


def gather_batch_to_root(batch: SubjectsBatch, *, root: int = 0, counts: list[int] | None = None,
                         out: dict[str, torch.Tensor] | None = None) -> dict[str, torch.Tensor] | None:
    """Gather every image tensor of the rank-local batches to ``root``.

    Returns ``{name: (sum_B, C, I, J, K) tensor}`` on root (rank order), None elsewhere.  Ranks may
    hold different batch sizes: pass ``counts`` (per-rank batch sizes) when they are known, else
    they are exchanged with one small all_gather.  ``out`` (root only, from `gather_buffers`)
    receives the data in place.  NCCL: every peer's block is received straight into its slice of
    the destination with one grouped batch of point-to-point operations, so the seven inbound
    NVLink transfers of an 8-GPU box run concurrently into rank 0; gloo (CPU tests): the same
    exchange on CPU tensors."""
    world, rank = dist.get_world_size(), dist.get_rank()
    names = list(batch.images)
    first = batch.images[names[0]].data
    if counts is None:
        sizes = [torch.zeros(1, dtype=torch.int64, device=first.device) for _ in range(world)]
        dist.all_gather(sizes, torch.tensor([first.shape[0]], dtype=torch.int64, device=first.device))
        counts = [int(s.item()) for s in sizes]
    offsets = [0]
    for c in counts:
        offsets.append(offsets[-1] + int(c))
    if rank == root and out is None:
        out = gather_buffers(batch, counts)
    requests = []
    for name in names:
        # as bytes: NCCL's point-to-point has no int16 ("Short") and the payload is opaque anyway
        data = batch.images[name].data.contiguous().view(torch.uint8)
        if rank == root:
            dest = out[name].view(torch.uint8)
            for src in range(world):
                block = dest[offsets[src]:offsets[src + 1]]
                if src == root:
                    block.copy_(data, non_blocking=True)
                elif counts[src]:
                    requests.append(dist.P2POp(dist.irecv, block, src))
        elif counts[rank]:
            requests.append(dist.P2POp(dist.isend, data, root))
    if requests:
        for work in dist.batch_isend_irecv(requests):
            work.wait()
    return out if rank == root else None


def bind_to_gpu_numa(local_rank: int) -> dict | None:
    """Pin the calling process to the CPU cores of the NUMA node its GPU hangs off, so that the
    rank's Python threads and the pinned staging buffers it allocates afterwards (first touch) sit
    next to the GPU's PCIe root.  Returns what was done ({"node": n, "cpus": k}) or None when the
    topology cannot be read (then nothing changes)."""
    import os
    import subprocess

    try:
        bus = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", str(local_rank)],
                             capture_output=True, text=True, timeout=20).stdout.strip().lower()
        if not bus:
            return None
        if bus.count(":") == 2 and len(bus.split(":")[0]) == 8:  # 00000000:1B:00.0 -> 0000:1b:00.0
            bus = bus[4:]
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read().strip())
        if node < 0:
            return None
        cpus: set[int] = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = cpus & os.sched_getaffinity(0)
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return {"node": node, "cpus": len(allowed), "gpu_bus": bus, "mempolicy": _prefer_memory_node(node)}
    except (OSError, ValueError, subprocess.SubprocessError):
        return None


def _prefer_memory_node(node: int) -> str | None:
    """`set_mempolicy(MPOL_PREFERRED, node)` for the calling thread (inherited by the threads it
    starts): pages allocated from now on — the page-locked staging buffers in particular — come
    from the GPU's NUMA node even when the allocating thread is not the one bound above.  Soft
    preference (falls back to other nodes when the node is full).  Returns "preferred" or None when
    the call is unavailable (then placement is first-touch, as before)."""
    import ctypes
    import platform

    number = {"x86_64": 238, "aarch64": 237}.get(platform.machine())
    if number is None or not 0 <= node < 64:
        return None
    try:
        syscall = ctypes.CDLL(None, use_errno=True).syscall
        syscall.restype = ctypes.c_long
        syscall.argtypes = [ctypes.c_long, ctypes.c_long, ctypes.c_void_p, ctypes.c_ulong]
        mask = ctypes.c_ulong(1 << node)
        mpol_preferred = 1
        return "preferred" if syscall(number, mpol_preferred, ctypes.addressof(mask), 65) == 0 else None
    except (OSError, AttributeError):
        return None
