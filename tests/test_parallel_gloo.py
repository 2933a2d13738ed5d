"""N>1 host path on CPU: world_size-2 gloo processes shard a batch, transform
their block with a deterministic stand-in for the kernels, and gather to rank 0.
The result must equal the unsharded run.  (The CUDA kernels themselves are
covered by the -m gpu tests; this pins sharding, seeding and the gather.)"""

import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _subjects(n):
    import torchio_b200 as tio

    return [tio.Subject(t1=tio.ScalarImage(torch.full((1, 4, 5, 6), float(i))), idx=i)
            for i in range(n)]


def _worker(rank, world, port, n, ragged, queue):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import torchio_b200 as tio
        from torchio_b200 import parallel

        mine = parallel.shard_subjects(_subjects(n))
        seed = parallel.seed_for_rank(100)
        assert seed == 100 + rank
        batch = tio.SubjectsBatch.from_subjects(mine)
        # stand-in "augmentation": params drawn from the rank's own stream
        shift = torch.rand(1).item()
        for ib in batch.images.values():
            ib.data = ib.data * 2 + 1
        out = parallel.gather_batch_to_root(batch)
        if rank == 0:
            queue.put((out["t1"].clone(), shift))
        else:
            assert out is None
            queue.put(("shift", rank, shift))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n", [4, 5])
def test_shard_transform_gather_world2(n):
    from torchio_b200 import parallel

    assert list(parallel.shard_range(5, 0, 2)) == [0, 1, 2]
    assert list(parallel.shard_range(5, 1, 2)) == [3, 4]
    ctx = mp.get_context("spawn")
    queue = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, n % 2, queue)) for r in range(2)]
    for p in procs:
        p.start()
    results = [queue.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    gathered = next(r for r in results if r[0] != "shift" if not isinstance(r[0], str))
    other = next(r for r in results if isinstance(r[0], str))
    data, shift0 = gathered
    assert data.shape == (n, 1, 4, 5, 6)
    expected = torch.stack([torch.full((1, 4, 5, 6), float(i)) * 2 + 1 for i in range(n)])
    assert torch.equal(data, expected)
    assert shift0 != other[2]  # ranks draw from different streams
