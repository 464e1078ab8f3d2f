"""Host logic of the multi-GPU path on CPU: contiguous batch sharding and the
optional all-gather of the outputs, world_size 2 and 3 over gloo.  The per-image seeds make a
sharded run see exactly the images of the unsharded run."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from polyblur_b200 import sharding, synthetic


def test_shard_range_partitions():
    for n in (1, 7, 32, 33, 256):
        for world in (1, 2, 3, 8):
            spans = [sharding.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharding.shard_range(4, 2, 2)


def test_synthetic_images_are_shard_invariant():
    full = synthetic.make("mosaic", 5, 3, 40, 56)
    for world in (2, 3):
        parts = []
        for r in range(world):
            a, b = sharding.shard_range(5, r, world)
            parts.append(synthetic.make("mosaic", b - a, 3, 40, 56, first_index=a))
        assert torch.equal(torch.cat(parts), full)
    assert not torch.equal(synthetic.make("white", 1, 3, 8, 8), synthetic.make("white", 1, 3, 8, 8, first_index=1))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, n_items, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        a, b = sharding.shard_range(n_items, rank, world)
        # stand-in for the per-rank deblurred slice: a deterministic function of the global index
        local = torch.stack([torch.full((2, 3, 4), float(i)) for i in range(a, b)]) if b > a else torch.zeros(0, 2, 3, 4)
        out = sharding.gather_outputs(local, n_items)
        ok = out.shape[0] == n_items and all(float(out[i, 0, 0, 0]) == float(i) for i in range(n_items))
        # max-over-ranks timing reduction used by bench.py
        t = torch.tensor([float(rank + 1)], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ok = ok and float(t) == float(world)
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n_items", [(2, 5), (3, 7), (2, 2)])
def test_gather_outputs_gloo(world, n_items):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_items, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(results) == [(r, True) for r in range(world)]


def test_numa_binding_is_a_no_op_without_topology():
    """bind_to_gpu_numa_node never raises: without a CUDA device (or on a single-node VM, numa_node = -1) it returns
    None and leaves the process affinity alone."""
    import os

    from polyblur_b200 import sharding
    before = os.sched_getaffinity(0)
    assert sharding.bind_to_gpu_numa_node(0) is None or isinstance(sharding.bind_to_gpu_numa_node(0), dict)
    if not torch.cuda.is_available():
        assert os.sched_getaffinity(0) == before
