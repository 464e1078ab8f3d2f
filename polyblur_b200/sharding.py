"""Batch sharding across the GPUs of one box: images are independent, so
rank r of N owns a contiguous slice of the batch and no collective sits on the data path.
The only (optional) collective is a final all-gather of the outputs."""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world: int):
    """Contiguous [start, stop) of rank `rank`; the first n_items % world ranks get one more."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, extra = divmod(n_items, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def bind_to_gpu_numa_node(device_index: int):
    """Pins the calling process to the CPUs of the NUMA node the GPU hangs off (one process per GPU: its pinned staging
    buffers are then allocated on that node and the host <-> device copies of N ranks do not cross the socket link).
    Returns {"node": n, "cpus": k} or None when the topology is flat, unknown or the affinity cannot be changed."""
    import os
    try:
        pr = torch.cuda.get_device_properties(device_index)
        bdf = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bdf}/numa_node") as f:
            node = int(f.read().strip())
        if node < 0:
            return None
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            spec = f.read().strip()
        cpus = set()
        for part in spec.split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return {"node": node, "cpus": len(cpus)}
    except (OSError, ValueError, AttributeError, RuntimeError):
        return None


def gather_outputs(local: torch.Tensor, n_items: int, group=None) -> torch.Tensor:
    """All-gather the per-rank output slices back into the full batch (NCCL on GPUs, gloo on
    CPU tensors).  Off the timed path; ragged shards are padded to the largest one."""
    world = dist.get_world_size(group)
    if world == 1:
        return local
    sizes = [shard_range(n_items, r, world) for r in range(world)]
    biggest = max(b - a for a, b in sizes)
    padded = local
    if local.shape[0] < biggest:
        pad = torch.zeros((biggest - local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype,
                          device=local.device)
        padded = torch.cat([local, pad], 0)
    parts = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(parts, padded.contiguous(), group=group)
    return torch.cat([p[: b - a] for p, (a, b) in zip(parts, sizes)], 0)


def pipeline_chunks(n_items: int, body: int, ramp=(1,)):
    """Chunk sizes for the host <-> device pipelines (deblurring._polyblur_host_pipelined,
    io._host_pipeline_u8): chunks of about `body` images, with the small `ramp` chunks first and, mirrored,
    last -- the first device->host copy can start after one image has been loaded and processed, and the
    copy left over when the last kernels finish is one image, not a whole chunk."""
    if n_items < 1:
        return []
    body = max(1, int(body))
    head, rem = [], n_items
    for r in ramp:
        if r <= body and rem - 2 * r >= body:
            head.append(int(r))
            rem -= 2 * r
    n_mid = -(-rem // body)
    mid = [rem // n_mid + (1 if i < rem % n_mid else 0) for i in range(n_mid)]
    return head + mid + head[::-1]
