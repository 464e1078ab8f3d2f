"""Concurrent pinned host<->device bandwidth of all ranks (torchrun), with and without binding the
process to the GPU's NUMA node first.  Explains the end-to-end scaling of bench.py at N > 1."""
import json
import os
import sys
import time

import torch
import torch.distributed as dist


def bw(n, dev, bind):
    if bind:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(dev)
        pynvml.nvmlDeviceSetCpuAffinity(h)
    h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
    h_in.fill_(1)
    h_out.fill_(1)
    d_a = torch.empty(n, dtype=torch.uint8, device="cuda")
    d_b = torch.empty(n, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def both():
        with torch.cuda.stream(s1):
            d_a.copy_(h_in, non_blocking=True)
        with torch.cuda.stream(s2):
            h_out.copy_(d_b, non_blocking=True)

    both()
    torch.cuda.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    for _ in range(5):
        both()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 5
    return n / dt / 1e9


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    dev = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(dev)
    n = 796 * 1024 * 1024
    res = {}
    for bind in (False, True):
        v = torch.tensor([bw(n, dev, bind)], dtype=torch.float64)
        out = [torch.zeros(1, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(out, v)
        res["bound" if bind else "unbound"] = [round(float(o), 1) for o in out]
    if rank == 0:
        res["affinity_after_bind"] = len(os.sched_getaffinity(0))
        print(json.dumps(res))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
