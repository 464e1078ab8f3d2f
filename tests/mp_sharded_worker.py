"""Worker of tests/test_gpu_round2.py::test_two_rank_gathered_output_is_bitwise_the_single_gpu_output.
Launched by torch.distributed.run with 2 (or more) ranks: rank r deblurs shard_range(B, r, world) of a seeded
batch on its GPU (cuda:LOCAL_RANK when the box has that many GPUs, else every rank on cuda:0), the slices are
gathered with sharding.gather_outputs and rank 0 compares with the whole batch run by one call."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import polyblur_b200 as pb  # noqa: E402
from polyblur_b200 import sharding, synthetic  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    ngpu = torch.cuda.device_count()
    assert ngpu >= 1, "needs a CUDA device"
    multi = ngpu >= world
    dev = torch.device("cuda", local if multi else 0)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl" if multi else "gloo", rank=rank, world_size=world)
    try:
        B, shape = 5, (3, 270, 480)                 # ragged: 3 + 2 images
        kw = dict(n_iter=3, alpha=6, beta=1)
        for kind in ("mosaic", "white"):
            a, b = sharding.shard_range(B, rank, world)
            mine = synthetic.make(kind, b - a, *shape, first_index=a).to(dev)
            out = pb.polyblur_deblurring(mine, **kw)
            full = sharding.gather_outputs(out if multi else out.cpu(), B)
            if rank == 0:
                whole = pb.polyblur_deblurring(synthetic.make(kind, B, *shape).to(dev), **kw)
                assert full.shape == whole.shape
                assert torch.equal(full.to(dev), whole), (kind, float((full.to(dev) - whole).abs().max()))
        dist.barrier()
        if rank == 0:
            print(f"SHARDED_BITWISE_OK world={world} backend={'nccl' if multi else 'gloo'} gpus={ngpu}", flush=True)
    finally:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
