"""Chunk schedules of the host pipelines, interleaved and repeated (the host link is shared with other
tenants of the box, so single timings are noisy): prints median and min wall ms per variant."""
import functools
import json
import os
import statistics
import sys
import time

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import polyblur_b200  # noqa: E402
from polyblur_b200 import deblurring, io as pbio, synthetic  # noqa: E402


def once(fn):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) * 1e3


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    rounds = int(sys.argv[2]) if len(sys.argv) > 2 else 7
    x = synthetic.make("mosaic", B, 3, 1080, 1920, device="cuda")
    xh = torch.empty(x.shape, dtype=torch.float32, pin_memory=True)
    xh.copy_(x)
    xu = (x.permute(0, 2, 3, 1) * 255).round().to(torch.uint8).contiguous().cpu().pin_memory()
    variants = {}
    for mc, ramp in ((16, ()), (16, (1,)), (32, ()), (32, (1,))):
        def f(mc=mc, ramp=ramp):
            orig = deblurring._polyblur_host_pipelined
            deblurring._polyblur_host_pipelined = functools.partial(orig, max_chunks=mc, ramp=ramp)
            try:
                polyblur_b200.polyblur_deblurring(xh, n_iter=3, alpha=6, beta=1)
            finally:
                deblurring._polyblur_host_pipelined = orig
        variants[f"f32 chunks{mc} ramp{ramp}"] = f
    for mc, ramp in ((4, ()), (4, (2,)), (4, (2, 4)), (4, (1, 2, 4)), (3, (2, 4)), (5, (2, 4)), (6, (2, 4)), (8, (2,))):
        def g(mc=mc, ramp=ramp):
            orig = pbio._host_pipeline_u8
            pbio._host_pipeline_u8 = functools.partial(orig, ramp=ramp)
            try:
                pbio.deblur_uint8(xu, n_iter=3, alpha=6, beta=1, max_chunks=mc)
            finally:
                pbio._host_pipeline_u8 = orig
        variants[f"u8 chunks{mc} ramp{ramp}"] = g
    yh = torch.empty(x.shape, dtype=torch.float32, pin_memory=True)
    s2 = torch.cuda.Stream()

    def both():
        x.copy_(xh, non_blocking=True)
        with torch.cuda.stream(s2):
            yh.copy_(x, non_blocking=True)
    variants["raw h2d+d2h concurrent"] = both
    variants["device only (batch %d)" % B] = lambda: polyblur_b200.polyblur_deblurring(x, n_iter=3, alpha=6, beta=1)
    x8 = x[:8].contiguous()
    variants["device only (batch 8) x4"] = lambda: [polyblur_b200.polyblur_deblurring(x8, n_iter=3, alpha=6, beta=1) for _ in range(4)]
    for f in variants.values():
        f()
    t = {k: [] for k in variants}
    for _ in range(rounds):
        for k, f in variants.items():
            t[k].append(once(f))
    print(json.dumps({k: {"median": round(statistics.median(v), 2), "min": round(min(v), 2)} for k, v in t.items()}, indent=1))


if __name__ == "__main__":
    main()
