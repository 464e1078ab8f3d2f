"""Where the end-to-end time of polyblur_deblurring(pinned CPU tensor) goes (tuning aid)."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

import torch

import polyblur_b200
from polyblur_b200 import deblurring, synthetic


def wall(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    x = synthetic.make("mosaic", B, 3, 1080, 1920, device="cuda")
    xh = torch.empty(x.shape, dtype=torch.float32, pin_memory=True)
    xh.copy_(x)
    res = {}
    res["pinned_alloc_ms"] = wall(lambda: torch.empty(x.shape, dtype=torch.float32, pin_memory=True))
    res["device_only_ms"] = wall(lambda: polyblur_b200.polyblur_deblurring(x, n_iter=3, alpha=6, beta=1))
    res["api_ms"] = wall(lambda: polyblur_b200.polyblur_deblurring(xh, n_iter=3, alpha=6, beta=1))
    for mc in (2, 4, 8, 16, 32):
        import functools
        orig = deblurring._polyblur_host_pipelined
        deblurring._polyblur_host_pipelined = functools.partial(orig, max_chunks=mc)
        try:
            res[f"api_chunks{mc}_ms"] = wall(lambda: polyblur_b200.polyblur_deblurring(xh, n_iter=3, alpha=6, beta=1))
        finally:
            deblurring._polyblur_host_pipelined = orig
    # plain copies of the same tensors
    yh = torch.empty(x.shape, dtype=torch.float32, pin_memory=True)
    res["h2d_ms"] = wall(lambda: x.copy_(xh, non_blocking=True))
    res["d2h_ms"] = wall(lambda: yh.copy_(x, non_blocking=True))
    print(json.dumps({k: round(v, 2) for k, v in res.items()}))


if __name__ == "__main__":
    main()
