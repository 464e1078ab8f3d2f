"""Pinned host<->device copy bandwidth of the box (ceiling of the end-to-end number in bench.py)."""
import json
import time

import torch


def main():
    n = 796 * 1024 * 1024
    h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
    d_a = torch.empty(n, dtype=torch.uint8, device="cuda")
    d_b = torch.empty(n, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    res = {}

    def timed(name, fn, reps=5):
        fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / reps
        res[name] = {"ms": round(dt * 1e3, 2), "GB/s": round(n / dt / 1e9, 2)}

    timed("h2d", lambda: d_a.copy_(h_in, non_blocking=True))
    timed("d2h", lambda: h_out.copy_(d_b, non_blocking=True))

    def both():
        with torch.cuda.stream(s1):
            d_a.copy_(h_in, non_blocking=True)
        with torch.cuda.stream(s2):
            h_out.copy_(d_b, non_blocking=True)

    timed("h2d+d2h concurrent (per direction)", both)
    # pageable for comparison (what a caller with plain numpy arrays pays)
    p_in = torch.empty(n, dtype=torch.uint8)
    timed("h2d pageable", lambda: d_a.copy_(p_in), reps=2)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
