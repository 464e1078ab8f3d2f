#!/usr/bin/env python
"""Times the BASELINE.json configurations that are not the bench headline (C1 peacock, C3/C5 4K
shapes, C4 single 12000x9000 image) on one GPU, device-resident, CUDA events, and checks the
size-independent properties (range, batch-composition invariance).  Prints one JSON line each."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import polyblur_b200 as pb  # noqa: E402
from polyblur_b200 import synthetic  # noqa: E402


def timed(fn, warm=2, reps=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def run(name, x, n_iter, **kw):
    out = pb.polyblur_deblurring(x, n_iter=n_iter, alpha=6, beta=1, **kw)
    ms = timed(lambda: pb.polyblur_deblurring(x, n_iter=n_iter, alpha=6, beta=1, **kw))
    B, C, H, W = x.shape
    ok = bool(out.min() >= 0 and out.max() <= 1 and torch.isfinite(out).all())
    solo = pb.polyblur_deblurring(x[-1:].contiguous(), n_iter=n_iter, alpha=6, beta=1, **kw)
    inv = bool(torch.equal(solo[0], out[-1]))
    print(json.dumps({"config": name, "shape": [B, C, H, W], "n_iter": n_iter, "ms": round(ms, 3),
                      "Mpix_s": round(B * H * W / 1e6 / (ms / 1e3), 1), "in_range": ok, "batch_invariant": inv,
                      "gb_s_36B": round(36 * n_iter * B * H * W / (ms / 1e3) / 1e9, 1), **{k: str(v) for k, v in kw.items()}}),
          flush=True)


def main():
    dev = "cuda"
    from PIL import Image
    img = np.asarray(Image.open(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "peacock_defocus.png")))
    x = torch.from_numpy(img.astype(np.float32) / 255).permute(2, 0, 1)[None].contiguous().to(dev)
    run("C1 peacock (device tensor)", x, 3)
    for kind in ("white", "mosaic"):
        run(f"C3/C5 4K batch 8 {kind}", synthetic.make(kind, 8, 3, 2160, 3840, device=dev), 3)
    run("C5 4K batch 8 mosaic n_iter=10", synthetic.make("mosaic", 8, 3, 2160, 3840, device=dev), 10)
    run("C5 4K batch 4 mosaic RF prefilter", synthetic.make("mosaic", 4, 3, 2160, 3840, device=dev), 3,
        prefiltering=True, prefilter="rf")
    for kind in ("white", "mosaic"):
        run(f"C4 single 12000x9000 {kind}", synthetic.make(kind, 1, 3, 9000, 12000, device=dev), 5)
    # patch decomposition on the device: one 4K image, 400-px patches, 25 % overlap = 7 x 13 patches
    x4k = synthetic.make("mosaic", 1, 3, 2160, 3840, device=dev)
    mod = pb.PolyblurDeblurring(patch_decomposition=True, patch_size=400, patch_overlap=0.25, batch_size=8)
    ms = timed(lambda: mod(x4k, n_iter=3, alpha=6, beta=1, b=0.768))
    print(json.dumps({"config": "f3 patch decomposition, 1 x 4K, 400-px patches (91 patches)", "n_iter": 3, "ms": round(ms, 3),
                      "Mpix_s": round(2160 * 3840 / 1e6 / (ms / 1e3), 1)}), flush=True)
    # forward + backward on 8 x 1080p, with and without the estimator's gradient
    xb = synthetic.make("mosaic", 8, 3, 1080, 1920, device=dev)
    for flag in (True, False):
        def step():
            xg = xb.clone().requires_grad_(True)
            pb.polyblur_deblurring(xg, n_iter=3, alpha=6, beta=1, estimate_grad=flag).sum().backward()
            return xg.grad
        ms = timed(step)
        gr = step()
        print(json.dumps({"config": f"forward + backward 8 x 1080p mosaic n_iter=3 estimate_grad={flag}", "shape": list(xb.shape),
                          "n_iter": 3, "ms": round(ms, 3), "Mpix_s": round(8 * 1080 * 1920 / 1e6 / (ms / 1e3), 1),
                          "finite": bool(torch.isfinite(gr).all())}), flush=True)


if __name__ == "__main__":
    main()
