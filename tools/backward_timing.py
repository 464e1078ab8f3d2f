"""Forward + backward time of polyblur_deblurring on 8 x 1080p (with / without the estimator's gradient)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import polyblur_b200 as pb  # noqa: E402
from polyblur_b200 import synthetic  # noqa: E402


def main():
    xb = synthetic.make("mosaic", 8, 3, 1080, 1920, device="cuda")
    for flag in (True, False):
        def step():
            xg = xb.clone().requires_grad_(True)
            pb.polyblur_deblurring(xg, n_iter=3, alpha=6, beta=1, estimate_grad=flag).sum().backward()
        for _ in range(2):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            step()
        e1.record()
        torch.cuda.synchronize()
        print(json.dumps({"estimate_grad": flag, "ms": round(e0.elapsed_time(e1) / 5, 2)}))


if __name__ == "__main__":
    main()
