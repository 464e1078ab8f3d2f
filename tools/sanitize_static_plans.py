"""The compile-time-plan kernels (full-HD and 4K sides: csrc/estimate3.cu, the second-generation FFT passes) under
compute-sanitizer:  compute-sanitizer --tool memcheck|racecheck|synccheck python tools/sanitize_static_plans.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import polyblur_b200 as pb  # noqa: E402
from polyblur_b200 import synthetic  # noqa: E402

for kind in ("mosaic", "white"):
    for (H, W) in ((1080, 1920), (2160, 3840), (1081, 1920), (1080, 1090)):
        x = synthetic.make(kind, 1, 3, H, W, device="cuda")
        y = pb.polyblur_deblurring(x, n_iter=2, alpha=6, beta=1)
        torch.cuda.synchronize()
        assert torch.isfinite(y).all() and float(y.min()) >= 0 and float(y.max()) <= 1
        print(kind, H, W, "ok", flush=True)
