"""Small run of every engine and option for compute-sanitizer (memcheck / racecheck / synccheck):
    compute-sanitizer --tool memcheck python tools/sanitize_smoke.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import polyblur_b200 as pb  # noqa: E402
from polyblur_b200 import autograd as ag, synthetic  # noqa: E402


def main():
    torch.manual_seed(0)
    for shape in [(2, 3, 120, 136), (1, 1, 37, 53), (1, 3, 64, 200)]:
        for kind in ("mosaic", "white"):
            x = synthetic.make(kind, *shape).cuda()
            for engine in (0, 1, 2):
                y = pb.polyblur_deblurring(x, n_iter=2, alpha=6, beta=1, engine=engine)
                assert bool(torch.isfinite(y).all())
    x = synthetic.make("mosaic", 2, 3, 96, 104).cuda()
    for kw in (dict(remove_halo=True), dict(edgetaping=True), dict(prefiltering=True), dict(q=0.01),
               dict(discard_saturation=True), dict(prefiltering=True, prefilter="rf")):
        y = pb.polyblur_deblurring(x, n_iter=2, alpha=2, beta=3, **kw)
        assert bool(torch.isfinite(y).all()), kw
    xg = x.clone().requires_grad_(True)
    pb.polyblur_deblurring(xg, n_iter=2, alpha=6, beta=1).sum().backward()
    assert bool(torch.isfinite(xg.grad).all())
    xg = x.clone().requires_grad_(True)
    pb.polyblur_deblurring(xg, n_iter=2, alpha=6, beta=1, estimate_grad=False, engine=2).sum().backward()
    assert bool(torch.isfinite(xg.grad).all())
    gx, gy = pb.filters.fourier_gradients(x)
    u8 = (x.permute(0, 2, 3, 1) * 255).round().to(torch.uint8).contiguous().cpu()
    from polyblur_b200 import io as pbio
    pbio.deblur_uint8(u8, n_iter=1, alpha=6, beta=1)
    torch.cuda.synchronize()
    print("sanitize smoke ok")


if __name__ == "__main__":
    main()
