#!/usr/bin/env python
"""C2 (32 x 3 x 1080 x 1920, n_iter=3) device-resident step time against pb_params.chunk_images (images per engine
pass: does a group whose half spectrum fits the 126 MB L2 beat the whole-batch passes?).  CUDA-graph replay, CUDA events."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import polyblur_b200 as pb  # noqa: E402
from polyblur_b200 import synthetic  # noqa: E402
from polyblur_b200.deblurring import GraphedPolyblur  # noqa: E402


def main():
    B, H, W = 32, 1080, 1920
    groups = [int(a) for a in sys.argv[1:]] or [0, 1, 2, 3, 4, 8, 16]
    for kind in ("mosaic", "white"):
        x = synthetic.make(kind, B, 3, H, W, device="cuda")
        ref = None
        for G in groups:
            g = GraphedPolyblur((B, 3, H, W), n_iter=3, alpha=6, beta=1, chunk_images=G)
            g.x.copy_(x)
            for _ in range(3):
                g.graph.replay()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                g.graph.replay()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 10
            same = None
            if ref is None:
                ref = g.out.clone()
            else:
                same = bool(torch.equal(ref, g.out))
            print(json.dumps({"dist": kind, "chunk_images": G, "ms": round(ms, 3), "Mpix_s": round(B * H * W / 1e3 / ms),
                              "bitwise_equal_to_first": same}), flush=True)
            del g
            torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
