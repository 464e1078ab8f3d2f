#!/usr/bin/env python
"""Prints the measured max-abs error of the CUDA path against the CPU oracle (float32 restatement of
the reference, and its float64 'truth' variant) for the cases DESIGN.md quotes.  Needs a GPU."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import polyblur_b200 as pb  # noqa: E402
from oracle import polyblur_oracle as po  # noqa: E402
from polyblur_b200 import synthetic  # noqa: E402


def maxabs(a, b):
    return float(np.max(np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64))))


def main():
    rows = []
    for kind, shape, n_iter in (("mosaic", (1, 3, 1080, 1920), 3), ("white", (1, 3, 1080, 1920), 3),
                                ("mosaic", (2, 3, 540, 960), 5), ("mosaic", (2, 3, 540, 960), 10)):
        x = synthetic.make(kind, *shape).numpy()
        ref32 = po.polyblur_deblurring(x, n_iter=n_iter, alpha=6, beta=1)
        ref64 = po.polyblur_deblurring(x, n_iter=n_iter, alpha=6, beta=1, dtype=np.float64)
        for engine, name in ((0, "auto"), (1, "spatial"), (2, "fft")):
            if engine == 1 and kind == "mosaic" and shape[-1] > 1000:
                continue                                  # 541-tap stencils at full HD: slow, same code as small cases
            out = pb.polyblur_deblurring(torch.from_numpy(x).cuda(), n_iter=n_iter, alpha=6, beta=1,
                                         engine=engine).cpu().numpy()
            rows.append({"input": kind, "shape": list(shape), "n_iter": n_iter, "engine": name,
                         "vs_oracle_f32": maxabs(out, ref32), "vs_truth_f64": maxabs(out, ref64),
                         "oracle_f32_vs_truth_f64": maxabs(ref32, ref64)})
            print(json.dumps(rows[-1]), flush=True)


if __name__ == "__main__":
    main()
