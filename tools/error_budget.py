"""Error budget of the CUDA path against the fp64 oracle, per stage (run on the GPU box).

For each test input: (A) estimator taps vs the fp64 oracle on the same input, (B) the
deconvolution alone with the oracle's kernel, (C) end to end.  Test tooling: imports oracle/.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import polyblur_oracle as po  # noqa: E402
import polyblur_b200 as pb  # noqa: E402


def cu(x):
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


def budget(name, x, **kw):
    n_iter = kw["n_iter"]
    t64, t32 = [], []
    ref64 = po.polyblur_deblurring(x, dtype=np.float64, trace=t64, **kw)
    ref32 = po.polyblur_deblurring(x, dtype=np.float32, trace=t32, **kw)
    out, est = pb.polyblur_deblurring(cu(x), return_estimates=True, **kw)
    out = out.cpu().numpy().astype(np.float64)
    est = est.cpu().numpy()
    print(f"== {name} {x.shape} {kw}")
    print(f"   e2e: cuda-vs-fp64 {np.abs(out - ref64).max():.2e}  cuda-vs-oracle32 {np.abs(out - ref32).max():.2e}"
          f"  oracle32-vs-fp64 {np.abs(ref32 - ref64).max():.2e}")
    for it in range(n_iter):
        m64 = t64[it]["mags"]
        print(f"   it{it}: mags rel err cuda {np.abs(est[it, :, :7] / m64 - 1).max():.2e} oracle32 "
              f"{np.abs(t32[it]['mags'] / m64 - 1).max():.2e} | sigma rel cuda "
              f"{np.abs(est[it, :, 8] / t64[it]['sigma'] - 1).max():.2e} oracle32 "
              f"{np.abs(t32[it]['sigma'] / t64[it]['sigma'] - 1).max():.2e} | theta {est[it, :, 7].astype(int)} "
              f"{t64[it]['theta_deg']} | sigma {est[it, :, 8]} rho {est[it, :, 9]}")
    # deconvolution alone: oracle fp64 kernel of iteration 0 applied to the input
    k = t64[0]["kernel"]
    d64 = po.inverse_filtering_rank3(x, k, alpha=kw["alpha"], b=kw["beta"], dtype=np.float64)
    d32 = po.inverse_filtering_rank3(x, k.astype(np.float32), alpha=kw["alpha"], b=kw["beta"], dtype=np.float32)
    dcu = pb.deblurring.inverse_filtering_rank3(cu(x), cu(k.astype(np.float32)), alpha=kw["alpha"], b=kw["beta"])
    dcu = dcu.cpu().numpy()
    print(f"   deconv alone (iter-0 kernel): cuda-vs-fp64 {np.abs(dcu - d64).max():.2e}  oracle32-vs-fp64 "
          f"{np.abs(d32 - d64).max():.2e}")
    # sensitivity: fp64 deconvolution with sigma perturbed by 1e-5 relative
    tr = t64[0]
    k2 = po.gaussian_kernel(tr["theta"], tr["sigma"] * (1 + 1e-5), tr["rho"], dtype=np.float64)
    d64b = po.inverse_filtering_rank3(x, k2, alpha=kw["alpha"], b=kw["beta"], dtype=np.float64)
    print(f"   sensitivity: d(out) for 1e-5 relative change of sigma = {np.abs(d64b - d64).max():.2e}")


def main():
    op = np.load(os.path.join(ROOT, "tests/golden/options.npz"))
    sc = np.load(os.path.join(ROOT, "tests/golden/small_cases.npz"))
    budget("options/module_default", op["in"], n_iter=2, c=0.352, b=0.468, alpha=2, beta=4)
    budget("options/default", op["in"], n_iter=2, c=0.352, b=0.768, alpha=6, beta=1)
    budget("mosaic_rgb_96x120", sc["mosaic_rgb_96x120/in"], n_iter=3, c=0.352, b=0.768, alpha=6, beta=1)


if __name__ == "__main__":
    main()
