"""Golden vectors for the backward pass, from torch.autograd over the LIVE reference.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden_vjp.py

Writes vjp.npz next to this file:
  * deconv_*: one reference ``inverse_filtering_rank3`` (method='fft') and the gradient of
    sum(y * ybar) with respect to the image, for a wide and a narrow Gaussian kernel;
  * chain_*: two Polyblur iterations in which the blur estimate of each iteration is taken from the
    reference estimator under no_grad (the estimate held constant), and the gradient with respect to the
    input; ``chain_full_grad`` is the gradient of the reference's own polyblur_deblurring, estimator in
    the graph, for comparison only.
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("POLYBLUR_REFERENCE", "/root/reference")
sk = types.ModuleType("skimage")
sk.img_as_float32 = lambda x: x
sys.modules["skimage"] = sk
sys.path.insert(0, REF)

import polyblur  # noqa: E402,F401  (the reference)
from polyblur import blur_estimation, deblurring, filters  # noqa: E402

torch.set_num_threads(1)


def mosaic(B, C, H, W, seed, sigma=(2.5, 1.2), theta_deg=30.0, block=12):
    g = torch.Generator().manual_seed(seed)
    small = torch.rand(B, C, -(-H // block), -(-W // block), generator=g)
    img = small.repeat_interleave(block, -2).repeat_interleave(block, -1)[..., :H, :W].contiguous()
    k = torch.from_numpy(filters.gaussian_filter(sigma, theta_deg * np.pi / 180, k_size=np.array([25, 25])))
    k = k[None, None].repeat(B, 1, 1, 1)
    return filters.convolve2d(img, k, method="fft").clamp(0, 1).contiguous()


def main():
    out = {}
    g = torch.Generator().manual_seed(11)
    # ---- one deconvolution ------------------------------------------------------------------
    B, C, H, W = 2, 3, 40, 52
    x = mosaic(B, C, H, W, seed=3)
    x[0] = x[0] * 1.3 - 0.1            # push part of the result outside [0,1] so that the clamp mask matters
    x = x.clamp(0, 1).contiguous()
    kw = torch.from_numpy(filters.gaussian_filter((2.5, 1.2), 30 * np.pi / 180, k_size=np.array([25, 25])))
    kn = torch.from_numpy(filters.gaussian_filter((0.45, 0.35), 100 * np.pi / 180, k_size=np.array([25, 25])))
    kernels = torch.stack([kw, kn])[:, None].float().contiguous()
    ybar = torch.randn(B, C, H, W, generator=g)
    for tag, (alpha, beta) in {"a6b1": (6, 1), "a2b4": (2, 4)}.items():
        xr = x.clone().requires_grad_(True)
        y = deblurring.inverse_filtering_rank3(xr, kernels, alpha=alpha, b=beta, method="fft")
        (gx,) = torch.autograd.grad((y * ybar).sum(), xr)
        out[f"deconv_{tag}_y"] = y.detach().numpy()
        out[f"deconv_{tag}_grad"] = gx.numpy()
    out["deconv_x"] = x.numpy()
    out["deconv_kernels"] = kernels.numpy()
    out["deconv_ybar"] = ybar.numpy()

    # ---- two iterations, estimate held constant ----------------------------------------------
    B, C, H, W = 2, 3, 64, 80
    x = mosaic(B, C, H, W, seed=5)
    ybar = torch.randn(B, C, H, W, generator=g)
    c, b, alpha, beta = 0.352, 0.768, 6, 1
    xr = x.clone().requires_grad_(True)
    cur = xr
    ks = []
    for _ in range(2):
        with torch.no_grad():
            k = blur_estimation.gaussian_blur_estimation(cur.detach(), q=0.0, n_angles=6, n_interpolated_angles=30,
                                                         c=c, b=b, ker_size=25)
        ks.append(k)
        cur = deblurring.inverse_filtering_rank3(cur, k, alpha=alpha, b=beta, method="fft")
    (gx,) = torch.autograd.grad((cur * ybar).sum(), xr)
    out["chain_x"] = x.numpy()
    out["chain_ybar"] = ybar.numpy()
    out["chain_y"] = cur.detach().numpy()
    out["chain_grad"] = gx.numpy()
    out["chain_kernels"] = torch.stack(ks).numpy()
    # the reference's own gradient (estimator in the graph), for the record
    xr = x.clone().requires_grad_(True)
    yf = deblurring.polyblur_deblurring(xr, n_iter=2, c=c, b=b, alpha=alpha, beta=beta)
    (gf,) = torch.autograd.grad((yf * ybar).sum(), xr)
    out["chain_full_grad"] = gf.numpy()
    rel = float((gf - gx).abs().max() / gx.abs().max())
    print("max |full - constant-estimate| / max |grad| =", rel)
    out["chain_full_rel_diff"] = np.float64(rel)
    np.savez_compressed(os.path.join(HERE, "vjp.npz"), **out)
    print("wrote vjp.npz:", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
