"""Golden vectors for the backward pass, from torch.autograd over the LIVE reference.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden_vjp.py

Writes vjp.npz next to this file:
  * deconv_*: one reference ``inverse_filtering_rank3`` (method='fft') and the gradient of
    sum(y * ybar) with respect to the image, for a wide and a narrow Gaussian kernel;
  * deconv_*_kernel_grad: the gradient of the same scalar with respect to the kernel taps;
  * est_*: the estimator alone, gradient of <gaussian_blur_estimation(x), kbar> with respect to x
    (image 1 has tied maxima);
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


def textured(B, C, H, W, seed):
    """Blurred mosaic plus a smooth random field: along a straight mosaic edge the gradient is constant to
    within rounding, so the arg-max pixel of a directional maximum (where autograd puts its sub-gradient)
    would depend on the last bit of the FFT; the texture makes every maximum unique by a clear margin."""
    g = torch.Generator().manual_seed(seed + 1000)
    base = mosaic(B, C, H, W, seed)
    n = torch.randn(B, C, H, W, generator=g)
    k = torch.from_numpy(filters.gaussian_filter((3.0, 3.0), 0.0, k_size=np.array([25, 25])))[None, None].repeat(B, 1, 1, 1)
    n = filters.convolve2d(n, k, method="fft")
    n = n / n.abs().amax(dim=(1, 2, 3), keepdim=True)
    return (0.1 + 0.8 * base + 0.08 * n).clamp(0, 1).contiguous()


def unique_maxima_margin(x):
    """Smallest relative gap between the two largest |cos gx - sin gy| over the 7 angles and the images."""
    gray = x.mean(1, keepdim=True)
    gn = blur_estimation.normalize(gray, q=0)
    gx, gy = blur_estimation.compute_gradients(gn, blur_estimation.get_saturation_mask(gray, False))
    ang = torch.linspace(0, np.pi, 7)
    worst = 1.0
    for b in range(x.shape[0]):
        for j in range(7):
            v = (torch.cos(ang[j]) * gx[b, 0] - torch.sin(ang[j]) * gy[b, 0]).abs().flatten()
            top = torch.topk(v, 2).values
            worst = min(worst, float((top[0] - top[1]) / top[0]))
    return worst


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
        # gradient with respect to the kernel taps of the same scalar
        kr = kernels.clone().requires_grad_(True)
        y = deblurring.inverse_filtering_rank3(x, kr, alpha=alpha, b=beta, method="fft")
        (gk,) = torch.autograd.grad((y * ybar).sum(), kr)
        out[f"deconv_{tag}_kernel_grad"] = gk.numpy()
    out["deconv_x"] = x.numpy()
    out["deconv_kernels"] = kernels.numpy()
    out["deconv_ybar"] = ybar.numpy()

    # ---- the estimator alone: gradient of <kernel(x), kbar> with respect to the image ----------
    B, C, H, W = 2, 3, 48, 60
    x = textured(B, C, H, W, seed=9)
    x[1, :, 5:9, 7:12] = x[1].max()          # ties at the maximum of image 1 (amax spreads its gradient)
    print("estimator case: smallest top-2 gap of a directional maximum =", unique_maxima_margin(x))
    assert unique_maxima_margin(x) > 1e-4
    kbar = torch.randn(B, 1, 25, 25, generator=g)
    xr = x.clone().requires_grad_(True)
    k = blur_estimation.gaussian_blur_estimation(xr, q=0.0, n_angles=6, n_interpolated_angles=30, c=0.352, b=0.768,
                                                 ker_size=25)
    (gx,) = torch.autograd.grad((k * kbar).sum(), xr)
    out["est_x"] = x.numpy()
    out["est_kbar"] = kbar.numpy()
    out["est_kernel"] = k.detach().numpy()
    out["est_grad"] = gx.numpy()

    # ---- the scalar chain alone: 7 maxima -> interpolation, arg-min direction, affine model, Gaussian taps ----
    g2 = torch.Generator().manual_seed(21)         # its own stream: the other cases keep their random draws
    m = torch.rand(6, 7, generator=g2) * 0.25 + 0.05
    m[0] = torch.tensor([0.30, 0.22, 0.15, 0.12, 0.16, 0.24, 0.30])
    m[1] = m[1] * 4                               # strong gradients: sigma, rho clamp to 0.3 (zero gradient)
    m[2] = m[2] * 0.2                             # weak gradients: clamp to 4.0
    kb = torch.randn(6, 25, 25, generator=g2)
    mr = m.clone().requires_grad_(True)
    thetas = torch.linspace(0, 180, 7).unsqueeze(0)
    interp = torch.arange(0, 180, 6.0).unsqueeze(0)
    m_n, m_o, th = blur_estimation.find_maximal_blur_direction(mr, thetas, interp)
    sg, rh = blur_estimation.compute_gaussian_parameters(m_n, m_o, c=0.352, b=0.768)
    kk = blur_estimation.create_gaussian_filter(th, sg, rh, ksize=25)
    (gm,) = torch.autograd.grad((kk[:, 0] * kb).sum(), mr)
    out["chainrule_m"] = m.numpy()
    out["chainrule_kbar"] = kb.numpy()
    out["chainrule_mbar"] = gm.numpy()
    out["chainrule_sigma_rho"] = torch.cat([sg, rh], dim=1).detach().numpy()

    # ---- two iterations, estimate held constant ----------------------------------------------
    B, C, H, W = 2, 3, 64, 80
    x = textured(B, C, H, W, seed=5)
    ybar = torch.randn(B, C, H, W, generator=g)
    print("chain case: smallest top-2 gap, iteration 1 =", unique_maxima_margin(x))
    assert unique_maxima_margin(x) > 1e-4
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
    with torch.no_grad():
        x1 = deblurring.inverse_filtering_rank3(x, ks[0], alpha=alpha, b=beta, method="fft")
    print("chain case: smallest top-2 gap, iteration 2 =", unique_maxima_margin(x1))
    assert unique_maxima_margin(x1) > 1e-4
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
    xr = x.clone().requires_grad_(True)
    y1 = deblurring.polyblur_deblurring(xr, n_iter=1, c=c, b=b, alpha=alpha, beta=beta)
    (g1,) = torch.autograd.grad((y1 * ybar).sum(), xr)
    out["chain_full_grad_1iter"] = g1.numpy()
    rel = float((gf - gx).abs().max() / gx.abs().max())
    print("max |full - constant-estimate| / max |grad| =", rel)
    out["chain_full_rel_diff"] = np.float64(rel)
    np.savez_compressed(os.path.join(HERE, "vjp.npz"), **out)
    print("wrote vjp.npz:", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
