"""Golden vectors for the backward pass THROUGH HALO MASKING, from torch.autograd over the LIVE reference.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden_vjp_halo.py

Writes vjp_halo.npz next to this file:
  * stage_*: ``inverse_filtering_rank3(x, k, remove_halo=True, method='fft')`` (deblurring.py:211-239 with the
    bug-compatible mask of :171-208) and the gradients of sum(y * ybar) with respect to the image and to the kernel
    taps, with ``grad_img=None`` (gradients of the image itself) and with the gradients of another image ``x0``
    (then also the gradient with respect to ``x0``);
  * loop_*: ``polyblur_deblurring(x, n_iter=2, alpha=6, beta=1, remove_halo=True)`` and the gradient with respect to
    the input, estimator in the graph;
  * sat_*: the same loop with ``discard_saturation=True`` on an over-range image (and with halo masking on top);
  * bil_* / pre_*: the 5x5 bilateral filter alone, the loop with ``prefiltering=True`` (two iterations) and with halo
    masking on top (one iteration) -- from an out-of-place restatement of the filter, because the reference's own
    scales an autograd-saved tensor in place and raises under autograd (see main()).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden_vjp as base  # noqa: E402  (imports the reference, defines the synthetic images)
from polyblur import deblurring, filters  # noqa: E402


def main():
    out = {}
    g = torch.Generator().manual_seed(17)
    B, C, H, W = 2, 3, 40, 52
    x = base.mosaic(B, C, H, W, seed=4)
    x0 = base.textured(B, C, H, W, seed=6)
    kw = torch.from_numpy(filters.gaussian_filter((2.5, 1.2), 30 * np.pi / 180, k_size=np.array([25, 25])))
    kn = torch.from_numpy(filters.gaussian_filter((1.1, 0.8), 100 * np.pi / 180, k_size=np.array([25, 25])))
    kernels = torch.stack([kw, kn])[:, None].float().contiguous()
    ybar = torch.randn(B, C, H, W, generator=g)
    # grad_img = None
    xr = x.clone().requires_grad_(True)
    kr = kernels.clone().requires_grad_(True)
    y = deblurring.inverse_filtering_rank3(xr, kr, alpha=6, b=1, remove_halo=True, method="fft")
    gx, gk = torch.autograd.grad((y * ybar).sum(), (xr, kr))
    out["stage_x"] = x.numpy()
    out["stage_x0"] = x0.numpy()
    out["stage_kernels"] = kernels.numpy()
    out["stage_ybar"] = ybar.numpy()
    out["stage_self_y"] = y.detach().numpy()
    out["stage_self_grad"] = gx.numpy()
    out["stage_self_kernel_grad"] = gk.numpy()
    # grad_img of another image
    xr = x.clone().requires_grad_(True)
    x0r = x0.clone().requires_grad_(True)
    y = deblurring.inverse_filtering_rank3(xr, kernels, alpha=6, b=1, remove_halo=True,
                                           grad_img=filters.fourier_gradients(x0r), method="fft")
    gx, g0 = torch.autograd.grad((y * ybar).sum(), (xr, x0r))
    out["stage_other_y"] = y.detach().numpy()
    out["stage_other_grad"] = gx.numpy()
    out["stage_other_grad_x0"] = g0.numpy()
    frac = float(((y.detach() - deblurring.inverse_filtering_rank3(x, kernels, alpha=6, b=1, method="fft")).abs() > 1e-6)
                 .float().mean())
    print("stage: fraction of pixels the mask changes =", frac)

    # two iterations of the loop
    B, C, H, W = 2, 3, 64, 80
    x = base.textured(B, C, H, W, seed=5)
    ybar = torch.randn(B, C, H, W, generator=g)
    assert base.unique_maxima_margin(x) > 1e-4
    xr = x.clone().requires_grad_(True)
    y = deblurring.polyblur_deblurring(xr, n_iter=2, alpha=6, beta=1, remove_halo=True)
    (gx,) = torch.autograd.grad((y * ybar).sum(), xr)
    with torch.no_grad():
        x1 = deblurring.polyblur_deblurring(x, n_iter=1, alpha=6, beta=1, remove_halo=True)
    print("loop: smallest top-2 gap, iteration 2 =", base.unique_maxima_margin(x1))
    assert base.unique_maxima_margin(x1) > 1e-4
    out["loop_x"] = x.numpy()
    out["loop_ybar"] = ybar.numpy()
    out["loop_y"] = y.detach().numpy()
    out["loop_grad"] = gx.numpy()

    # the estimator's saturation mask (discard_saturation=True): an image in [0.65, 1.9], so that the
    # strongest gradients of the first iteration sit on masked pixels
    xs = (x * 1.5 + 0.5).contiguous()           # not clamped: the estimator normalises by min / max (:92-109)
    sat = float((xs.mean(1) > 0.99).float().mean())
    print("saturation case: fraction of masked pixels =", sat)
    assert sat > 0.01
    with torch.no_grad():
        k_on = base.blur_estimation.gaussian_blur_estimation(xs, q=0.0, c=0.352, b=0.768, discard_saturation=True)
        k_off = base.blur_estimation.gaussian_blur_estimation(xs, q=0.0, c=0.352, b=0.768, discard_saturation=False)
    print("saturation case: max |kernel(mask) - kernel(no mask)| =", float((k_on - k_off).abs().max()))
    assert float((k_on - k_off).abs().max()) > 1e-5
    xr = xs.clone().requires_grad_(True)
    y = deblurring.polyblur_deblurring(xr, n_iter=2, alpha=6, beta=1, discard_saturation=True)
    (gx,) = torch.autograd.grad((y * ybar).sum(), xr)
    out["sat_x"] = xs.numpy()
    out["sat_y"] = y.detach().numpy()
    out["sat_grad"] = gx.numpy()
    xr = xs.clone().requires_grad_(True)
    y = deblurring.polyblur_deblurring(xr, n_iter=1, alpha=6, beta=1, discard_saturation=True, remove_halo=True)
    (gx,) = torch.autograd.grad((y * ybar).sum(), xr)
    out["sat_halo_y"] = y.detach().numpy()
    out["sat_halo_grad"] = gx.numpy()

    # the bilateral prefilter (filters.py:107-148, deblurring.py:80-84, 99-110): the stage alone, two iterations of the
    # loop with prefiltering=True, and one with halo masking on top
    # The reference's own filter cannot be differentiated: bilateral_filter_loop_ scales the output of torch.exp in
    # place (filters.py:133, `F *= gw[y]...`), so autograd raises "modified by an inplace operation".  The goldens come
    # from the same arithmetic with that one statement out of place (checked bit-identical in the forward direction).
    from polyblur import utils as rutils

    def bilateral_oop(I, ksize=5, sigma_spatial=5.0, sigma_color=0.1):
        t = torch.arange(-ksize // 2 + 1, ksize // 2 + 1, device=I.device)
        xx, yy = torch.meshgrid(t, t, indexing="xy")
        gw = torch.exp(-(xx * xx + yy * yy) / (2 * sigma_spatial * sigma_spatial))
        I_padded = rutils.pad_with_kernel(I, ksize=ksize)
        var2 = 2 * sigma_color * sigma_color
        b_, c_, h, w = I.shape
        J = torch.zeros_like(I)
        Wt = torch.zeros_like(I)
        for yk in range(gw.shape[0]):
            I_shifted = rutils.extract_tiles(I_padded[..., yk:yk + h, :], kernel_size=(h, w), stride=1)
            F = I_shifted - I.unsqueeze(1)
            F = torch.exp(-F * F / var2)
            F = F * gw[yk].view(-1, 1, 1, 1)
            J = J + torch.sum(F * I_shifted, dim=1)
            Wt = Wt + torch.sum(F, dim=1)
        return J / (Wt + 1e-5)

    xb = (x + 0.02 * torch.randn(x.shape, generator=g)).clamp(0, 1).contiguous()
    with torch.no_grad():
        assert torch.equal(filters.bilateral_filter(xb), bilateral_oop(xb)), "restatement differs from the reference"
    filters.bilateral_filter = bilateral_oop          # deblurring.edge_aware_filtering looks it up at call time
    xr = xb.clone().requires_grad_(True)
    yb = filters.bilateral_filter(xr)
    (gx,) = torch.autograd.grad((yb * ybar).sum(), xr)
    out["bil_x"] = xb.numpy()
    out["bil_y"] = yb.detach().numpy()
    out["bil_grad"] = gx.numpy()
    assert base.unique_maxima_margin(xb) > 1e-4
    xr = xb.clone().requires_grad_(True)
    y = deblurring.polyblur_deblurring(xr, n_iter=2, alpha=6, beta=1, prefiltering=True)
    (gx,) = torch.autograd.grad((y * ybar).sum(), xr)
    with torch.no_grad():
        x1 = deblurring.polyblur_deblurring(xb, n_iter=1, alpha=6, beta=1, prefiltering=True)
    print("prefilter: smallest top-2 gap, iteration 2 =", base.unique_maxima_margin(x1))
    out["pre_y"] = y.detach().numpy()
    out["pre_grad"] = gx.numpy()
    xr = xb.clone().requires_grad_(True)
    y = deblurring.polyblur_deblurring(xr, n_iter=1, alpha=6, beta=1, prefiltering=True, remove_halo=True)
    (gx,) = torch.autograd.grad((y * ybar).sum(), xr)
    out["pre_halo_y"] = y.detach().numpy()
    out["pre_halo_grad"] = gx.numpy()
    np.savez_compressed(os.path.join(HERE, "vjp_halo.npz"), **out)
    for k, v in out.items():
        print(k, v.shape, float(np.abs(v).max()))


if __name__ == "__main__":
    main()
