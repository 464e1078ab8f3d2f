"""Stage-level drop-ins for polyblur/blur_estimation.py (reference)."""
from __future__ import annotations

import torch

from . import _lib
from .filters import _prep


def estimate_parameters(imgc: torch.Tensor, c=0.362, b=0.464, q=0.0, discard_saturation=False):
    """Run the estimator and return the taps the reference computes on the way
    (blur_estimation.py:59-70): dict of (B,7) mags, (B,) theta_deg, sigma, rho, m_normal, m_ortho.

    Note the reference's own default is q=1e-4; the Polyblur loop always passes q
    explicitly (deblurring.py:71-74).  q > 0 = quantile range normalisation (radix select on
    the device, csrc/quantile.cu).
    """
    x, dev, src = _prep(imgc, "gaussian_blur_estimation")
    B, C, H, W = x.shape
    with torch.cuda.device(dev):
        p = _lib.default_params()
        p.q = float(q)
        ws = _lib.workspace(B, C, H, W, p, dev)
        est = torch.empty(B, _lib.PB_EST_STRIDE, dtype=torch.float32, device=dev)
        flags = _lib.FLAG_DISCARD_SATURATION if discard_saturation else 0
        rc = _lib.lib().pb_estimate_f32(x.data_ptr(), B, C, H, W, float(c), float(b), float(q), flags,
                                        est.data_ptr(), ws.data_ptr(), ws.numel(), _lib.stream_ptr(dev))
        _lib.check(rc, "pb_estimate_f32")
    est = est.to(src)
    return dict(mags=est[:, :7], theta_deg=est[:, 7], sigma=est[:, 8], rho=est[:, 9],
                m_normal=est[:, 10], m_ortho=est[:, 11])


def create_gaussian_filter(thetas, sigmas, rhos, ksize=25):
    """(B,1) thetas [rad], sigmas, rhos -> (B,1,ksize,ksize) normalised kernels
    (blur_estimation.py:211-232)."""
    dev = _lib.require_cuda(sigmas)
    src = sigmas.device
    th = thetas.detach().to(dev, torch.float32).reshape(-1).contiguous()
    sg = sigmas.detach().to(dev, torch.float32).reshape(-1).contiguous()
    rh = rhos.detach().to(dev, torch.float32).reshape(-1).contiguous()
    B = sg.numel()
    with torch.cuda.device(dev):
        ws = torch.empty(B * 4096, dtype=torch.uint8, device=dev)
        k = torch.empty(B, 1, ksize, ksize, dtype=torch.float32, device=dev)
        rc = _lib.lib().pb_make_kernel_f32(th.data_ptr(), sg.data_ptr(), rh.data_ptr(), B, int(ksize),
                                           k.data_ptr(), ws.data_ptr(), ws.numel(), _lib.stream_ptr(dev))
        _lib.check(rc, "pb_make_kernel_f32")
    return k.to(src)


def gaussian_blur_estimation(imgc, q=0.0001, n_angles=6, n_interpolated_angles=30, c=0.362, b=0.464,
                             ker_size=25, discard_saturation=False, multichannel=False, thetas=None,
                             interpolated_thetas=None, return_2d_filters=True):
    """One anisotropic Gaussian blur kernel per image (blur_estimation.py:18-79).

    Same signature as the reference.  The kernel is always estimated on the channel mean:
    ``multichannel=True`` is a no-op for RGB and crashes for other channel counts in the
    reference.  ``return_2d_filters=False`` returns
    (sigma, rho, theta) -- what the reference's broken separable path intended (:75-77).
    """
    if n_angles != 6 or n_interpolated_angles != 30:
        raise ValueError("only n_angles=6, n_interpolated_angles=30 are meaningful (SURVEY.md B.10)")
    e = estimate_parameters(imgc, c=c, b=b, q=q, discard_saturation=discard_saturation)
    theta = (e["theta_deg"] * 3.14159274101257324 / 180.0).float()
    if not return_2d_filters:
        return e["sigma"][:, None], e["rho"][:, None], theta[:, None]
    return create_gaussian_filter(theta[:, None], e["sigma"][:, None], e["rho"][:, None], ksize=ker_size)
