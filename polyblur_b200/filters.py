"""Stage-level drop-ins for polyblur/filters.py (reference), backed by the CUDA library."""
from __future__ import annotations

import numpy as np
import torch

from . import _lib


def _prep(images: torch.Tensor, what: str, differentiable: bool = False):
    if not isinstance(images, torch.Tensor) or images.ndim != 4:
        raise ValueError(f"{what}: expected a (B,C,H,W) tensor")
    if images.dtype != torch.float32:
        raise TypeError(f"{what}: float32 only (got {images.dtype}), like the reference")
    if images.requires_grad and torch.is_grad_enabled() and not differentiable:
        # the reference's stage is written in differentiable torch ops; dropping the autograd history
        # silently would hand back wrong (zero) gradients
        raise NotImplementedError(f"{what}: no backward pass is built for this stage; call it under "
                                  "torch.no_grad() or on a detached tensor")
    dev = _lib.require_cuda(images)
    src_device = images.device
    x = images.detach().to(dev, non_blocking=True).contiguous()
    return x, dev, src_device


def fourier_gradients(images: torch.Tensor):
    """Spectral derivative along W and H of every channel -> (gx, gy).

    Replaces filters.fourier_gradients (polyblur/filters.py:159-186): full complex fft2,
    multiply by 2*pi*f*i, ifft2.  Here: independent 1-D spectral derivatives of the rows and
    the columns, computed by the on-chip FFT kernels (csrc/estimate.cu k_rows / k_cols).
    """
    if isinstance(images, torch.Tensor) and images.requires_grad and torch.is_grad_enabled():
        return _FourierGradients.apply(images)
    x, dev, src = _prep(images, "fourier_gradients", differentiable=True)
    B, C, H, W = x.shape
    with torch.cuda.device(dev):
        p = _lib.default_params()
        ws = _lib.workspace(B, C, H, W, p, dev)
        gx = torch.empty_like(x)
        gy = torch.empty_like(x)
        rc = _lib.lib().pb_fourier_gradients_f32(x.data_ptr(), gx.data_ptr(), gy.data_ptr(), B, C, H, W,
                                                 ws.data_ptr(), ws.numel(), _lib.stream_ptr(dev))
        _lib.check(rc, "pb_fourier_gradients_f32")
    return gx.to(src), gy.to(src)


class _FourierGradients(torch.autograd.Function):
    """The spectral derivative is a real skew-symmetric circulant (i omega is odd), so its transpose is its
    negative: grad_in = -(D_x gx_bar + D_y gy_bar)."""

    @staticmethod
    def forward(ctx, images):
        return fourier_gradients(images.detach())

    @staticmethod
    def backward(ctx, gx_bar, gy_bar):
        ax, _ = fourier_gradients(gx_bar.detach().contiguous())
        _, by = fourier_gradients(gy_bar.detach().contiguous())
        return -(ax + by)


def bilateral_filter(I, ksize=5, sigma_spatial=5.0, sigma_color=0.1):
    """5x5 bilateral filter with per-channel range weights (polyblur/filters.py:107-148)."""
    if ksize != 5:
        raise NotImplementedError("only the reference's 5x5 window is built")
    if isinstance(I, torch.Tensor) and I.requires_grad and torch.is_grad_enabled():
        return _BilateralFilter.apply(I, float(sigma_spatial), float(sigma_color))
    x, dev, src = _prep(I, "bilateral_filter", differentiable=True)
    B, C, H, W = x.shape
    with torch.cuda.device(dev):
        out = torch.empty_like(x)
        rc = _lib.lib().pb_bilateral_f32(x.data_ptr(), out.data_ptr(), B, C, H, W, float(sigma_spatial),
                                         float(sigma_color), _lib.stream_ptr(dev))
        _lib.check(rc, "pb_bilateral_f32")
    return out.to(src)


class _BilateralFilter(torch.autograd.Function):
    """Backward of the 5x5 bilateral filter (pb_bilateral_vjp_f32): both the averaged samples and every range weight
    are differentiated, as torch.autograd does over filters.py:107-148."""

    @staticmethod
    def forward(ctx, I, sigma_spatial, sigma_color):
        x = I.detach()
        ctx.save_for_backward(x)
        ctx.sig = (sigma_spatial, sigma_color)
        return bilateral_filter(x, 5, sigma_spatial, sigma_color)

    @staticmethod
    def backward(ctx, gbar):
        (x0,) = ctx.saved_tensors
        x, dev, src = _prep(x0, "bilateral_filter", differentiable=True)
        g = gbar.detach().to(dev, torch.float32).contiguous()
        B, C, H, W = x.shape
        with torch.cuda.device(dev):
            gin = torch.empty_like(x)
            rc = _lib.lib().pb_bilateral_vjp_f32(x.data_ptr(), g.data_ptr(), gin.data_ptr(), B, C, H, W,
                                                 float(ctx.sig[0]), float(ctx.sig[1]), _lib.stream_ptr(dev))
            _lib.check(rc, "pb_bilateral_vjp_f32")
        return gin.to(src), None, None


def gaussian_filter(sigma, theta, shift=np.array([0.0, 0.0]), k_size=np.array([15, 15])):
    """NumPy generator of a generalised 2-D Gaussian kernel (polyblur/filters.py:198-234);
    used by the CLI's synthetic degradation, not on the GPU path."""
    l1, l2 = sigma
    th = -theta
    Q = np.array([[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]])
    cov = Q @ np.diag([l1 ** 2, l2 ** 2]) @ Q.T
    inv = np.linalg.inv(cov)
    kx, ky = int(k_size[0]), int(k_size[1])
    mu = np.array([kx // 2, ky // 2], dtype=np.float64) - np.asarray(shift, dtype=np.float64)
    X, Y = np.meshgrid(np.arange(kx), np.arange(ky))
    zx, zy = X - mu[0], Y - mu[1]
    quad = inv[0, 0] * zx * zx + 2 * inv[0, 1] * zx * zy + inv[1, 1] * zy * zy
    raw = np.exp(-0.5 * quad).astype(np.float32)
    if raw.sum() < 1e-2:
        k = np.zeros_like(raw)
        k[kx // 2, ky // 2] = 1
        return k
    return raw / raw.sum()
