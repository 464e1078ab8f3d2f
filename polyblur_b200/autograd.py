"""Backward pass: gradients with respect to the image.

The reference is differentiable because it is written in torch ops (README.md:70).  Two levels here:

* ``inverse_filtering_rank3_vjp`` / ``DeconvolutionFunction`` -- the exact vector-Jacobian product of
  ``deblurring.inverse_filtering_rank3`` (default flags, polyblur/deblurring.py:211-239) with respect
  to ``img`` for a given kernel: replicate pad, circular polynomial filter, crop and clamp, transposed.
  It matches torch.autograd over the reference (tests/golden/vjp.npz).
* ``polyblur_deblurring_grad`` / ``PolyblurFunction`` -- the Polyblur loop with the blur estimate of
  every iteration held constant (as if ``gaussian_blur_estimation`` ran under ``torch.no_grad()``).
  The reference also differentiates through its estimator; that term is a sub-gradient through the
  arg-max pixels of the seven directional maxima and min / max of the gray image, concentrated on the
  rows and columns through those few pixels, and is deliberately not reproduced.

Only the default options are differentiable (no halo masking, edgetaper, prefilter); the kernels and
the unclamped iterates are kept for the backward pass (n_iter extra images of memory).
"""
from __future__ import annotations

import torch

from . import _lib
from . import blur_estimation


def _deconv_noclamp(x: torch.Tensor, k: torch.Tensor, alpha, beta, engine) -> torch.Tensor:
    """Forward deconvolution without the final clamp (pb_deconv_ex_f32, PB_FLAG_NO_CLAMP)."""
    B, Cn, H, W = x.shape
    ks = k.shape[-1]
    dev = x.device
    with torch.cuda.device(dev):
        p = _lib.default_params()
        p.ker_size = ks
        p.engine = int(engine)
        ws = _lib.workspace(B, Cn, H, W, p, dev)
        out = torch.empty_like(x)
        rc = _lib.lib().pb_deconv_ex_f32(x.data_ptr(), out.data_ptr(), B, Cn, H, W, k.data_ptr(), ks, float(alpha),
                                         float(beta), int(engine), _lib.FLAG_NO_CLAMP, None, None, ws.data_ptr(),
                                         ws.numel(), _lib.stream_ptr(dev))
        _lib.check(rc, "pb_deconv_ex_f32")
    return out


def inverse_filtering_rank3_vjp(grad_out: torch.Tensor, kernel: torch.Tensor, alpha=2, b=4, preclamp=None,
                                correlate=False, engine=_lib.ENGINE_AUTO) -> torch.Tensor:
    """grad_img of ``inverse_filtering_rank3(img, kernel, alpha, b)`` (default flags) for the upstream gradient
    ``grad_out``; ``preclamp`` = the unclamped forward result (None: the clamp is taken as inactive)."""
    if grad_out.dtype != torch.float32 or grad_out.ndim != 4:
        raise TypeError("grad_out must be a float32 (B,C,H,W) tensor")
    dev = _lib.require_cuda(grad_out)
    src = grad_out.device
    g = grad_out.detach().to(dev).contiguous()
    B, Cn, H, W = g.shape
    k = kernel.detach().to(dev, torch.float32)
    if correlate:
        k = torch.rot90(k, k=2, dims=(-2, -1))
    ks = k.shape[-1]
    if k.shape[-2] != ks or k.shape[1] != 1:
        raise ValueError("one square kernel per image: (B,1,k,k) or (1,1,k,k)")
    k = k.expand(B, 1, ks, ks).contiguous()
    pre = None
    if preclamp is not None:
        pre = preclamp.detach().to(dev, torch.float32).contiguous()
        if pre.shape != g.shape:
            raise ValueError("preclamp must have grad_out's shape")
    with torch.cuda.device(dev):
        n = _lib.lib().pb_deconv_vjp_workspace_bytes(B, Cn, H, W, ks, int(engine))
        ws = torch.empty(n, dtype=torch.uint8, device=dev)
        gin = torch.empty_like(g)
        rc = _lib.lib().pb_deconv_vjp_f32(g.data_ptr(), _lib.ptr(pre), gin.data_ptr(), B, Cn, H, W, k.data_ptr(), ks,
                                          float(alpha), float(b), int(engine), ws.data_ptr(), ws.numel(),
                                          _lib.stream_ptr(dev))
        _lib.check(rc, "pb_deconv_vjp_f32")
    return gin.to(src)


class DeconvolutionFunction(torch.autograd.Function):
    """``inverse_filtering_rank3(img, kernel, alpha, b)`` with default flags, differentiable in ``img``."""

    @staticmethod
    def forward(ctx, img, kernel, alpha, beta, engine):
        dev = _lib.require_cuda(img)
        x = img.detach().to(dev).contiguous()
        k = kernel.detach().to(dev, torch.float32)
        k = k.expand(x.shape[0], 1, k.shape[-2], k.shape[-1]).contiguous()
        v = _deconv_noclamp(x, k, alpha, beta, engine)
        ctx.save_for_backward(k, v)
        ctx.meta = (alpha, beta, engine, img.device)
        return v.clamp(0.0, 1.0).to(img.device)

    @staticmethod
    def backward(ctx, grad_out):
        k, v = ctx.saved_tensors
        alpha, beta, engine, src = ctx.meta
        g = inverse_filtering_rank3_vjp(grad_out.contiguous(), k, alpha, beta, preclamp=v, engine=engine)
        return g.to(src), None, None, None, None


class PolyblurFunction(torch.autograd.Function):
    """The Polyblur loop (polyblur/deblurring.py:68-88, default options), differentiable in the image with the
    per-iteration blur estimates held constant."""

    @staticmethod
    def forward(ctx, img, n_iter, c, b, alpha, beta, ker_size, q, discard_saturation, engine):
        dev = _lib.require_cuda(img)
        cur = img.detach().to(dev).contiguous()
        saved = []
        for _ in range(int(n_iter)):
            k = blur_estimation.gaussian_blur_estimation(cur, q=q, c=c, b=b, ker_size=ker_size,
                                                         discard_saturation=discard_saturation)
            v = _deconv_noclamp(cur, k.contiguous(), alpha, beta, engine)
            saved += [k, v]
            cur = v.clamp(0.0, 1.0)
        ctx.save_for_backward(*saved)
        ctx.meta = (alpha, beta, engine, img.device)
        return cur.to(img.device)

    @staticmethod
    def backward(ctx, grad_out):
        saved = ctx.saved_tensors
        alpha, beta, engine, src = ctx.meta
        dev = saved[0].device
        g = grad_out.detach().to(dev).contiguous()
        for t in range(len(saved) // 2 - 1, -1, -1):
            g = inverse_filtering_rank3_vjp(g, saved[2 * t], alpha, beta, preclamp=saved[2 * t + 1], engine=engine)
        return (g.to(src),) + (None,) * 9


def polyblur_deblurring_grad(img: torch.Tensor, n_iter=1, c=0.352, b=0.768, alpha=2, beta=3, ker_size=25, q=0.0,
                             discard_saturation=False, engine=_lib.ENGINE_AUTO) -> torch.Tensor:
    """Differentiable ``polyblur_deblurring`` for (B,C,H,W) float32 tensors (default options only)."""
    if img.dtype != torch.float32 or img.ndim != 4:
        raise TypeError("img must be a float32 (B,C,H,W) tensor")
    return PolyblurFunction.apply(img, n_iter, c, b, alpha, beta, ker_size, q, discard_saturation, engine)
