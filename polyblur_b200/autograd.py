"""Backward pass: gradients with respect to the image.

The reference is differentiable because it is written in torch ops (README.md:70).  Two levels here:

* ``inverse_filtering_rank3_vjp`` / ``DeconvolutionFunction`` -- the exact vector-Jacobian product of
  ``deblurring.inverse_filtering_rank3`` (default flags, polyblur/deblurring.py:211-239) with respect
  to ``img`` for a given kernel: replicate pad, circular polynomial filter, crop and clamp, transposed.
  It matches torch.autograd over the reference (tests/golden/vjp.npz).
* ``polyblur_deblurring_grad`` / ``PolyblurFunction`` -- the Polyblur loop.  With
  ``estimate_grad=True`` (default, what the reference does) the gradient also flows through the blur
  estimator of every iteration: kernel taps -> sigma, rho -> interpolated directional maxima -> the
  arg-max pixels of ``|cos gx - sin gy|`` -> spectral derivative -> range normalisation (including the
  min / max pixels) -> channel mean.  The image-sized steps run in CUDA (csrc/backward.cu:
  ``pb_estimate_trace_f32``, ``pb_kernel_grad_f32``, ``pb_estimator_vjp_f32``); the chain of a dozen
  scalars per image between the 25 x 25 kernel gradient and the 7 maxima is evaluated with torch autograd
  on (B, 7) / (B, 25, 25) tensors.  ``estimate_grad=False`` holds the estimates constant (the
  estimator under ``torch.no_grad()``), which removes the spiky rows / columns through the arg-max
  pixels from the gradient.

Halo masking (``remove_halo=True``, polyblur/deblurring.py:171-208) is differentiable as a composite:
``EstimateKernelFunction`` (the estimator as a node: image -> kernel taps) -> ``DeconvolutionFunction`` without
its clamp -> ``halo_masking`` written with the differentiable ``filters.fourier_gradients`` of this package and
elementwise torch operations -> clamp.  It matches torch.autograd over the reference (tests/golden/vjp_halo.npz),
including the path through the gradients of the original image that the mask is built from.

``discard_saturation=True`` is differentiable as well (the trace's arg-max search leaves the saturated pixels out, as
the forward estimator does), and so is ``prefiltering=True`` (the reference's 5x5 bilateral filter with its own
backward kernel, ``pb_bilateral_vjp_f32``).  Edgetaper and the quantile normalisation are not differentiable here
(they raise); the kernels, the estimator traces and the unclamped iterates are kept for the backward pass (n_iter
extra images of memory).
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


def _bw_workspace(B, Cn, H, W, ks, engine, dev):
    n = _lib.lib().pb_backward_workspace_bytes(B, Cn, H, W, int(ks), int(engine))
    return torch.empty(n, dtype=torch.uint8, device=dev)


def estimate_trace(x: torch.Tensor, discard_saturation: bool = False):
    """Forward trace of the estimator (pb_estimate_trace_ex_f32): (B,24) floats and (B,8) pixel indices;
    ``discard_saturation`` leaves the pixels with gray > 0.99 out of the arg-max search (blur_estimation.py:83-88)."""
    B, Cn, H, W = x.shape
    dev = x.device
    with torch.cuda.device(dev):
        ws = _bw_workspace(B, Cn, H, W, 25, _lib.ENGINE_AUTO, dev)
        tf = torch.empty(B, 24, dtype=torch.float32, device=dev)
        tp = torch.empty(B, 8, dtype=torch.int32, device=dev)
        flags = _lib.FLAG_DISCARD_SATURATION if discard_saturation else 0
        rc = _lib.lib().pb_estimate_trace_ex_f32(x.data_ptr(), B, Cn, H, W, flags, tf.data_ptr(), tp.data_ptr(),
                                                 ws.data_ptr(), ws.numel(), _lib.stream_ptr(dev))
        _lib.check(rc, "pb_estimate_trace_ex_f32")
    return tf, tp


def kernel_grad(x: torch.Tensor, grad_out: torch.Tensor, preclamp, k: torch.Tensor, alpha, beta, engine):
    """d <grad_out, inverse_filtering_rank3(x, k)> / d k  -> (B,1,ks,ks)  (pb_kernel_grad_f32)."""
    B, Cn, H, W = x.shape
    ks = k.shape[-1]
    dev = x.device
    with torch.cuda.device(dev):
        ws = _bw_workspace(B, Cn, H, W, ks, engine, dev)
        kb = torch.empty(B, 1, ks, ks, dtype=torch.float32, device=dev)
        rc = _lib.lib().pb_kernel_grad_f32(x.data_ptr(), grad_out.data_ptr(), _lib.ptr(preclamp), B, Cn, H, W,
                                           k.data_ptr(), ks, float(alpha), float(beta), int(engine), kb.data_ptr(),
                                           ws.data_ptr(), ws.numel(), _lib.stream_ptr(dev))
        _lib.check(rc, "pb_kernel_grad_f32")
    return kb


_KEYS_WEIGHTS: dict = {}


def _keys_weights(dev) -> torch.Tensor:
    """(30,7) normalised Keys weights (pb_keys_weights), uploaded once per device."""
    key = str(dev)
    if key not in _KEYS_WEIGHTS:
        import ctypes as C
        buf = (C.c_float * 210)()
        _lib.lib().pb_keys_weights(buf)
        _KEYS_WEIGHTS[key] = torch.tensor(list(buf), dtype=torch.float32, device=dev).view(30, 7)
    return _KEYS_WEIGHTS[key]


def _maxima_grad(m: torch.Tensor, kbar: torch.Tensor, c: float, b: float, ks: int) -> torch.Tensor:
    """Gradient with respect to the 7 directional maxima given the gradient with respect to the kernel taps:
    blur_estimation.py:138-232 (Keys interpolation, arg-min direction, affine model with clamping, Gaussian
    taps) on (B,7) tensors, differentiated by torch autograd."""
    dev = m.device
    W30 = _keys_weights(dev)
    with torch.enable_grad():
        mm = m.detach().clone().requires_grad_(True)
        mags = mm @ W30.t()                                          # (B,30)
        i_min = torch.argmin(mags, dim=-1, keepdim=True)
        theta_deg = i_min * 6
        m_n = torch.take_along_dim(mags, i_min, dim=-1)
        i_o = ((theta_deg + 90) % 180) // 6
        m_o = torch.take_along_dim(mags, i_o, dim=-1)
        cc, bb = c * c, b * b
        sigma = torch.sqrt(torch.clamp(cc / (m_n * m_n + 1e-8) - bb, min=0.09, max=16.0))
        rho = torch.sqrt(torch.clamp(cc / (m_o * m_o + 1e-8) - bb, min=0.09, max=16.0))
        th = -(theta_deg.float() * 3.14159274101257324 / 180.0)
        cs, sn = torch.cos(th), torch.sin(th)
        il1, il2 = 1.0 / (sigma * sigma), 1.0 / (rho * rho)
        s00 = cs * cs * il1 + sn * sn * il2
        s01 = sn * cs * (il1 - il2)
        s11 = cs * cs * il2 + sn * sn * il1
        t = torch.arange(ks, device=dev) - ((ks - 1) // 2)
        X, Y = torch.meshgrid(t, t, indexing="xy")
        X, Y = X.float()[None], Y.float()[None]
        quad = s00[:, :, None] * X * X + 2.0 * s01[:, :, None] * X * Y + s11[:, :, None] * Y * Y
        E = torch.exp(-0.5 * quad)
        K = E / E.sum(dim=(-1, -2), keepdim=True)
        (K * kbar.view(-1, ks, ks)).sum().backward()
    return mm.grad.contiguous()


def estimator_vjp(x: torch.Tensor, mbar: torch.Tensor, tf: torch.Tensor, tp: torch.Tensor, grad_img: torch.Tensor):
    """Adds the gradient through the estimator into grad_img (pb_estimator_vjp_f32)."""
    B, Cn, H, W = x.shape
    dev = x.device
    with torch.cuda.device(dev):
        ws = _bw_workspace(B, Cn, H, W, 25, _lib.ENGINE_AUTO, dev)
        rc = _lib.lib().pb_estimator_vjp_f32(x.data_ptr(), mbar.data_ptr(), tf.data_ptr(), tp.data_ptr(),
                                             grad_img.data_ptr(), B, Cn, H, W, ws.data_ptr(), ws.numel(),
                                             _lib.stream_ptr(dev))
        _lib.check(rc, "pb_estimator_vjp_f32")


class DeconvolutionFunction(torch.autograd.Function):
    """``inverse_filtering_rank3(img, kernel, alpha, b)`` with default flags, differentiable in ``img`` and in
    the kernel taps."""

    @staticmethod
    def forward(ctx, img, kernel, alpha, beta, engine, clamp=True):
        dev = _lib.require_cuda(img)
        x = img.detach().to(dev).contiguous()
        k = kernel.detach().to(dev, torch.float32)
        k = k.expand(x.shape[0], 1, k.shape[-2], k.shape[-1]).contiguous()
        v = _deconv_noclamp(x, k, alpha, beta, engine)
        ctx.save_for_backward(x, k, v)
        ctx.meta = (alpha, beta, engine, img.device, tuple(kernel.shape), kernel.device, kernel.dtype, bool(clamp))
        return (v.clamp(0.0, 1.0) if clamp else v).to(img.device)

    @staticmethod
    def backward(ctx, grad_out):
        x, k, v = ctx.saved_tensors
        alpha, beta, engine, src, kshape, kdev, kdtype, clamp = ctx.meta
        go = grad_out.detach().to(x.device).contiguous()
        pre = v if clamp else None          # clamp=False: the unclamped result was handed out (halo masking follows)
        g = gk = None
        if ctx.needs_input_grad[0]:
            g = inverse_filtering_rank3_vjp(go, k, alpha, beta, preclamp=pre, engine=engine).to(src)
        if ctx.needs_input_grad[1]:
            # d <grad_out, y> / d taps (pb_kernel_grad_f32); a kernel shared by the batch collects every image's term
            gk = kernel_grad(x, go, pre, k, alpha, beta, engine)
            if kshape[0] == 1 and gk.shape[0] != 1:
                gk = gk.sum(dim=0, keepdim=True)
            gk = gk.to(kdev, kdtype)
        return g, gk, None, None, None, None


class EstimateKernelFunction(torch.autograd.Function):
    """``gaussian_blur_estimation(img)`` (default options, polyblur/blur_estimation.py:18-79) as a node of the
    autograd graph: image -> (B,1,ks,ks) kernel taps; backward = taps -> 7 directional maxima (``_maxima_grad``) ->
    arg-max pixels, transposed spectral derivative, range normalisation (``pb_estimator_vjp_f32``)."""

    @staticmethod
    def forward(ctx, img, c, b, ker_size, discard_saturation=False):
        dev = _lib.require_cuda(img)
        x = img.detach().to(dev).contiguous()
        k = blur_estimation.gaussian_blur_estimation(x, q=0.0, c=c, b=b, ker_size=ker_size,
                                                     discard_saturation=bool(discard_saturation)).contiguous()
        tf, tp = estimate_trace(x, bool(discard_saturation))
        ctx.save_for_backward(x, tf, tp)
        ctx.meta = (float(c), float(b), int(ker_size), img.device)
        return k.to(img.device)

    @staticmethod
    def backward(ctx, kbar):
        x, tf, tp = ctx.saved_tensors
        c, b, ks, src = ctx.meta
        kb = kbar.detach().to(x.device, torch.float32).contiguous()
        mbar = _maxima_grad(tf[:, :7], kb, c, b, ks)
        gin = torch.zeros_like(x)
        estimator_vjp(x, mbar, tf, tp, gin)
        return gin.to(src), None, None, None, None


def halo_masking(img: torch.Tensor, imout: torch.Tensor, grad_img=None) -> torch.Tensor:
    """Differentiable ``halo_masking`` (polyblur/deblurring.py:171-208), bug-compatible like the fused kernels
    (csrc/stages.cu): M = -gx gout_x - gy gy (:174), z = max(M / (nM + M), 0), out = imout + z (img - imout).
    The spectral derivatives are this package's CUDA kernels (``filters.fourier_gradients``, differentiable)."""
    from . import filters
    gx, gy = filters.fourier_gradients(img) if grad_img is None else grad_img
    gout_x, _ = filters.fourier_gradients(imout)
    M = (-gx * gout_x) + (-gy * gy)
    nM = torch.sum(gx * gx + gy * gy, dim=(-2, -1), keepdim=True)
    z = torch.clamp(M / (nM + M), min=0)
    return imout + z * (img - imout)


def inverse_filtering_rank3_halo(img, kernel, alpha, beta, grad_img, engine) -> torch.Tensor:
    """Differentiable ``inverse_filtering_rank3(..., remove_halo=True)`` (no edgetaper): deblurring.py:225-239."""
    v = DeconvolutionFunction.apply(img, kernel, alpha, beta, engine, False)
    return torch.clamp(halo_masking(img, v, grad_img), 0.0, 1.0)


def polyblur_deblurring_halo_grad(img: torch.Tensor, n_iter=1, c=0.352, b=0.768, alpha=2, beta=3, ker_size=25,
                                  engine=_lib.ENGINE_AUTO, estimate_grad=True, discard_saturation=False,
                                  remove_halo=True, prefiltering=False) -> torch.Tensor:
    """Differentiable ``polyblur_deblurring`` with ``remove_halo`` and / or ``prefiltering`` (deblurring.py:60-88) as a
    composite of autograd nodes.  The halo mask of every iteration is built from the gradients of the ORIGINAL image,
    which therefore also receives gradient through them; with the prefilter the smooth component (5x5 bilateral filter,
    differentiable: ``filters.bilateral_filter``) is deconvolved and the residual added back (:80-84)."""
    if img.dtype != torch.float32 or img.ndim != 4:
        raise TypeError("img must be a float32 (B,C,H,W) tensor")
    from . import filters
    dev = _lib.require_cuda(img)
    x = img.to(dev)
    grad_img = filters.fourier_gradients(x) if remove_halo else None
    cur = x
    for _ in range(int(n_iter)):
        if estimate_grad:
            k = EstimateKernelFunction.apply(cur, c, b, ker_size, bool(discard_saturation))
        else:
            with torch.no_grad():
                k = blur_estimation.gaussian_blur_estimation(cur.detach(), q=0.0, c=c, b=b, ker_size=ker_size,
                                                             discard_saturation=bool(discard_saturation))
        src, noise = cur, None
        if prefiltering:
            src = filters.bilateral_filter(cur)              # edge_aware_filtering (:99-110)
            noise = cur - src
        if remove_halo:
            y = inverse_filtering_rank3_halo(src, k, alpha, beta, grad_img, engine)
        else:
            y = DeconvolutionFunction.apply(src, k, alpha, beta, engine, True)
        cur = (y if noise is None else y + noise).clamp(0.0, 1.0)
    return cur.to(img.device)


class PolyblurFunction(torch.autograd.Function):
    """The Polyblur loop (polyblur/deblurring.py:68-88, default options), differentiable in the image."""

    @staticmethod
    def forward(ctx, img, n_iter, c, b, alpha, beta, ker_size, engine, estimate_grad, discard_saturation=False):
        dev = _lib.require_cuda(img)
        x0 = img.detach().to(dev).contiguous()
        cur = x0
        saved = [x0]
        for _ in range(int(n_iter)):
            k = blur_estimation.gaussian_blur_estimation(cur, q=0.0, c=c, b=b, ker_size=ker_size,
                                                         discard_saturation=bool(discard_saturation)).contiguous()
            v = _deconv_noclamp(cur, k, alpha, beta, engine)
            saved += [k, v]
            if estimate_grad:
                saved += list(estimate_trace(cur, bool(discard_saturation)))
            cur = v.clamp(0.0, 1.0)
        ctx.save_for_backward(*saved)
        ctx.meta = (int(n_iter), c, b, alpha, beta, int(ker_size), engine, bool(estimate_grad), img.device)
        return cur.to(img.device)

    @staticmethod
    def backward(ctx, grad_out):
        saved = ctx.saved_tensors
        n_iter, c, b, alpha, beta, ks, engine, estimate_grad, src = ctx.meta
        per = 4 if estimate_grad else 2
        x0 = saved[0]
        g = grad_out.detach().to(x0.device).contiguous()
        for t in range(n_iter - 1, -1, -1):
            rec = saved[1 + per * t: 1 + per * (t + 1)]
            k, v = rec[0], rec[1]
            gin = inverse_filtering_rank3_vjp(g, k, alpha, beta, preclamp=v, engine=engine)
            if estimate_grad:
                xt = x0 if t == 0 else saved[1 + per * (t - 1) + 1].clamp(0.0, 1.0)
                tf, tp = rec[2], rec[3]
                kb = kernel_grad(xt, g, v, k, alpha, beta, engine)
                mbar = _maxima_grad(tf[:, :7], kb, float(c), float(b), ks)
                estimator_vjp(xt, mbar, tf, tp, gin)
            g = gin
        return (g.to(src),) + (None,) * 9


def polyblur_deblurring_grad(img: torch.Tensor, n_iter=1, c=0.352, b=0.768, alpha=2, beta=3, ker_size=25,
                             engine=_lib.ENGINE_AUTO, estimate_grad=True, discard_saturation=False) -> torch.Tensor:
    """Differentiable ``polyblur_deblurring`` for (B,C,H,W) float32 tensors (default options, optionally with the
    estimator's saturation mask)."""
    if img.dtype != torch.float32 or img.ndim != 4:
        raise TypeError("img must be a float32 (B,C,H,W) tensor")
    return PolyblurFunction.apply(img, n_iter, c, b, alpha, beta, ker_size, engine, estimate_grad,
                                  bool(discard_saturation))
