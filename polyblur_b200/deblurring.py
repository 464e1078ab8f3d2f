"""The two public surfaces of the reference, kept as drop-ins:

    polyblur_deblurring(img, n_iter, c, b, alpha, beta, ...)   polyblur/deblurring.py:23-96
    PolyblurDeblurring(nn.Module)                              polyblur/deblurring.py:250-394

Host code only: argument handling mirrors the reference, all arithmetic happens in
libpolyblur_sm100.so (hand-written sm_100a kernels) on the tensor's CUDA device, enqueued
on torch's current stream with no host synchronisation.  ndarray / CPU tensor input is
copied to the GPU and back.  There is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch
from torch import nn

from . import _lib, utils

_METHODS = ("fft", "direct", "direct_separable")


def _make_params(n_iter, c, b, alpha, beta, sigma_r, sigma_s, ker_size, q, remove_halo, edgetaping,
                 prefiltering, discard_saturation, engine=_lib.ENGINE_AUTO, prefilter="bilateral",
                 tap_rel_threshold=0.0):
    p = _lib.default_params()
    p.n_iter = int(n_iter)
    p.c, p.b, p.alpha, p.beta = float(c), float(b), float(alpha), float(beta)
    p.sigma_r, p.sigma_s, p.q = float(sigma_r), float(sigma_s), float(q)
    p.ker_size = int(ker_size)
    flags = 0
    if remove_halo:
        flags |= _lib.FLAG_REMOVE_HALO
    if edgetaping:
        flags |= _lib.FLAG_EDGETAPER | _lib.FLAG_EDGETAPER_BATCHMAX
    if prefiltering:
        flags |= _lib.FLAG_PREFILTER_RF if prefilter == "rf" else _lib.FLAG_PREFILTER
    if discard_saturation:
        flags |= _lib.FLAG_DISCARD_SATURATION
    p.flags = flags
    p.engine = int(engine)
    p.tap_rel_threshold = float(tap_rel_threshold)
    return p


def polyblur_device(x: torch.Tensor, p: "_lib.PbParams", out: torch.Tensor | None = None,
                    return_estimates: bool = False):
    """Run the whole Polyblur loop on a contiguous float32 CUDA tensor (B,C,H,W).

    Thin wrapper over ``pb_polyblur_f32``; everything is enqueued on the current stream.
    """
    assert x.is_cuda and x.dtype == torch.float32 and x.is_contiguous() and x.ndim == 4
    B, Cn, H, W = x.shape
    dev = x.device
    with torch.cuda.device(dev):
        if out is None:
            out = torch.empty_like(x)
        ws = _lib.workspace(B, Cn, H, W, p, dev)
        est = None
        if return_estimates:
            est = torch.empty(max(p.n_iter, 1), B, _lib.PB_EST_STRIDE, dtype=torch.float32, device=dev)
        rc = _lib.lib().pb_polyblur_f32(x.data_ptr(), out.data_ptr(), B, Cn, H, W, C.byref(p),
                                        ws.data_ptr(), ws.numel(), _lib.ptr(est), _lib.stream_ptr(dev))
        _lib.check(rc, "pb_polyblur_f32")
    return (out, est) if return_estimates else out


def polyblur_deblurring(img, n_iter=1, c=0.352, b=0.768, alpha=2, beta=3, sigma_r=0.8, sigma_s=2.0,
                        ker_size=25, q=0.0, n_angles=6, n_interpolated_angles=30, remove_halo=False,
                        edgetaping=False, prefiltering=False, discard_saturation=False,
                        multichannel_kernel=False, method='fft', verbose=False, **engine_kw):
    """Blind deblurring by polynomial reblurring; same signature and defaults as the
    reference (polyblur/deblurring.py:23-25).

    ``img``: (H,W) / (H,W,C) ndarray -> float32 ndarray squeezed like ``utils.to_array``;
    or a (B,C,H,W) float32 tensor -> tensor on the same device.  Every ``method`` gives the
    reference's ``'fft'`` result (its other methods are broken for batches).
    Extra keyword arguments (``engine``, ``prefilter``, ``tap_rel_threshold``,
    ``return_estimates``) are engine knobs that the reference does not have.
    """
    if method not in _METHODS:
        raise ValueError(f"unknown method {method!r}; expected one of {_METHODS}")
    if n_angles != 6 or n_interpolated_angles != 30:
        raise ValueError("only n_angles=6 and n_interpolated_angles=30 work in the reference "
                         "(SURVEY.md Appendix B.10); other values are rejected here")
    return_estimates = bool(engine_kw.pop("return_estimates", False))
    p = _make_params(n_iter, c, b, alpha, beta, sigma_r, sigma_s, ker_size, q, remove_halo, edgetaping,
                     prefiltering, discard_saturation, **engine_kw)

    flag_numpy = isinstance(img, np.ndarray)
    if flag_numpy:
        x = utils.to_tensor(img).unsqueeze(0)
    else:
        x = img
        if not isinstance(x, torch.Tensor):
            raise TypeError("img must be a numpy array or a torch tensor")
        if x.ndim != 4:
            raise ValueError("tensor input must be (B,C,H,W)")
        if x.dtype != torch.float32:
            raise TypeError(f"float32 only (got {x.dtype}), like the reference")
    if n_iter == 0:                     # the reference returns its input object (:60,:96)
        return utils.to_array(x) if flag_numpy else img

    dev = _lib.require_cuda(x)
    src_device = x.device
    xd = x.detach()
    if not xd.is_cuda and not flag_numpy:
        xd = xd.pin_memory() if not xd.is_pinned() and xd.numel() > (1 << 20) else xd
    xd = xd.to(dev, non_blocking=True).contiguous()
    res = polyblur_device(xd, p, return_estimates=return_estimates)
    out, est = res if return_estimates else (res, None)
    if flag_numpy:
        out = utils.to_array(out)
    elif src_device != dev:
        # device -> pinned host memory on the same stream, one synchronisation at the end
        host = torch.empty(out.shape, dtype=out.dtype, pin_memory=True)
        host.copy_(out, non_blocking=True)
        torch.cuda.current_stream(dev).synchronize()
        out = host
    if return_estimates:
        return out, est.to(src_device)
    return out


def inverse_filtering_rank3(img, kernel, alpha=2, b=4, correlate=False, remove_halo=False,
                            do_edgetaper=False, grad_img=None, method='direct', engine=_lib.ENGINE_AUTO):
    """pad -> polynomial deconvolution on the torus -> crop -> clamp
    (polyblur/deblurring.py:211-239), for explicit kernels (B,1,k,k) or (1,1,k,k).

    Always the 'fft' (circular, replicate-padded) semantics.  ``correlate`` rotates the
    kernel by 180 degrees like the reference."""
    if remove_halo or do_edgetaper:
        raise NotImplementedError("remove_halo / do_edgetaper are not built yet in inverse_filtering_rank3")
    if img.dtype != torch.float32 or img.ndim != 4:
        raise TypeError("img must be a float32 (B,C,H,W) tensor")
    dev = _lib.require_cuda(img)
    src = img.device
    x = img.detach().to(dev).contiguous()
    B, Cn, H, W = x.shape
    k = kernel.detach().to(dev, torch.float32)
    if correlate:
        k = torch.rot90(k, k=2, dims=(-2, -1))
    ks = k.shape[-1]
    if k.shape[-2] != ks:
        raise ValueError("square kernels only")
    if k.shape[1] != 1:
        raise NotImplementedError("one kernel per image (B,1,k,k); per-channel kernels are not supported")
    k = k.expand(B, 1, ks, ks).contiguous()
    with torch.cuda.device(dev):
        p = _lib.default_params()
        ws = _lib.workspace(B, Cn, H, W, p, dev)
        out = torch.empty_like(x)
        rc = _lib.lib().pb_deconv_f32(x.data_ptr(), out.data_ptr(), B, Cn, H, W, k.data_ptr(), ks,
                                      float(alpha), float(b), int(engine), ws.data_ptr(), ws.numel(),
                                      _lib.stream_ptr(dev))
        _lib.check(rc, "pb_deconv_f32")
    return out.to(src)


class PolyblurDeblurring(nn.Module):
    """nn.Module wrapper with the reference's constructor and forward signature
    (polyblur/deblurring.py:250-347).  No parameters, no buffers.

    ``patch_decomposition=True`` raises NameError in the reference; here it
    raises NotImplementedError until the patch path is built.
    """

    def __init__(self, patch_decomposition=False, patch_size=400, patch_overlap=0.25, batch_size=1):
        super().__init__()
        self.batch_size = batch_size
        self.patch_decomposition = patch_decomposition
        self.patch_size = (patch_size, patch_size)
        self.patch_overlap = patch_overlap

    def forward(self, images, n_iter=1, c=0.352, b=0.468, alpha=2, beta=4, sigma_s=2, ker_size=25,
                sigma_r=0.4, q=0.0, n_angles=6, n_interpolated_angles=30, remove_halo=False,
                edgetaping=False, prefiltering=False, discard_saturation=False, multichannel_kernel=False,
                method='fft', device=None):
        if self.patch_decomposition:
            raise NotImplementedError("patch_decomposition is not built yet (it raises NameError in the reference)")
        if device is not None and isinstance(images, torch.Tensor):
            images = images.to(device)
        return polyblur_deblurring(images, n_iter=n_iter, c=c, b=b, alpha=alpha, beta=beta, ker_size=ker_size,
                                   sigma_s=sigma_s, sigma_r=sigma_r, remove_halo=remove_halo,
                                   edgetaping=edgetaping, prefiltering=prefiltering,
                                   discard_saturation=discard_saturation,
                                   multichannel_kernel=multichannel_kernel, method=method, q=q,
                                   n_angles=n_angles, n_interpolated_angles=n_interpolated_angles)
