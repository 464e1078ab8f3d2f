"""8-bit image path around the engine (what polyblur's main.py does on the host,
main.py:80 ``img_as_float32`` and :146 ``img_as_ubyte``).

    deblur_uint8(images_u8, n_iter=3, alpha=6, beta=1, ...) -> uint8 images of the same layout

The conversion uint8 HWC -> float32 NCHW in [0,1] and back happens on the device, so one byte per
sample crosses PCIe in each direction instead of four.  Host batches are pipelined over three
streams (H2D / convert + Polyblur + convert / D2H per chunk of images).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib, sharding
from .deblurring import _make_params


def _convert_in(xu8: torch.Tensor, xf: torch.Tensor, stream: int) -> None:
    B, H, W, Cn = xu8.shape
    rc = _lib.lib().pb_u8hwc_to_f32nchw(xu8.data_ptr(), xf.data_ptr(), B, H, W, Cn, stream)
    _lib.check(rc, "pb_u8hwc_to_f32nchw")


def _convert_out(yf: torch.Tensor, yu8: torch.Tensor, stream: int) -> None:
    B, Cn, H, W = yf.shape
    rc = _lib.lib().pb_f32nchw_to_u8hwc(yf.data_ptr(), yu8.data_ptr(), B, Cn, H, W, stream)
    _lib.check(rc, "pb_f32nchw_to_u8hwc")


def deblur_uint8(images, n_iter=1, c=0.352, b=0.768, alpha=2, beta=3, sigma_r=0.8, sigma_s=2.0, ker_size=25,
                 q=0.0, remove_halo=False, edgetaping=False, prefiltering=False, discard_saturation=False,
                 max_chunks=4, **engine_kw):
    """Polyblur on 8-bit images.  ``images``: uint8 ndarray (H,W), (H,W,C) or (B,H,W,C), or a uint8
    torch tensor (B,H,W,C) on the CPU or on a CUDA device; the result has the same kind, layout and
    device.  Keyword arguments as ``polyblur_deblurring``."""
    is_np = isinstance(images, np.ndarray)
    if is_np:
        arr = np.ascontiguousarray(images)
        x = torch.from_numpy(arr if arr.flags.writeable else arr.copy())     # torch wants a writable buffer
    else:
        x = images
    if not isinstance(x, torch.Tensor) or x.dtype != torch.uint8:
        raise TypeError("deblur_uint8 expects uint8 images")
    shape_in = tuple(x.shape)
    if x.ndim == 2:
        x = x[None, :, :, None]
    elif x.ndim == 3:
        x = x[None]
    elif x.ndim != 4:
        raise ValueError("expected (H,W), (H,W,C) or (B,H,W,C)")
    x = x.contiguous()
    B, H, W, Cn = x.shape
    p = _make_params(n_iter, c, b, alpha, beta, sigma_r, sigma_s, ker_size, q, remove_halo, edgetaping,
                     prefiltering, discard_saturation, **engine_kw)
    dev = _lib.require_cuda(x)
    if n_iter == 0:
        return images
    with torch.cuda.device(dev):
        if not x.is_cuda and (p.flags & _lib.FLAG_EDGETAPER_BATCHMAX):
            # batch-global edgetaper normalisation (edgetaper.py:15,21): one engine call for the whole batch
            return_cpu = True
            x = x.to(dev)
        else:
            return_cpu = False
        if x.is_cuda:
            st = _lib.stream_ptr(dev)
            xf = torch.empty(B, Cn, H, W, dtype=torch.float32, device=dev)
            yf = torch.empty_like(xf)
            out = torch.empty_like(x)
            ws = _lib.workspace(B, Cn, H, W, p, dev)
            _convert_in(x, xf, st)
            rc = _lib.lib().pb_polyblur_f32(xf.data_ptr(), yf.data_ptr(), B, Cn, H, W, C.byref(p), ws.data_ptr(),
                                            ws.numel(), None, st)
            _lib.check(rc, "pb_polyblur_f32")
            _convert_out(yf, out, st)
            if return_cpu:
                out = out.cpu()
        else:
            out = _host_pipeline_u8(x, p, dev, max_chunks)
    out = out.reshape(shape_in)
    return out.numpy() if is_np else out


def _host_pipeline_u8(x: torch.Tensor, p, dev: torch.device, max_chunks: int, ramp=(2, 4)) -> torch.Tensor:
    """(B,H,W,C) uint8 CPU tensor in -> same out.  One byte per sample over PCIe: the kernels are the bottleneck, so
    few large chunks for their efficiency (each is one CUDA-graph launch of a cached engine: conversion, Polyblur,
    conversion -- deblurring._run_host_pipeline), with small chunks first and last so that the kernels start early and
    the copy left over at the end is short."""
    from .deblurring import _run_host_pipeline
    B = x.shape[0]
    if not x.is_pinned():
        x = x.pin_memory()
    host = torch.empty(x.shape, dtype=torch.uint8, pin_memory=True)
    sizes = sharding.pipeline_chunks(B, -(-B // max(1, min(max_chunks, B))), ramp)
    return _run_host_pipeline(x, host, sizes, p, dev, uint8_io=True)
