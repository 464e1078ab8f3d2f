"""Stage-level drop-in for polyblur/edgetaper.py (reference)."""
from __future__ import annotations

import torch

from . import _lib
from .filters import _prep


def edgetaper(img, kernel, n_tapers=3, method='fft', batch_max=True):
    """Taper the borders of an (already padded) image with its blurred self
    (edgetaper.py:26-33, method='fft' semantics = circular correlation on the image's torus).

    ``batch_max=True`` reproduces the reference's normalisation of the taper weights by the
    maximum over the whole batch (edgetaper.py:15,21)."""
    x, dev, src = _prep(img, "edgetaper")
    B, C, H, W = x.shape
    k = kernel.detach().to(dev, torch.float32)
    ks = k.shape[-1]
    if k.shape[-2] != ks or k.shape[1] != 1:
        raise ValueError("kernel must be (B,1,k,k) or (1,1,k,k)")
    k = k.expand(B, 1, ks, ks).contiguous()
    with torch.cuda.device(dev):
        ws = torch.empty(B * 4096 + B * (H + W + 128) * 4 + 4096 + x.numel() * 4, dtype=torch.uint8, device=dev)
        out = torch.empty_like(x)
        flags = _lib.FLAG_EDGETAPER_BATCHMAX if batch_max else 0
        rc = _lib.lib().pb_edgetaper_f32(x.data_ptr(), out.data_ptr(), B, C, H, W, k.data_ptr(), ks,
                                         int(n_tapers), flags, ws.data_ptr(), ws.numel(), _lib.stream_ptr(dev))
        _lib.check(rc, "pb_edgetaper_f32")
    return out.to(src)
