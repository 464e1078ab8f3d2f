"""Host-side helpers mirroring polyblur/utils.py of the reference (pure torch/numpy)."""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F


def to_tensor(array):
    """(H,W) / (H,W,C) ndarray -> (C,H,W) float tensor, no rescaling (utils.py:8-21)."""
    array = np.asarray(array)
    if array.ndim == 2:
        array = array[None]
    else:
        array = np.transpose(array, (2, 0, 1))
    return torch.from_numpy(np.ascontiguousarray(array)).float()


def to_array(tensor):
    """(..,C,H,W) tensor -> squeezed (H,W) / (H,W,C) ndarray (utils.py:24-31)."""
    t = tensor.detach().squeeze().cpu()
    if t.ndim == 2:
        return t.numpy()
    return t.permute(1, 2, 0).numpy()


def to_float(img):
    """uint8/uint16 ndarray -> float32 in [0,1] (utils.py:34-38, skimage.img_as_float32)."""
    img = np.asarray(img)
    if img.dtype == np.uint8:
        return img.astype(np.float32) / 255.0
    if img.dtype == np.uint16:
        return img.astype(np.float32) / 65535.0
    return img.astype(np.float32)


def to_uint(img):
    """float ndarray -> uint8 exactly like the reference's utils.to_uint (utils.py:41-45):
    ``(255 * img).astype(np.uint8)``, i.e. TRUNCATION towards zero (0.999 -> 254), not rounding.  Input outside
    [0,1] is clipped first (the reference's cast is undefined there).  The command line and the 8-bit pipeline use
    :func:`to_ubyte` instead, as main.py:146 does."""
    img = np.clip(np.asarray(img, dtype=np.float32), 0, 1)
    return (255 * img).astype(np.uint8)


def to_ubyte(img):
    """float ndarray in [0,1] -> uint8 with rounding to nearest: skimage.img_as_ubyte, which the reference's
    command line applies to the result (main.py:146); csrc/io.cu does the same on the device."""
    img = np.clip(np.asarray(img, dtype=np.float64), 0, 1)
    return np.rint(img * 255.0).astype(np.uint8)


def pad_with_kernel(img, kernel=None, ksize=3, mode="replicate"):
    """Pad by kernel.shape[-1]//2 (utils.py:48-53)."""
    ks = kernel.shape[-1] // 2 if kernel is not None else ksize // 2
    return F.pad(img, (ks, ks, ks, ks), mode=mode)


def crop_with_kernel(img, kernel=None, ksize=3):
    """Crop by kernel.shape[-1]//2 (utils.py:56-61)."""
    ks = kernel.shape[-1] // 2 if kernel is not None else ksize // 2
    return img[..., ks:-ks, ks:-ks]
