"""Stage-level drop-in for polyblur/domain_transform.py (reference): the Gastal-Oliveira
domain-transform recursive filter, run by the CUDA library (csrc/stages.cu)."""
from __future__ import annotations

import torch

from . import _lib
from .filters import _prep


def recursive_filter(I, sigma_s=60, sigma_r=0.4, num_iterations=3, joint_image=None):
    """Edge-aware smoothing with the recursive filter (domain_transform.py:6-63).

    Same signature as the reference; also what the native prototype exports
    (polyblur/domain_transform/RF.cpp:98), here correct for any batch size."""
    x, dev, src = _prep(I, "recursive_filter")
    B, C, H, W = x.shape
    j = None
    if joint_image is not None:
        j, _, _ = _prep(joint_image, "recursive_filter(joint_image)")
        if j.shape[0] != B or j.shape[-2:] != x.shape[-2:]:
            raise ValueError("joint_image must have the batch and spatial size of I")
        if j.shape[1] != C:
            raise NotImplementedError("joint_image with a different channel count is not supported")
    with torch.cuda.device(dev):
        ws = torch.empty(2 * (B * H * W * 4 + 256), dtype=torch.uint8, device=dev)
        out = torch.empty_like(x)
        rc = _lib.lib().pb_recursive_filter_f32(x.data_ptr(), _lib.ptr(j), out.data_ptr(), B, C, H, W,
                                                float(sigma_s), float(sigma_r), int(num_iterations),
                                                ws.data_ptr(), ws.numel(), _lib.stream_ptr(dev))
        _lib.check(rc, "pb_recursive_filter_f32")
    return out.to(src)


def normalized_convolution(I, sigma_s=60, sigma_r=0.4, num_iterations=3):
    """Edge-aware smoothing with the normalized convolution (box filter in the transformed
    domain): the reference's native prototype polyblur/domain_transform/NC.cpp:143-204, which
    only handles one 3-channel image; any batch and channel count here."""
    x, dev, src = _prep(I, "normalized_convolution")
    B, C, H, W = x.shape
    with torch.cuda.device(dev):
        ws = torch.empty(2 * (B * H * W * 4 + 256) + 2 * (B * C * H * W * 4 + 256), dtype=torch.uint8, device=dev)
        out = torch.empty_like(x)
        rc = _lib.lib().pb_normalized_convolution_f32(x.data_ptr(), out.data_ptr(), B, C, H, W, float(sigma_s),
                                                      float(sigma_r), int(num_iterations), ws.data_ptr(),
                                                      ws.numel(), _lib.stream_ptr(dev))
        _lib.check(rc, "pb_normalized_convolution_f32")
    return out.to(src)
