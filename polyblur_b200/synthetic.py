"""Synthetic inputs of SURVEY.md 8(d): float32 NCHW in [0,1], per-image seeds so that a
sharded run sees exactly the images of the unsharded run.

    white  : torch.rand -- the literal "synthetic RGB batch"; the estimator saturates at
             sigma = rho = 0.3 (near-delta kernel).
    mosaic : random 60 px constant blocks circularly blurred by an anisotropic Gaussian
             (sigma 2.5 / 1.2 at 30 degrees) -- a mildly blurred image with wide estimated
             kernels (sigma ~2.9 -> 2.2 -> 1.4 over three iterations).

Input generation only (not on the measured path): it uses torch ops on whatever device is
asked for.
"""
from __future__ import annotations

import math

import torch


def _gaussian_kernel(sigma, rho, theta, ksize=25):
    t = torch.arange(ksize, dtype=torch.float64) - ksize // 2
    Y, X = torch.meshgrid(t, t, indexing="ij")
    c, s = math.cos(-theta), math.sin(-theta)
    a = c * c / sigma ** 2 + s * s / rho ** 2
    b = s * c * (1 / sigma ** 2 - 1 / rho ** 2)
    d = c * c / rho ** 2 + s * s / sigma ** 2
    k = torch.exp(-0.5 * (a * X * X + 2 * b * X * Y + d * Y * Y))
    return (k / k.sum()).float()


def white(n, C, H, W, first_index=0, base_seed=0, device="cpu"):
    out = torch.empty(n, C, H, W, dtype=torch.float32)
    for i in range(n):
        g = torch.Generator().manual_seed(base_seed + first_index + i)
        out[i] = torch.rand(C, H, W, generator=g)
    return out.to(device)


def mosaic(n, C, H, W, first_index=0, base_seed=0, device="cpu", block=60, sigma=2.5, rho=1.2,
           theta_deg=30.0):
    k = _gaussian_kernel(sigma, rho, theta_deg * math.pi / 180).to(device)
    kp = torch.zeros(H, W, device=device)
    kp[:25, :25] = k
    kp = torch.roll(kp, (-12, -12), dims=(0, 1))
    K = torch.fft.rfft2(kp)
    out = torch.empty(n, C, H, W, dtype=torch.float32, device=device)
    for i in range(n):
        g = torch.Generator().manual_seed(base_seed + first_index + i)
        small = torch.rand(C, -(-H // block), -(-W // block), generator=g).to(device)
        img = small.repeat_interleave(block, -2).repeat_interleave(block, -1)[..., :H, :W]
        out[i] = torch.fft.irfft2(torch.fft.rfft2(img) * K, s=(H, W)).clamp_(0, 1)
    return out


def make(kind, n, C, H, W, first_index=0, base_seed=0, device="cpu"):
    if kind == "white":
        return white(n, C, H, W, first_index, base_seed, device)
    if kind == "mosaic":
        return mosaic(n, C, H, W, first_index, base_seed, device)
    raise ValueError(f"unknown synthetic distribution {kind!r}")
