"""torch-CPU restatement of the reference's default path -- TEST / BASELINE INFRASTRUCTURE.

Same algorithm as ``oracle/polyblur_oracle.py`` (which is the parity checker), but written
with the ATen CPU ops the reference itself dispatches to (torch.fft, amax, matmul), so that
its *speed* is representative of the reference when ``bench.py`` times the CPU baseline
(`cpu_baseline` / `--impl reference`, kind "port").  The numpy oracle is 3x slower than the
reference because scipy's pocketfft is slower than ATen's; timing it would flatter the GPU.

Only ``bench.py``'s CPU legs and ``tests/`` import this module.  Pinned against the golden
vectors of the live reference in tests/test_oracle_golden.py.
File:line citations are into the reference checkout (polyblur/...).
"""
from __future__ import annotations

import math

import torch


def fourier_gradients(x):
    """filters.py:159-186: fft2, shift, multiply by i*2*pi*f, unshift, real(ifft2)."""
    h, w = x.shape[-2:]
    U = torch.fft.fftshift(torch.fft.fft2(x), dim=(-2, -1))
    fh = ((torch.arange(h) - h // 2) / h).view(-1, 1)
    fw = ((torch.arange(w) - w // 2) / w).view(1, -1)
    rot = torch.complex(-U.imag, U.real)
    gx = torch.fft.ifft2(torch.fft.ifftshift(2 * math.pi * fw * rot, dim=(-2, -1))).real
    gy = torch.fft.ifft2(torch.fft.ifftshift(2 * math.pi * fh * rot, dim=(-2, -1))).real
    return gx, gy


_ANGLES = torch.linspace(0, math.pi, 7).view(1, 7, 1, 1)


def _keys_matrix():
    xs = torch.linspace(0, 180, 7).long() / 30
    xn = torch.arange(0, 180, 6).long() / 30
    d = (xn[:, None] - xs[None, :]).abs()
    near = (d < 1).float()
    far = ((d >= 1) & (d < 2)).float()
    w = far * (((-0.5 * d + 2.5) * d - 4) * d + 2) + near * ((1.5 * d - 2.5) * d * d + 1)
    return w / (w.sum(-1, keepdim=True) + 1e-5)


_KEYS = _keys_matrix()


def estimate(img, c, b):
    """blur_estimation.py:18-79 (q = 0, gray path) -> (B,1,25,25) kernels."""
    g = img.mean(dim=1, keepdim=True)
    lo = g.amin(dim=(-1, -2), keepdim=True)
    hi = g.amax(dim=(-1, -2), keepdim=True)
    g = ((g - lo) / (hi - lo)).clamp(0, 1)
    gx, gy = fourier_gradients(g)
    mags = (torch.cos(_ANGLES) * gx - torch.sin(_ANGLES) * gy).abs().amax(dim=(-1, -2))     # (B,7)
    interp = (_KEYS[None] @ mags[..., None]).squeeze(-1)                                      # (B,30)
    imin = interp.argmin(dim=-1, keepdim=True)
    deg = 6 * imin
    iort = ((deg + 90) % 180) // 6
    m_n = interp.gather(-1, imin)
    m_o = interp.gather(-1, iort)
    sigma = (c * c / (m_n * m_n + 1e-8) - b * b).clamp(0.09, 16.0).sqrt()
    rho = (c * c / (m_o * m_o + 1e-8) - b * b).clamp(0.09, 16.0).sqrt()
    theta = -(deg.float() * math.pi / 180)
    cs, sn = torch.cos(theta), torch.sin(theta)
    i1, i2 = 1.0 / (sigma * sigma), 1.0 / (rho * rho)
    a00 = (cs * cs * i1 + sn * sn * i2).view(-1, 1, 1)
    a01 = (sn * cs * (i1 - i2)).view(-1, 1, 1)
    a11 = (cs * cs * i2 + sn * sn * i1).view(-1, 1, 1)
    t = (torch.arange(25) - 12).float()
    Y, X = torch.meshgrid(t, t, indexing="ij")
    quad = X * (a00 * X + a01 * Y) + Y * (a01 * X + a11 * Y)
    k = torch.exp(-0.5 * quad)
    k = k / k.sum(dim=(-1, -2), keepdim=True)
    return k[:, None], dict(mags=mags, theta_deg=deg[:, 0], sigma=sigma[:, 0], rho=rho[:, 0])


def deconvolve(img, kernel, alpha, beta):
    """deblurring.py:211-239 + :141-169 + filters.py:255-273, default flags."""
    pad = kernel.shape[-1] // 2
    p = torch.nn.functional.pad(img, (pad, pad, pad, pad), mode="replicate")
    h, w = p.shape[-2:]
    otf = torch.zeros(kernel.shape[0], 1, h, w)
    otf[..., :kernel.shape[-2], :kernel.shape[-1]] = kernel
    K = torch.fft.fft2(torch.roll(otf, (-pad, -pad), dims=(-2, -1)))
    Y = torch.fft.fft2(p)
    a3, a2, a1 = alpha / 2 - beta + 2, 3 * beta - alpha - 6, 5 - 3 * beta + alpha / 2
    X = a3 * Y
    X = K * X + a2 * Y
    X = K * X + a1 * Y
    X = K * X + beta * Y
    out = torch.fft.ifft2(X).real[..., pad:-pad, pad:-pad]
    return out.clamp(0.0, 1.0)


def polyblur_deblurring(img, n_iter=1, c=0.352, b=0.768, alpha=2, beta=3, faithful_cost=True, trace=None):
    """deblurring.py:23-96 on a (B,C,H,W) float32 CPU tensor, default flags, method='fft'.
    ``faithful_cost`` also evaluates the reference's unconditional, unused init gradient (:61)."""
    if faithful_cost:
        fourier_gradients(img)
    cur = img
    for _ in range(n_iter):
        k, tr = estimate(cur, c, b)
        if trace is not None:
            trace.append(tr)
        cur = deconvolve(cur, k, alpha, beta).clip(0.0, 1.0)
    return cur
