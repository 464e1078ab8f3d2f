"""CPU oracle for the Polyblur hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

This file is a numpy / scipy.fft restatement of the algorithm that teboli/polyblur
runs for ``polyblur_deblurring(..., method='fft')`` (the reference's default and only
batched-correct path).  It exists so that the CUDA engine in
``polyblur_b200/`` can be checked against something that runs on any CPU box.

Rules (tier framing, item 3):
  * only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
    ``--impl reference`` legs may import this module;
  * nothing under ``polyblur_b200/`` imports it -- the product fails loudly when the
    CUDA extension is missing instead of falling back to this code.

Parity pin: the reference ships no tests and no reproducible golden output
(SURVEY.md section 4), so the oracle is pinned against outputs of the *live* reference
generated in the build container by ``tests/golden/make_golden.py`` (committed with
its ``.npz`` outputs).  ``tests/test_oracle_golden.py`` checks every function below
against those vectors.

Every function cites the reference file:line it follows (paths relative to the
reference checkout).  All functions take / return ``(B, C, H, W)`` numpy arrays and a
``dtype`` (float32 = the reference's arithmetic, float64 = the "truth" used to
measure the fp32 noise floor).
"""
from __future__ import annotations

import math
import os

import numpy as np
from scipy import fft as _fft

__all__ = [
    "to_tensor", "to_array", "pad_with_kernel", "crop_with_kernel",
    "fourier_gradients", "fourier_gradients_1d", "gray_mean", "saturation_mask",
    "normalize", "gradient_magnitudes", "keys_weights", "find_direction",
    "gaussian_parameters", "gaussian_kernel", "gaussian_blur_estimation",
    "polynomial_coefficients", "compute_polynomial_fft", "compute_polynomial_torus",
    "inverse_filtering_rank3", "halo_masking", "edgetaper_alpha", "edgetaper",
    "convolve2d_fft", "bilateral_filter", "recursive_filter", "normalized_convolution",
    "polyblur_deblurring", "gaussian_filter_np",
]

_WORKERS = int(os.environ.get("PB_ORACLE_WORKERS", "0")) or (os.cpu_count() or 1)


def set_workers(n: int) -> None:
    """Number of pocketfft threads (the CPU baseline states it as ``cores``)."""
    global _WORKERS
    _WORKERS = max(1, int(n))


def get_workers() -> int:
    return _WORKERS


def _cdt(dtype):
    return np.complex64 if np.dtype(dtype) == np.float32 else np.complex128


# --------------------------------------------------------------------------------------
# utils.py
# --------------------------------------------------------------------------------------
def to_tensor(x: np.ndarray) -> np.ndarray:
    """(H,W) or (H,W,C) ndarray -> (C,H,W) float32, no rescaling (utils.py:8-21)."""
    x = np.asarray(x)
    if x.ndim == 2:
        x = x[None]
    else:
        x = np.transpose(x, (2, 0, 1))
    return np.ascontiguousarray(x).astype(np.float32)


def to_array(x: np.ndarray) -> np.ndarray:
    """(..,C,H,W) -> squeezed (H,W) or (H,W,C) ndarray (utils.py:24-31)."""
    x = np.squeeze(x)
    if x.ndim == 2:
        return x
    return np.transpose(x, (1, 2, 0))


def pad_with_kernel(img, ks, mode="edge"):
    """Replicate ('edge') or circular ('wrap') pad by ks (utils.py:48-53)."""
    return np.pad(img, ((0, 0), (0, 0), (ks, ks), (ks, ks)), mode=mode)


def crop_with_kernel(img, ks):
    """utils.py:56-61."""
    return img[..., ks:-ks, ks:-ks]


# --------------------------------------------------------------------------------------
# filters.py : spectral gradient
# --------------------------------------------------------------------------------------
def _shifted_freq(n, dtype):
    # filters.py:175-176: (arange(n) - n//2) / n, evaluated in the working precision
    # (the reference divides an int64 tensor by a Python int -> float32).
    return ((np.arange(n) - n // 2).astype(dtype) / dtype(n)).astype(dtype)


def fourier_gradients(images, dtype=np.float32):
    """Spectral derivative along W (gx) and H (gy); filters.py:159-186.

    Follows the reference's operation order: full complex fft2, fftshift, multiply by
    2*pi*f*(-Im + i*Re), ifftshift, real(ifft2).
    """
    dtype = np.dtype(dtype).type
    x = np.asarray(images, dtype=dtype)
    h, w = x.shape[-2:]
    U = _fft.fft2(x.astype(_cdt(dtype)), axes=(-2, -1), workers=_WORKERS)
    U = _fft.fftshift(U, axes=(-2, -1))
    fh = _shifted_freq(h, dtype)[:, None]
    fw = _shifted_freq(w, dtype)[None, :]
    two_pi = dtype(2 * np.pi)
    rot = (-U.imag + 1j * U.real).astype(_cdt(dtype))          # = i * U
    gxU = _fft.ifftshift((two_pi * fw) * rot, axes=(-2, -1))
    gx = _fft.ifft2(gxU, axes=(-2, -1), workers=_WORKERS).real.astype(dtype)
    gyU = _fft.ifftshift((two_pi * fh) * rot, axes=(-2, -1))
    gy = _fft.ifft2(gyU, axes=(-2, -1), workers=_WORKERS).real.astype(dtype)
    return gx, gy


def fourier_gradients_1d(images, dtype=np.float32):
    """Same operator written as independent 1-D row / column spectral derivatives.

    This is the form the CUDA kernels use: the Nyquist bin of
    an even length contributes nothing to the real part, so it is zeroed explicitly.
    Agreement with :func:`fourier_gradients` is checked in tests (1e-6 in fp32).
    """
    dtype = np.dtype(dtype).type
    x = np.asarray(images, dtype=dtype)
    out = []
    for axis in (-1, -2):
        n = x.shape[axis]
        k = np.arange(n)
        f = np.where(k < (n + 1) // 2, k, k - n).astype(np.float64) / n
        if n % 2 == 0:
            f[n // 2] = 0.0
        mult = (2j * np.pi * f).astype(_cdt(dtype))
        shape = [1] * x.ndim
        shape[axis] = n
        U = _fft.fft(x.astype(_cdt(dtype)), axis=axis, workers=_WORKERS)
        g = _fft.ifft(U * mult.reshape(shape), axis=axis, workers=_WORKERS).real
        out.append(g.astype(dtype))
    return out[0], out[1]


# --------------------------------------------------------------------------------------
# blur_estimation.py
# --------------------------------------------------------------------------------------
def gray_mean(img, dtype=np.float32):
    """Channel mean, kept as (B,1,H,W) (blur_estimation.py:36-37).

    torch's mean over a size-3 dim is ((c0 + c1) + c2) / 3 in the working precision.
    """
    dtype = np.dtype(dtype).type
    x = np.asarray(img, dtype=dtype)
    acc = x[:, 0].copy()
    for c in range(1, x.shape[1]):
        acc = acc + x[:, c]
    return (acc / dtype(x.shape[1]))[:, None]


def saturation_mask(gray, discard_saturation, threshold=0.99):
    """blur_estimation.py:83-88 (mask is taken on the *un-normalised* gray)."""
    if discard_saturation:
        return gray > threshold
    return np.zeros(gray.shape, dtype=bool)


def _quantile_linear(flat, q, dtype):
    # torch.quantile(..., interpolation='linear') on the sorted values:
    # pos = q*(n-1); lerp(v[floor], v[ceil], frac) evaluated in the working precision.
    n = flat.shape[-1]
    srt = np.sort(flat, axis=-1)
    pos = dtype(q) * dtype(n - 1)
    lo = int(np.floor(pos))
    hi = min(lo + 1, n - 1)
    frac = dtype(pos - dtype(lo))
    a = srt[..., lo]
    b = srt[..., hi]
    return (a + frac * (b - a)).astype(dtype)


def normalize(gray, q=0.0, dtype=np.float32):
    """clamp((x - lo) / (hi - lo), 0, 1) with per-image min/max or quantiles.

    blur_estimation.py:96-109 and :92-93.  A constant image divides by zero exactly
    as in the reference (NaN output).
    """
    dtype = np.dtype(dtype).type
    g = np.asarray(gray, dtype=dtype)
    b, c = g.shape[:2]
    if q > 0:
        flat = g.reshape(b, c, -1)
        lo = _quantile_linear(flat, q, dtype)[..., None, None]
        hi = _quantile_linear(flat, 1 - q, dtype)[..., None, None]
    else:
        lo = g.min(axis=(-1, -2), keepdims=True)
        hi = g.max(axis=(-1, -2), keepdims=True)
    with np.errstate(divide="ignore", invalid="ignore"):
        out = (g - lo) / (hi - lo)
    return np.clip(out, dtype(0), dtype(1)).astype(dtype)


def direction_angles(dtype=np.float32, n_angles=6):
    """linspace(0, pi, n_angles+1) in the working precision (blur_estimation.py:129)."""
    dtype = np.dtype(dtype).type
    return np.linspace(0, np.pi, n_angles + 1).astype(dtype)


def gradient_magnitudes(gx, gy, dtype=np.float32):
    """max over pixels of |cos(phi_j) gx - sin(phi_j) gy| for 7 angles -> (B,7).

    blur_estimation.py:122-134 (the channel mean there is a no-op: C == 1 here).
    """
    dtype = np.dtype(dtype).type
    ang = direction_angles(dtype)
    cs = np.cos(ang).astype(dtype)
    sn = np.sin(ang).astype(dtype)
    gxm = gx.mean(axis=1) if gx.shape[1] > 1 else gx[:, 0]
    gym = gy.mean(axis=1) if gy.shape[1] > 1 else gy[:, 0]
    out = np.empty((gx.shape[0], ang.size), dtype=dtype)
    for j in range(ang.size):
        out[:, j] = np.abs(cs[j] * gxm - sn[j] * gym).max(axis=(-1, -2))
    return out


def keys_weights(dtype=np.float32, n_angles=6, n_interp=30):
    """(30,7) matrix of normalised Keys cubic weights (blur_estimation.py:138-148,157-158).

    Sample positions are theta_j / 30 with theta_j = long(linspace(0,180,7)) and query
    positions long(arange(0,180,6)) / 30, both divided in float32 by the reference.
    The normaliser carries the reference's +1e-5.
    """
    dtype = np.dtype(dtype).type
    thetas = np.linspace(0, 180, n_angles + 1).astype(np.int64)
    interp = np.arange(0, 180, 180 / n_interp).astype(np.int64)
    x = (thetas.astype(dtype) / dtype(n_interp))
    xn = (interp.astype(dtype) / dtype(n_interp))
    d = np.abs(xn[:, None] - x[None, :]).astype(dtype)
    m1 = (d < 1).astype(dtype)
    m2 = ((d >= 1) & (d < 2)).astype(dtype)
    w = m2 * (((dtype(-0.5) * d + dtype(2.5)) * d - dtype(4)) * d + dtype(2)) \
        + m1 * ((dtype(1.5) * d - dtype(2.5)) * d * d + dtype(1))
    w = w / (w.sum(axis=-1, keepdims=True) + dtype(1e-5))
    return w.astype(dtype), interp


def find_direction(mags, dtype=np.float32):
    """Interpolate 7 -> 30 magnitudes, argmin -> blur direction.

    blur_estimation.py:151-167.  Returns (m_normal, m_ortho, theta_rad, theta_deg, interp).
    """
    dtype = np.dtype(dtype).type
    w, interp_deg = keys_weights(dtype)
    # (x_new @ y[..., None]) : float32 matmul, accumulation order j = 0..6
    interp = np.zeros((mags.shape[0], w.shape[0]), dtype=dtype)
    for j in range(w.shape[1]):
        interp = interp + w[None, :, j] * mags[:, j:j + 1].astype(dtype)
    i_min = np.argmin(interp, axis=-1)                       # first occurrence
    theta_deg = interp_deg[i_min]
    m_normal = np.take_along_axis(interp, i_min[:, None], axis=-1)[:, 0]
    theta_ortho = (theta_deg + 90) % 180
    i_ortho = (theta_ortho / (180 / w.shape[0])).astype(np.int64)
    m_ortho = np.take_along_axis(interp, i_ortho[:, None], axis=-1)[:, 0]
    theta = (theta_deg.astype(np.float32).astype(np.float64) * np.pi / 180).astype(dtype)
    return m_normal, m_ortho, theta, theta_deg, interp


def gaussian_parameters(m_normal, m_ortho, c, b, dtype=np.float32):
    """sigma, rho from the affine model, clamped to [0.3, 4] (blur_estimation.py:171-185)."""
    dtype = np.dtype(dtype).type
    cc = c * c
    bb = b * b

    def one(m):
        m = m.astype(dtype)
        v = (dtype(cc) / (m * m + dtype(1e-8)) - dtype(bb)).astype(dtype)
        return np.sqrt(np.clip(v, dtype(0.09), dtype(16.0))).astype(dtype)

    return one(m_normal), one(m_ortho)


def gaussian_kernel(theta, sigma, rho, ksize=25, dtype=np.float32):
    """(B,1,k,k) normalised anisotropic Gaussian (blur_estimation.py:189-232).

    x runs along columns, y along rows (meshgrid 'xy'); the rotation uses -theta.
    """
    dtype = np.dtype(dtype).type
    th = (-np.asarray(theta, dtype=dtype)).astype(dtype)
    c = np.cos(th).astype(dtype)
    s = np.sin(th).astype(dtype)
    cc, ss, sc = c * c, s * s, s * c
    il1 = (dtype(1.0) / (sigma.astype(dtype) * sigma.astype(dtype))).astype(dtype)
    il2 = (dtype(1.0) / (rho.astype(dtype) * rho.astype(dtype))).astype(dtype)
    a00 = (cc * il1 + ss * il2).astype(dtype)
    a01 = (sc * (il1 - il2)).astype(dtype)
    a11 = (cc * il2 + ss * il1).astype(dtype)
    t = (np.arange(ksize) - (ksize - 1) // 2).astype(dtype)
    X, Y = np.meshgrid(t, t, indexing="xy")
    X = X[None].astype(dtype)
    Y = Y[None].astype(dtype)
    # Z^T (S^-1 Z):  u = a00 x + a01 y ; v = a01 x + a11 y ; q = x u + y v
    u = a00[:, None, None] * X + a01[:, None, None] * Y
    v = a01[:, None, None] * X + a11[:, None, None] * Y
    quad = (X * u + Y * v).astype(dtype)
    k = np.exp(dtype(-0.5) * quad).astype(dtype)
    k = k / k.sum(axis=(-1, -2), keepdims=True)
    return k[:, None].astype(dtype)


def gaussian_blur_estimation(img, c=0.362, b=0.464, q=0.0001, ker_size=25,
                             discard_saturation=False, dtype=np.float32, trace=None):
    """One blur estimate per image (blur_estimation.py:18-79), gray path only.

    ``multichannel=True`` is a no-op for RGB and crashes for other C in the reference
    (SURVEY.md Appendix B.9), so only the channel-mean path exists here.
    """
    gray = gray_mean(img, dtype)
    mask = saturation_mask(gray, discard_saturation)
    gn = normalize(gray, q=q, dtype=dtype)
    gx, gy = fourier_gradients(gn, dtype)
    gx = np.where(mask, 0, gx).astype(gx.dtype)
    gy = np.where(mask, 0, gy).astype(gy.dtype)
    mags = gradient_magnitudes(gx, gy, dtype)
    m_n, m_o, theta, theta_deg, interp = find_direction(mags, dtype)
    sigma, rho = gaussian_parameters(m_n, m_o, c, b, dtype)
    kernel = gaussian_kernel(theta, sigma, rho, ker_size, dtype)
    if trace is not None:
        trace.append(dict(mags=mags, interp=interp, m_normal=m_n, m_ortho=m_o,
                          theta=theta, theta_deg=theta_deg, sigma=sigma, rho=rho,
                          kernel=kernel))
    return kernel


# --------------------------------------------------------------------------------------
# deblurring.py : polynomial deconvolution
# --------------------------------------------------------------------------------------
def polynomial_coefficients(alpha, beta):
    """a3, a2, a1, b of the degree-3 approximate inverse (deblurring.py:160-162)."""
    a3 = alpha / 2 - beta + 2
    a2 = 3 * beta - alpha - 6
    a1 = 5 - 3 * beta + alpha / 2
    return a3, a2, a1, beta


def _p2o(kernel, shape, dtype):
    """Zero-embed, roll by -k//2, fft2 (filters.py:255-273)."""
    kh, kw = kernel.shape[-2:]
    otf = np.zeros(kernel.shape[:-2] + tuple(shape), dtype=dtype)
    otf[..., :kh, :kw] = kernel
    otf = np.roll(otf, (-int(kh / 2), -int(kw / 2)), axis=(-2, -1))
    return _fft.fft2(otf.astype(_cdt(dtype)), axes=(-2, -1), workers=_WORKERS)


def compute_polynomial_fft(img, kernel, alpha, beta, dtype=np.float32):
    """Horner evaluation of the polynomial in the Fourier domain (deblurring.py:141-169)."""
    dtype = np.dtype(dtype).type
    cdt = _cdt(dtype)
    h, w = img.shape[-2:]
    Y = _fft.fft2(np.asarray(img, dtype=dtype).astype(cdt), axes=(-2, -1), workers=_WORKERS)
    K = _p2o(np.asarray(kernel, dtype=dtype), (h, w), dtype)
    a3, a2, a1, b = (dtype(v) for v in polynomial_coefficients(alpha, beta))
    X = a3 * Y
    X = K * X + a2 * Y
    X = K * X + a1 * Y
    X = K * X + b * Y
    return _fft.ifft2(X.astype(cdt), axes=(-2, -1), workers=_WORKERS).real.astype(dtype)


def torus_source_index(n, pad, coords):
    """Gather map of SURVEY.md Appendix A.6: padded (possibly out-of-range) coordinate
    -> source row/column of the *unpadded* image: clamp((c mod (n+2 pad)) - pad, 0, n-1)."""
    return np.clip(np.mod(coords, n + 2 * pad) - pad, 0, n - 1)


def compute_polynomial_torus(img, kernel, alpha, beta, dtype=np.float64):
    """Spatial-domain twin of :func:`compute_polynomial_fft` on the *padded* image:
    three circular convolutions on the (H',W') torus, Horner order: out[p] = sum_d K[d] v[p - d],
    which is what the p2o / fft2 product computes (filters.py:255-273), also for kernels that
    are not point-symmetric (pinned by tests/golden/round2.npz "asym/*").  O(625 H W) per
    step -- small inputs only.  Used to validate the gather map the CUDA kernel uses.
    """
    dtype = np.dtype(dtype).type
    p = np.asarray(img, dtype=dtype)
    k = np.asarray(kernel, dtype=dtype)
    r = k.shape[-1] // 2
    a3, a2, a1, b = (dtype(v) for v in polynomial_coefficients(alpha, beta))

    def corr(v):          # circular convolution with k
        out = np.zeros_like(v)
        for dy in range(-r, r + 1):
            for dx in range(-r, r + 1):
                out += k[:, :, dy + r, dx + r][..., None, None] * np.roll(v, (dy, dx), axis=(-2, -1))
        return out

    o = a3 * p
    o = corr(o) + a2 * p
    o = corr(o) + a1 * p
    o = corr(o) + b * p
    return o.astype(dtype)


def halo_masking(img, imout, grad_img, dtype=np.float32):
    """Bug-compatible halo masking (deblurring.py:173-208): M uses gy*gy, not gy*goy."""
    dtype = np.dtype(dtype).type
    # grad_img defaults to the gradients of the image it is handed (deblurring.py:200-203)
    gx, gy = fourier_gradients(img, dtype) if grad_img is None else grad_img
    ox, oy = fourier_gradients(imout, dtype)
    M = (-gx * ox) + (-gy * gy)
    nM = np.sum(gx * gx + gy * gy, axis=(-2, -1), keepdims=True, dtype=dtype)
    with np.errstate(divide="ignore", invalid="ignore"):
        z = np.maximum(M / (nM + M), dtype(0))
    return (imout + z * (img - imout)).astype(dtype)


def convolve2d_fft(img, kernel, dtype=np.float32):
    """filters.convolve2d(method='fft') (filters.py:31-35): circular pad, FFT product, crop."""
    dtype = np.dtype(dtype).type
    ks = kernel.shape[-1] // 2
    xp = pad_with_kernel(np.asarray(img, dtype=dtype), ks, mode="wrap")
    X = _fft.fft2(xp.astype(_cdt(dtype)), axes=(-2, -1), workers=_WORKERS)
    K = _p2o(np.asarray(kernel, dtype=dtype), xp.shape[-2:], dtype)
    y = _fft.ifft2(K * X, axes=(-2, -1), workers=_WORKERS).real.astype(dtype)
    return crop_with_kernel(y, ks)


def edgetaper_alpha(kernel, img_shape, dtype=np.float32):
    """Taper weights from the kernel's projected autocorrelations (edgetaper.py:10-23).

    The max is over the whole batch, as in the reference.
    """
    dtype = np.dtype(dtype).type
    k = np.asarray(kernel, dtype=dtype)
    vs = []
    for axis_sum, n in ((-1, img_shape[0] - 1), (-2, img_shape[1] - 1)):
        proj = k.sum(axis=axis_sum).astype(dtype)
        z = _fft.fft(proj.astype(_cdt(dtype)), n=n, axis=-1)
        za = np.abs(z).astype(dtype)
        z = _fft.ifft((za * za).astype(_cdt(dtype)), axis=-1).real.astype(dtype)
        z = np.concatenate([z, z[..., 0:1]], axis=-1)
        vs.append((dtype(1) - z / z.max()).astype(dtype))
    return (vs[0][..., :, None] * vs[1][..., None, :]).astype(dtype)


def edgetaper(img, kernel, n_tapers=3, dtype=np.float32):
    """edgetaper.py:26-33 with method='fft'."""
    dtype = np.dtype(dtype).type
    img = np.asarray(img, dtype=dtype)
    alpha = edgetaper_alpha(kernel, img.shape[-2:], dtype)
    for _ in range(n_tapers):
        blurred = convolve2d_fft(img, kernel, dtype)
        img = (alpha * img + (dtype(1.0) - alpha) * blurred).astype(dtype)
    return img


def inverse_filtering_rank3(img, kernel, alpha=2, b=4, remove_halo=False, do_edgetaper=False,
                            grad_img=None, dtype=np.float32, spatial=False):
    """pad -> [edgetaper] -> polynomial -> crop -> [halo] -> clamp (deblurring.py:211-239)."""
    dtype = np.dtype(dtype).type
    ks = kernel.shape[-1] // 2
    p = pad_with_kernel(np.asarray(img, dtype=dtype), ks)
    if do_edgetaper:
        p = edgetaper(p, kernel, dtype=dtype)
    if spatial:
        o = compute_polynomial_torus(p, kernel, alpha, b, dtype)
    else:
        o = compute_polynomial_fft(p, kernel, alpha, b, dtype)
    o = crop_with_kernel(o, ks)
    if remove_halo:
        cur = crop_with_kernel(p, ks)
        o = halo_masking(cur, o, grad_img, dtype)
    return np.clip(o, dtype(0), dtype(1)).astype(dtype)


def inverse_filtering_rank3_vjp(img, kernel, grad_out, alpha=2, b=4, dtype=np.float64):
    """What torch.autograd computes over the reference's inverse_filtering_rank3 (default flags,
    deblurring.py:211-239 with utils.py:48-61 and deblurring.py:141-169) for the scalar
    <grad_out, output>: the gradients with respect to ``img`` and to the kernel taps.

    Forward: y = clamp(C T R x); R replicate pad, T circular filter with the spectrum
    P(K^) = ((a3 K^ + a2) K^ + a1) K^ + b on the padded torus, C crop.  Backward: the clamp passes
    the gradient where the unclamped value lies in [0, 1] (torch.clamp), C^T zero-embeds, T^T
    multiplies by conj P(K^), R^T folds the border back (sums the padded rows / columns that
    replicate an edge pixel).  Kernel: K~[d] = sum_p z[p] v[p - d] with v = P'(K) (*) R x, i.e. the
    circular cross-correlation of z and v read at the kernel's offsets.
    Returns (grad_img (B,C,H,W), grad_kernel (B,1,k,k), unclamped forward result)."""
    dtype = np.dtype(dtype).type
    cdt = _cdt(dtype)
    img = np.asarray(img, dtype=dtype)
    kernel = np.asarray(kernel, dtype=dtype)
    ksz = kernel.shape[-1]
    pad = ksz // 2
    B, C, H, W = img.shape
    xp = pad_with_kernel(img, pad)
    Hp, Wp = xp.shape[-2:]
    a3, a2, a1, b0 = (dtype(v) for v in polynomial_coefficients(alpha, b))
    K = _p2o(np.broadcast_to(kernel, (B, 1, ksz, ksz)), (Hp, Wp), dtype)
    P = ((a3 * K + a2) * K + a1) * K + b0
    dP = (3 * a3 * K + 2 * a2) * K + a1
    X = _fft.fft2(xp.astype(cdt), axes=(-2, -1), workers=_WORKERS)
    pre = crop_with_kernel(_fft.ifft2(P * X, axes=(-2, -1), workers=_WORKERS).real, pad)
    z = np.zeros((B, C, Hp, Wp), dtype=dtype)
    z[..., pad:pad + H, pad:pad + W] = np.asarray(grad_out, dtype=dtype) * ((pre >= 0) & (pre <= 1))
    Z = _fft.fft2(z.astype(cdt), axes=(-2, -1), workers=_WORKERS)
    t = _fft.ifft2(np.conj(P) * Z, axes=(-2, -1), workers=_WORKERS).real
    # R^T: interior pixels map to one padded position, edge pixels collect their replicas
    g = t[..., pad:pad + H, :].copy()
    g[..., 0, :] += t[..., :pad, :].sum(axis=-2)
    g[..., H - 1, :] += t[..., pad + H:, :].sum(axis=-2)
    gi = g[..., pad:pad + W].copy()
    gi[..., 0] += g[..., :pad].sum(axis=-1)
    gi[..., W - 1] += g[..., pad + W:].sum(axis=-1)
    # kernel taps: correlation of z with v = P'(K) (*) xp at offsets d = index - pad
    corr = _fft.ifft2(Z * np.conj(dP * X), axes=(-2, -1), workers=_WORKERS).real.sum(axis=1)
    idx = (np.arange(ksz) - pad)
    gk = corr[:, idx % Hp][:, :, idx % Wp][:, None]
    return gi.astype(dtype), gk.astype(dtype), pre.astype(dtype)


# --------------------------------------------------------------------------------------
# optional prefilters
# --------------------------------------------------------------------------------------
def bilateral_filter(img, ksize=5, sigma_spatial=5.0, sigma_color=0.1, dtype=np.float32):
    """5x5 bilateral filter with per-channel range weights (filters.py:107-148)."""
    dtype = np.dtype(dtype).type
    I = np.asarray(img, dtype=dtype)
    r = ksize // 2
    t = np.arange(-ksize // 2 + 1, ksize // 2 + 1)
    xx, yy = np.meshgrid(t, t, indexing="xy")
    gw = np.exp(-(xx * xx + yy * yy).astype(dtype) / dtype(2 * sigma_spatial * sigma_spatial)).astype(dtype)
    P = pad_with_kernel(I, r)
    h, w = I.shape[-2:]
    var2 = dtype(2 * sigma_color * sigma_color)
    J = np.zeros_like(I)
    W = np.zeros_like(I)
    for y in range(ksize):
        Jy = np.zeros_like(I)
        Wy = np.zeros_like(I)
        for x in range(ksize):
            S = P[..., y:y + h, x:x + w]
            F = S - I
            F = np.exp(-F * F / var2).astype(dtype) * gw[y, x]
            Jy += F * S
            Wy += F
        J += Jy
        W += Wy
    return (J / (W + dtype(1e-5))).astype(dtype)


def recursive_filter(img, sigma_s=60, sigma_r=0.4, num_iterations=3, joint_image=None,
                     dtype=np.float32):
    """Gastal-Oliveira domain-transform recursive filter (domain_transform.py:6-85)."""
    dtype = np.dtype(dtype).type
    I = np.asarray(img, dtype=dtype)
    J = I if joint_image is None else np.asarray(joint_image, dtype=dtype)
    dIdx = np.zeros((J.shape[0],) + J.shape[-2:], dtype=dtype)
    dIdy = np.zeros_like(dIdx)
    dIdx[..., :, 1:] = np.abs(np.diff(J, axis=-1)).sum(axis=1, dtype=dtype)
    dIdy[..., 1:, :] = np.abs(np.diff(J, axis=-2)).sum(axis=1, dtype=dtype)
    ratio = dtype(sigma_s / sigma_r)
    dHdx = (dtype(1) + ratio * dIdx).astype(dtype)
    dVdy = (dtype(1) + ratio * dIdy).astype(dtype)
    F = I.copy()
    N = num_iterations

    def sweep(F, V, axis):
        # V: (B,H,W) feedback weights; recurrence along `axis` (-1 rows, -2 columns)
        Fm = np.moveaxis(F, axis, -1)
        Vm = np.moveaxis(V[:, None], axis, -1)
        n = Fm.shape[-1]
        for i in range(1, n):
            Fm[..., i] += Vm[..., i] * (Fm[..., i - 1] - Fm[..., i])
        for i in range(n - 2, -1, -1):
            Fm[..., i] += Vm[..., i + 1] * (Fm[..., i + 1] - Fm[..., i])
        return F

    for i in range(N):
        sigma_i = sigma_s * math.sqrt(3) * 2 ** (N - (i + 1)) / math.sqrt(4 ** N - 1)
        a = math.exp(-math.sqrt(2) / sigma_i)
        F = sweep(F, np.power(dtype(a), dHdx).astype(dtype), -1)
        F = sweep(F, np.power(dtype(a), dVdy).astype(dtype), -2)
    return F.astype(dtype)


def normalized_convolution(img, sigma_s=60, sigma_r=0.4, num_iterations=3, dtype=np.float32):
    """Domain-transform normalized convolution as written in NC.cpp:143-204 / :50-140
    (searchsorted restatement of SURVEY.md Appendix A.11)."""
    dtype = np.dtype(dtype).type
    I = np.asarray(img, dtype=dtype)
    B, C, H, W = I.shape
    dIdx = np.zeros((B, H, W), dtype=dtype)
    dIdy = np.zeros((B, H, W), dtype=dtype)
    dIdx[..., :, 1:] = np.abs(np.diff(I, axis=-1)).sum(axis=1, dtype=dtype)
    dIdy[..., 1:, :] = np.abs(np.diff(I, axis=-2)).sum(axis=1, dtype=dtype)
    ratio = dtype(sigma_s / sigma_r)
    ctH = np.cumsum(dtype(1) + ratio * dIdx, axis=-1, dtype=dtype)
    ctV = np.cumsum(dtype(1) + ratio * dIdy, axis=-2, dtype=dtype)
    F = I.copy()
    N = num_iterations

    def box(F, ct, radius):
        # rows of F: (B,C,n_rows,n); ct: (B,n_rows,n)
        out = np.empty_like(F)
        n = F.shape[-1]
        for b in range(F.shape[0]):
            for r in range(F.shape[2]):
                row = ct[b, r]
                lo = np.searchsorted(row, row - radius, side="right")
                hi = np.searchsorted(row, row + radius, side="right")
                for ch in range(F.shape[1]):
                    sat = np.concatenate([[dtype(0)], np.cumsum(F[b, ch, r], dtype=dtype)])
                    out[b, ch, r] = (sat[hi] - sat[lo]) / ((hi - lo).astype(dtype) + dtype(1e-4))
        return out

    for i in range(N):
        sigma_i = sigma_s * math.sqrt(3) * 2 ** (N - (i + 1)) / math.sqrt(4 ** N - 1)
        radius = dtype(math.sqrt(3) * sigma_i)
        F = box(F, ctH, radius)
        Ft = np.ascontiguousarray(np.swapaxes(F, -1, -2))
        Ft = box(Ft, np.ascontiguousarray(np.swapaxes(ctV, -1, -2)), radius)
        F = np.ascontiguousarray(np.swapaxes(Ft, -1, -2))
    return F.astype(dtype)


# --------------------------------------------------------------------------------------
# deblurring.py : driver
# --------------------------------------------------------------------------------------
def polyblur_deblurring(img, n_iter=1, c=0.352, b=0.768, alpha=2, beta=3, sigma_r=0.8,
                        sigma_s=2.0, ker_size=25, q=0.0, remove_halo=False, edgetaping=False,
                        prefiltering=False, discard_saturation=False, prefilter="bilateral",
                        dtype=np.float32, trace=None, faithful_cost=False):
    """Outer loop of the reference (deblurring.py:23-96), method='fft' semantics.

    ``img``: (H,W)/(H,W,C) ndarray (returns the squeezed ndarray form, like the
    reference's ndarray path) or (B,C,H,W) ndarray (returns the same layout).
    ``prefilter``: 'bilateral' is what the reference's ``prefiltering=True`` runs
    (deblurring.py:108); 'rf' enables the commented-out domain-transform call (:107).
    ``faithful_cost``: also evaluate the reference's unconditional, unused
    ``fourier_gradients(img)`` (:61) so that CPU timings carry the same work.
    """
    dtype = np.dtype(dtype).type
    x = np.asarray(img)
    squeeze = x.ndim < 4
    if squeeze:
        x = to_tensor(x)[None]
    x = x.astype(dtype)
    if n_iter == 0:
        return to_array(x) if squeeze else x
    grad_img = None
    if remove_halo or faithful_cost:
        grad_img = fourier_gradients(x, dtype)
    cur = x
    for _ in range(n_iter):
        kernel = gaussian_blur_estimation(cur, c=c, b=b, q=q, ker_size=ker_size,
                                          discard_saturation=discard_saturation,
                                          dtype=dtype, trace=trace)
        if prefiltering:
            if prefilter == "rf":
                smooth = recursive_filter(cur, sigma_s=sigma_s, sigma_r=sigma_r,
                                          num_iterations=1, dtype=dtype)
            else:
                smooth = bilateral_filter(cur, dtype=dtype)
            noise = cur - smooth
            cur = inverse_filtering_rank3(smooth, kernel, alpha=alpha, b=beta,
                                          remove_halo=remove_halo, do_edgetaper=edgetaping,
                                          grad_img=grad_img, dtype=dtype)
            cur = cur + noise
        else:
            cur = inverse_filtering_rank3(cur, kernel, alpha=alpha, b=beta,
                                          remove_halo=remove_halo, do_edgetaper=edgetaping,
                                          grad_img=grad_img, dtype=dtype)
        cur = np.clip(cur, dtype(0), dtype(1)).astype(dtype)
    return to_array(cur) if squeeze else cur


# --------------------------------------------------------------------------------------
# filters.gaussian_filter (numpy kernel generator used for synthetic degradation)
# --------------------------------------------------------------------------------------
def gaussian_filter_np(sigma, theta, k_size=(25, 25)):
    """Generalised 2-D Gaussian with std (sigma[0], sigma[1]) and angle theta
    (filters.py:198-234), float32, normalised; dirac if the mass underflows."""
    l1, l2 = sigma
    th = -theta
    Q = np.array([[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]])
    S = Q @ np.diag([l1 ** 2, l2 ** 2]) @ Q.T
    inv = np.linalg.inv(S)
    kx, ky = int(k_size[0]), int(k_size[1])
    mu = np.array([kx // 2, ky // 2], dtype=np.float64)
    X, Y = np.meshgrid(np.arange(kx), np.arange(ky))
    zx = X - mu[0]
    zy = Y - mu[1]
    quad = inv[0, 0] * zx * zx + 2 * inv[0, 1] * zx * zy + inv[1, 1] * zy * zy
    raw = np.exp(-0.5 * quad).astype(np.float32)
    if raw.sum() < 1e-2:
        k = np.zeros_like(raw)
        k[kx // 2, ky // 2] = 1
        return k
    return raw / raw.sum()
