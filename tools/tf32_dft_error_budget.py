#!/usr/bin/env python
"""Error budget of a DFT-as-GEMM on TF32 tensor cores, emulated on the CPU (numpy), BEFORE any kernel is written.

The FFT engine is bound by the float32 pipe; the one way past it would be to run the radix-R stages as
small GEMMs (2R x 2R real DFT matrix times the data) on tcgen05 with kind::tf32.  The question here is only numeric:
what does a 2016-point row transform -> multiply by a deconvolution filter -> inverse transform lose against float64

  * as a float32 FFT (pocketfft on complex64: what the CUDA-core engine's arithmetic amounts to),
  * with the stage matrices and the data rounded to TF32 (10-bit mantissa), products exact, float32 accumulation,
  * with the 3 x TF32 split (hi hi + hi lo + lo hi, the usual float32 emulation on TF32 tensor cores),

on image-like rows in [0, 1], against the end-to-end tolerance of 1e-5 (three iterations share it)?
Stages 16 x 14 x 9 like PlanX2016, twiddles between the stages applied in float32 (CUDA cores).
"""
import json

import numpy as np
import scipy.fft


def tf32(x):
    """round-to-nearest-even of float32 to 10 mantissa bits"""
    u = np.asarray(x, np.float32).view(np.uint32).astype(np.uint64)
    u = (u + 0x0FFF + ((u >> 13) & 1)) & 0xFFFFE000
    return u.astype(np.uint32).view(np.float32)


def gemm(A, B, mode):
    """real A (m, k) @ B (k, n) with float32 accumulation; mode 'f32' | 'tf32' | '3xtf32'"""
    A = A.astype(np.float32)
    B = B.astype(np.float32)
    if mode == "f32":
        return A @ B
    Ah, Bh = tf32(A), tf32(B)
    if mode == "tf32":
        return Ah @ Bh
    Al, Bl = tf32(A - Ah), tf32(B - Bh)
    return (Al @ Bh + Ah @ Bl) + Ah @ Bh


def dft_stage(x, R, sign, mode):
    """x: (..., R, M) complex64 -> DFT over the R axis as ONE real GEMM with the (2R x 2R) matrix [[C, -S], [S, C]]"""
    k = np.arange(R)
    ang = sign * 2 * np.pi * np.outer(k, k) / R
    C, S = np.cos(ang), np.sin(ang)
    W = np.block([[C, -S], [S, C]])                                   # (2R, 2R) real, float64 -> rounded inside gemm
    shp = x.shape
    xr = np.concatenate([x.real, x.imag], axis=-2)                    # (..., 2R, M)
    flat = xr.reshape(-1, 2 * R, shp[-1])
    out = np.stack([gemm(W, f, mode) for f in flat]).reshape(shp[:-2] + (2 * R, shp[-1]))
    return (out[..., :R, :] + 1j * out[..., R:, :]).astype(np.complex64)


def fft_by_gemm(x, radices, sign, mode):
    """decimation-in-frequency Cooley-Tukey, every butterfly layer one GEMM, twiddles in float32; returns natural order"""
    n = x.shape[-1]
    if len(radices) == 0:
        return x
    R = radices[0]
    M = n // R
    y = dft_stage(x.reshape(x.shape[:-1] + (R, M)), R, sign, mode)   # y[q, j] = sum_m x[j + m M] w_R^{q m}
    if M == 1:
        return y.reshape(x.shape)
    q = np.arange(R)[:, None]
    j = np.arange(M)[None, :]
    tw = np.exp(sign * 2j * np.pi * q * j / n).astype(np.complex64)
    y = (y * tw).astype(np.complex64)
    sub = fft_by_gemm(y, radices[1:], sign, mode)                     # (..., R, M): X[q + R k'] for each q
    return np.swapaxes(sub, -1, -2).reshape(x.shape)                  # index k = q + R k'


def main():
    rng = np.random.default_rng(0)
    n, rows = 2016, 64
    # image-like rows: blurred mosaic + a little noise, two real rows packed into one complex sequence
    blocks = np.repeat(rng.random((2 * rows, n // 12 + 1)), 12, axis=1)[:, :n]
    kern = np.exp(-0.5 * (np.arange(-12, 13) / 2.5) ** 2)
    kern /= kern.sum()
    img = np.stack([np.convolve(r, kern, mode="same") for r in blocks]) + 0.01 * rng.random((2 * rows, n))
    img = np.clip(img, 0, 1)
    z = (img[0::2] + 1j * img[1::2])
    # deconvolution filter of a sigma = 2.5 Gaussian, alpha = 6, beta = 1 (deblurring.py:160-167)
    K = np.fft.fft(np.roll(np.pad(kern, (0, n - 25)), -12)).real
    a3, a2, a1, b = 6 / 2 - 1 + 2, 3 * 1 - 6 - 6, 5 - 3 * 1 + 6 / 2, 1
    Hf = ((a3 * K + a2) * K + a1) * K + b
    truth = np.fft.ifft(np.fft.fft(z) * Hf)
    res = {"n": n, "rows": 2 * rows, "filter_max": float(np.abs(Hf).max())}
    y32 = scipy.fft.ifft(scipy.fft.fft(z.astype(np.complex64)) * Hf.astype(np.float32)).astype(np.complex64)
    res["float32 FFT (pocketfft complex64)"] = float(np.abs(y32 - truth).max())
    for mode in ("f32", "tf32", "3xtf32"):
        X = fft_by_gemm(z.astype(np.complex64), [16, 14, 9], -1, mode)
        fwd_err = float(np.abs(X - np.fft.fft(z)).max() / np.abs(np.fft.fft(z)).max())
        Y = (X * Hf.astype(np.float32)).astype(np.complex64)
        y = fft_by_gemm(Y, [16, 14, 9], +1, mode) / np.float32(n)
        res[f"GEMM stages, {mode}"] = {"deconvolved row max-abs error": float(np.abs(y - truth).max()),
                                       "forward spectrum error / max |X|": fwd_err}
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
