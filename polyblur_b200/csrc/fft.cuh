// Batched mixed-radix Stockham FFT in shared memory (any length, sm_100a).
//
// One CTA transforms `nb` independent complex sequences of length n that live in shared
// memory as buf[f * n + i].  Every stage reads one buffer and writes the other
// (autosort: no bit reversal), so two buffers of nb * n float2 are needed.  Radices
// 2,3,4,5,7,8 are unrolled in registers; any other prime factor goes through a generic
// O(R^2) butterfly, so every n works (n = 1104 = 8*3*2*23 is the padded torus of a 1080p
// image).  Twiddles come from a table tw[k] = exp(-2*pi*i*k/n), k < n, rounded to fp32
// from fp64 (k_twiddles), so no stage accumulates angle error.
//
// Used for (a) the spectral derivative of filters.fourier_gradients (filters.py:159-186),
// where two real rows/columns are packed into one complex transform, and (b) the
// blur-independent deconvolution engine.
#pragma once
#include "common.cuh"

namespace pb {

__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
// multiply by -i (forward-transform rotation):  -i (x + i y) = y - i x
__device__ __forceinline__ float2 mul_mi(float2 a) { return make_float2(a.y, -a.x); }

template <int R>
__device__ __forceinline__ void dft_reg(float2 (&v)[R]);

template <>
__device__ __forceinline__ void dft_reg<2>(float2 (&v)[2]) {
    float2 a = v[0], b = v[1];
    v[0] = cadd(a, b);
    v[1] = csub(a, b);
}

template <>
__device__ __forceinline__ void dft_reg<3>(float2 (&v)[3]) {
    const float S = 0.86602540378443864676f;  // sin(2 pi / 3)
    float2 t = cadd(v[1], v[2]);
    float2 d = csub(v[1], v[2]);
    float2 m = make_float2(v[0].x - 0.5f * t.x, v[0].y - 0.5f * t.y);
    float2 r = make_float2(S * d.y, -S * d.x);   // -i * S * d
    v[0] = cadd(v[0], t);
    v[1] = cadd(m, r);
    v[2] = csub(m, r);
}

template <>
__device__ __forceinline__ void dft_reg<4>(float2 (&v)[4]) {
    float2 t0 = cadd(v[0], v[2]);
    float2 t1 = csub(v[0], v[2]);
    float2 t2 = cadd(v[1], v[3]);
    float2 t3 = mul_mi(csub(v[1], v[3]));
    v[0] = cadd(t0, t2);
    v[1] = cadd(t1, t3);
    v[2] = csub(t0, t2);
    v[3] = csub(t1, t3);
}

template <>
__device__ __forceinline__ void dft_reg<5>(float2 (&v)[5]) {
    const float C1 = 0.30901699437494742410f;   // cos(2 pi / 5)
    const float C2 = -0.80901699437494742410f;  // cos(4 pi / 5)
    const float S1 = 0.95105651629515357212f;   // sin(2 pi / 5)
    const float S2 = 0.58778525229247312917f;   // sin(4 pi / 5)
    float2 a1 = cadd(v[1], v[4]), b1 = csub(v[1], v[4]);
    float2 a2 = cadd(v[2], v[3]), b2 = csub(v[2], v[3]);
    float2 e1 = make_float2(v[0].x + C1 * a1.x + C2 * a2.x, v[0].y + C1 * a1.y + C2 * a2.y);
    float2 e2 = make_float2(v[0].x + C2 * a1.x + C1 * a2.x, v[0].y + C2 * a1.y + C1 * a2.y);
    float2 s1 = make_float2(S1 * b1.x + S2 * b2.x, S1 * b1.y + S2 * b2.y);
    float2 s2 = make_float2(S2 * b1.x - S1 * b2.x, S2 * b1.y - S1 * b2.y);
    float2 r1 = mul_mi(s1), r2 = mul_mi(s2);
    v[0] = make_float2(v[0].x + a1.x + a2.x, v[0].y + a1.y + a2.y);
    v[1] = cadd(e1, r1);
    v[4] = csub(e1, r1);
    v[2] = cadd(e2, r2);
    v[3] = csub(e2, r2);
}

template <>
__device__ __forceinline__ void dft_reg<7>(float2 (&v)[7]) {
    const float C1 = 0.62348980185873353053f;   // cos(2 pi / 7)
    const float C2 = -0.22252093395631440429f;  // cos(4 pi / 7)
    const float C3 = -0.90096886790241912624f;  // cos(6 pi / 7)
    const float S1 = 0.78183148246802980871f;   // sin(2 pi / 7)
    const float S2 = 0.97492791218182360702f;   // sin(4 pi / 7)
    const float S3 = 0.43388373911755812048f;   // sin(6 pi / 7)
    float2 a1 = cadd(v[1], v[6]), b1 = csub(v[1], v[6]);
    float2 a2 = cadd(v[2], v[5]), b2 = csub(v[2], v[5]);
    float2 a3 = cadd(v[3], v[4]), b3 = csub(v[3], v[4]);
    // q = 1: cos(1,2,3) sin(1,2,3);  q = 2: cos(2,4->3,6->1) sin(2, 4 -> -3, 6 -> -1)
    // q = 3: cos(3, 6->1, 9->2) sin(3, 6 -> -1, 9 -> 2)
    float2 e1 = make_float2(v[0].x + C1 * a1.x + C2 * a2.x + C3 * a3.x, v[0].y + C1 * a1.y + C2 * a2.y + C3 * a3.y);
    float2 e2 = make_float2(v[0].x + C2 * a1.x + C3 * a2.x + C1 * a3.x, v[0].y + C2 * a1.y + C3 * a2.y + C1 * a3.y);
    float2 e3 = make_float2(v[0].x + C3 * a1.x + C1 * a2.x + C2 * a3.x, v[0].y + C3 * a1.y + C1 * a2.y + C2 * a3.y);
    float2 s1 = make_float2(S1 * b1.x + S2 * b2.x + S3 * b3.x, S1 * b1.y + S2 * b2.y + S3 * b3.y);
    float2 s2 = make_float2(S2 * b1.x - S3 * b2.x - S1 * b3.x, S2 * b1.y - S3 * b2.y - S1 * b3.y);
    float2 s3 = make_float2(S3 * b1.x - S1 * b2.x + S2 * b3.x, S3 * b1.y - S1 * b2.y + S2 * b3.y);
    float2 r1 = mul_mi(s1), r2 = mul_mi(s2), r3 = mul_mi(s3);
    v[0] = make_float2(v[0].x + a1.x + a2.x + a3.x, v[0].y + a1.y + a2.y + a3.y);
    v[1] = cadd(e1, r1);
    v[6] = csub(e1, r1);
    v[2] = cadd(e2, r2);
    v[5] = csub(e2, r2);
    v[3] = cadd(e3, r3);
    v[4] = csub(e3, r3);
}

template <>
__device__ __forceinline__ void dft_reg<8>(float2 (&v)[8]) {
    const float H = 0.70710678118654752440f;
    float2 e[4] = {v[0], v[2], v[4], v[6]};
    float2 o[4] = {v[1], v[3], v[5], v[7]};
    dft_reg<4>(e);
    dft_reg<4>(o);
    // w8^1 = H (1 - i), w8^2 = -i, w8^3 = H (-1 - i)
    float2 o1 = make_float2(H * (o[1].x + o[1].y), H * (o[1].y - o[1].x));
    float2 o2 = mul_mi(o[2]);
    float2 o3 = make_float2(H * (o[3].y - o[3].x), -H * (o[3].x + o[3].y));
    v[0] = cadd(e[0], o[0]);
    v[4] = csub(e[0], o[0]);
    v[1] = cadd(e[1], o1);
    v[5] = csub(e[1], o1);
    v[2] = cadd(e[2], o2);
    v[6] = csub(e[2], o2);
    v[3] = cadd(e[3], o3);
    v[7] = csub(e[3], o3);
}

// One Stockham stage of radix R over nb sequences.  Ns = product of earlier radices.
template <int R>
__device__ __forceinline__ void fft_stage(const float2* __restrict__ src, float2* __restrict__ dst,
                                          int n, int nb, int Ns, const float2* __restrict__ tw) {
    const int nbf = n / R;
    const int tstride = n / (Ns * R);
    const int total = nb * nbf;
    for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
        const int f = idx / nbf;
        const int j = idx - f * nbf;
        const int k = j % Ns;
        const float2* s = src + f * n + j;
        float2 v[R];
#pragma unroll
        for (int r = 0; r < R; ++r) v[r] = s[r * nbf];
        if (Ns > 1) {
            const int t = k * tstride;
#pragma unroll
            for (int r = 1; r < R; ++r) v[r] = cmul(v[r], __ldg(&tw[r * t]));
        }
        dft_reg<R>(v);
        float2* d = dst + f * n + (j - k) * R + k;
#pragma unroll
        for (int q = 0; q < R; ++q) d[q * Ns] = v[q];
    }
}

// Generic radix (any R >= 2): one thread per output point, O(R) shared-memory reads each.
__device__ __forceinline__ void fft_stage_generic(const float2* __restrict__ src, float2* __restrict__ dst,
                                                  int n, int nb, int Ns, int R,
                                                  const float2* __restrict__ tw) {
    const int nbf = n / R;
    const int tstride = n / (Ns * R);
    const int rstride = n / R;
    const int total = nb * n;
    for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
        const int f = idx / n;
        const int rem = idx - f * n;
        const int q = rem / nbf;
        const int j = rem - q * nbf;
        const int k = j % Ns;
        const float2* s = src + f * n + j;
        float2 acc = s[0];
        int t1 = 0;   // r * k * tstride            (< n)
        int t2 = 0;   // ((q * r) mod R) * (n / R)   (< n)
        for (int r = 1; r < R; ++r) {
            t1 += k * tstride;
            t2 += q * rstride;
            if (t2 >= n) t2 -= n;
            int t = t1 + t2;
            if (t >= n) t -= n;
            acc = cadd(acc, cmul(s[r * nbf], __ldg(&tw[t])));
        }
        dst[f * n + (j - k) * R + k + q * Ns] = acc;
    }
}

// Forward DFT of nb sequences.  On return `a` points at the buffer holding the result
// (natural order) and `b` at the other one.  All threads of the CTA must call this; the
// data written by the caller must be visible (the function starts with no barrier, the
// caller syncs), and a barrier has been executed after the last stage.
__device__ __forceinline__ void fft_forward(float2*& a, float2*& b, int nb, const FftPlan& plan,
                                            const float2* __restrict__ tw) {
    const int n = plan.n;
    int Ns = 1;
    for (int s = 0; s < plan.ns; ++s) {
        const int R = plan.radix[s];
        switch (R) {
            case 2: fft_stage<2>(a, b, n, nb, Ns, tw); break;
            case 3: fft_stage<3>(a, b, n, nb, Ns, tw); break;
            case 4: fft_stage<4>(a, b, n, nb, Ns, tw); break;
            case 5: fft_stage<5>(a, b, n, nb, Ns, tw); break;
            case 7: fft_stage<7>(a, b, n, nb, Ns, tw); break;
            case 8: fft_stage<8>(a, b, n, nb, Ns, tw); break;
            default: fft_stage_generic(a, b, n, nb, Ns, R, tw); break;
        }
        __syncthreads();
        float2* t = a; a = b; b = t;
        Ns *= R;
    }
}

// Angular frequency of DFT bin k as the reference builds it (filters.py:175-181):
// f = (k' - n//2) / n in float32 for the shifted index, times float32(2*pi).  The Nyquist
// bin of an even length carries no real signal and is zeroed because two
// real sequences share one complex transform here.
__device__ __forceinline__ float bin_omega(int k, int n) {
    int kk = (k < (n + 1) / 2) ? k : k - n;
    if ((n & 1) == 0 && k == n / 2) return 0.0f;
    float f = __fdiv_rn((float)kk, (float)n);
    return __fmul_rn(6.283185307179586f, f);
}

// Spectral derivative of nb packed sequences (real part = one signal, imaginary = another):
// z' = IDFT(i * omega * DFT(z)).  Uses IDFT(Y) = conj(DFT(conj(Y))) / n; the result left in
// `a` is DFT(conj(Y)), i.e. the caller reads  d(real signal) = a.x / n,  d(imag signal) = -a.y / n.
__device__ __forceinline__ void fft_derivative(float2*& a, float2*& b, int nb, const FftPlan& plan,
                                               const float2* __restrict__ tw) {
    const int n = plan.n;
    fft_forward(a, b, nb, plan, tw);
    for (int idx = threadIdx.x; idx < nb * n; idx += blockDim.x) {
        const int k = idx % n;
        const float w = bin_omega(k, n);
        float2 z = a[idx];
        // i*w*z = w * (-z.y + i z.x);  conj -> (-w z.y, -w z.x)
        a[idx] = make_float2(-w * z.y, -w * z.x);
    }
    __syncthreads();
    fft_forward(a, b, nb, plan, tw);
}

// tw[k] = exp(-2 pi i k / n) for k < n, computed in fp64 and rounded once.
__global__ void k_twiddles(float2* __restrict__ tw, int n);

}  // namespace pb
