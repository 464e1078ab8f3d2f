// In-place large-radix FFT in shared memory (sm_100a), second-generation core.
//
// A CTA transforms `nb` complex sequences of length n stored as x[f * stride + i] (stride >= n;
// an odd stride keeps the transposing loads / stores of the column kernels off one bank).  The forward
// transform is decimation-in-frequency (natural order in, digit-scrambled order out), the
// inverse-direction transform is its transpose, decimation-in-time (scrambled in, natural out).
// A convolution / spectral derivative never needs the spectrum in natural order, so no
// reordering pass exists: a butterfly reads and writes the same R slots (truly in place, one
// buffer, one barrier per stage).  Radices up to 16 are done in registers -- 6, 9, 10, 12, 14,
// 15 and 16 as two nested small DFTs with compile-time inner twiddles -- so the sizes Polyblur
// meets need three or four stages (1920 = 16*12*10, 1080 = 12*10*9, 1152 = 16*9*8,
// 3840 = 16*16*15, 2160 = 16*15*9).
//
// Everything here is __host__ __device__ with the thread index passed explicitly, so that the
// arithmetic can be unit-tested on the CPU by running the "threads" one after the other
// (tests/test_fft2_host.py builds csrc/fft2_host_test.cu).
//
// Used by (a) the spectral derivative of filters.fourier_gradients (polyblur/filters.py:159-186)
// inside the blur estimator and (b) the blur-independent FFT deconvolution engine that replaces
// compute_polynomial_fft (polyblur/deblurring.py:141-169).
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <mutex>
#include <unordered_map>
#include <utility>

#if defined(__CUDACC__)
#define PB_HD __host__ __device__ __forceinline__
#define PB_HDC __host__ __device__
#else
#define PB_HD inline
#define PB_HDC
#endif
#if defined(__CUDA_ARCH__)
#define PB_LDG(p) __ldg(p)
#else
#define PB_LDG(p) (*(p))
#endif

#include "fft_twc.cuh"

#define PB_FFT2_MAX_STAGES 8

namespace pb {

struct Fft2Plan {
    int n;
    int ns;
    int radix[PB_FFT2_MAX_STAGES];   // DIF order: radix[0] is applied at sub-length n
    // Stage twiddles live in one table, q-major per stage so that the lanes of a warp (consecutive
    // j) read consecutive entries:  stw[tw_off[s] + (q - 1) M_s + j] = exp(-2 pi i j q / L_s),
    // q = 1..R_s-1, j < M_s = L_s / R_s  (no entries for a stage with M_s = 1).
    int tw_off[PB_FFT2_MAX_STAGES];
    int tw_total;
};

// ---- small complex helpers ---------------------------------------------------------------
PB_HD float2 c_add(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
PB_HD float2 c_sub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
PB_HD float2 c_mul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
PB_HD float2 c_mul_mi(float2 a) { return make_float2(a.y, -a.x); }     // * (-i)
PB_HD float2 c_mul_pi(float2 a) { return make_float2(-a.y, a.x); }     // * (+i)

// ---- register butterflies: v <- DFT_R(v), forward sign -------------------------------------
template <int R>
struct Dft;

template <>
struct Dft<2> {
    static PB_HD void run(float2 (&v)[2]) {
        const float2 a = v[0], b = v[1];
        v[0] = c_add(a, b);
        v[1] = c_sub(a, b);
    }
};

template <>
struct Dft<3> {
    static PB_HD void run(float2 (&v)[3]) {
        const float S = 0.86602540378443864676f;
        const float2 t = c_add(v[1], v[2]);
        const float2 d = c_sub(v[1], v[2]);
        const float2 m = make_float2(v[0].x - 0.5f * t.x, v[0].y - 0.5f * t.y);
        const float2 r = make_float2(S * d.y, -S * d.x);
        v[0] = c_add(v[0], t);
        v[1] = c_add(m, r);
        v[2] = c_sub(m, r);
    }
};

template <>
struct Dft<4> {
    static PB_HD void run(float2 (&v)[4]) {
        const float2 t0 = c_add(v[0], v[2]);
        const float2 t1 = c_sub(v[0], v[2]);
        const float2 t2 = c_add(v[1], v[3]);
        const float2 t3 = c_mul_mi(c_sub(v[1], v[3]));
        v[0] = c_add(t0, t2);
        v[1] = c_add(t1, t3);
        v[2] = c_sub(t0, t2);
        v[3] = c_sub(t1, t3);
    }
};

template <>
struct Dft<5> {
    static PB_HD void run(float2 (&v)[5]) {
        const float C1 = 0.30901699437494742410f, C2 = -0.80901699437494742410f;
        const float S1 = 0.95105651629515357212f, S2 = 0.58778525229247312917f;
        const float2 a1 = c_add(v[1], v[4]), b1 = c_sub(v[1], v[4]);
        const float2 a2 = c_add(v[2], v[3]), b2 = c_sub(v[2], v[3]);
        const float2 e1 = make_float2(v[0].x + C1 * a1.x + C2 * a2.x, v[0].y + C1 * a1.y + C2 * a2.y);
        const float2 e2 = make_float2(v[0].x + C2 * a1.x + C1 * a2.x, v[0].y + C2 * a1.y + C1 * a2.y);
        const float2 r1 = c_mul_mi(make_float2(S1 * b1.x + S2 * b2.x, S1 * b1.y + S2 * b2.y));
        const float2 r2 = c_mul_mi(make_float2(S2 * b1.x - S1 * b2.x, S2 * b1.y - S1 * b2.y));
        v[0] = make_float2(v[0].x + a1.x + a2.x, v[0].y + a1.y + a2.y);
        v[1] = c_add(e1, r1);
        v[4] = c_sub(e1, r1);
        v[2] = c_add(e2, r2);
        v[3] = c_sub(e2, r2);
    }
};

template <>
struct Dft<7> {
    static PB_HD void run(float2 (&v)[7]) {
        const float C1 = 0.62348980185873353053f, C2 = -0.22252093395631440429f, C3 = -0.90096886790241912624f;
        const float S1 = 0.78183148246802980871f, S2 = 0.97492791218182360702f, S3 = 0.43388373911755812048f;
        const float2 a1 = c_add(v[1], v[6]), b1 = c_sub(v[1], v[6]);
        const float2 a2 = c_add(v[2], v[5]), b2 = c_sub(v[2], v[5]);
        const float2 a3 = c_add(v[3], v[4]), b3 = c_sub(v[3], v[4]);
        const float2 e1 = make_float2(v[0].x + C1 * a1.x + C2 * a2.x + C3 * a3.x, v[0].y + C1 * a1.y + C2 * a2.y + C3 * a3.y);
        const float2 e2 = make_float2(v[0].x + C2 * a1.x + C3 * a2.x + C1 * a3.x, v[0].y + C2 * a1.y + C3 * a2.y + C1 * a3.y);
        const float2 e3 = make_float2(v[0].x + C3 * a1.x + C1 * a2.x + C2 * a3.x, v[0].y + C3 * a1.y + C1 * a2.y + C2 * a3.y);
        const float2 r1 = c_mul_mi(make_float2(S1 * b1.x + S2 * b2.x + S3 * b3.x, S1 * b1.y + S2 * b2.y + S3 * b3.y));
        const float2 r2 = c_mul_mi(make_float2(S2 * b1.x - S3 * b2.x - S1 * b3.x, S2 * b1.y - S3 * b2.y - S1 * b3.y));
        const float2 r3 = c_mul_mi(make_float2(S3 * b1.x - S1 * b2.x + S2 * b3.x, S3 * b1.y - S1 * b2.y + S2 * b3.y));
        v[0] = make_float2(v[0].x + a1.x + a2.x + a3.x, v[0].y + a1.y + a2.y + a3.y);
        v[1] = c_add(e1, r1);
        v[6] = c_sub(e1, r1);
        v[2] = c_add(e2, r2);
        v[5] = c_sub(e2, r2);
        v[3] = c_add(e3, r3);
        v[4] = c_sub(e3, r3);
    }
};

template <>
struct Dft<8> {
    static PB_HD void run(float2 (&v)[8]) {
        const float H = 0.70710678118654752440f;
        float2 e[4] = {v[0], v[2], v[4], v[6]};
        float2 o[4] = {v[1], v[3], v[5], v[7]};
        Dft<4>::run(e);
        Dft<4>::run(o);
        const float2 o1 = make_float2(H * (o[1].x + o[1].y), H * (o[1].y - o[1].x));
        const float2 o2 = c_mul_mi(o[2]);
        const float2 o3 = make_float2(H * (o[3].y - o[3].x), -H * (o[3].x + o[3].y));
        v[0] = c_add(e[0], o[0]);
        v[4] = c_sub(e[0], o[0]);
        v[1] = c_add(e[1], o1);
        v[5] = c_sub(e[1], o1);
        v[2] = c_add(e[2], o2);
        v[6] = c_sub(e[2], o2);
        v[3] = c_add(e[3], o3);
        v[7] = c_sub(e[3], o3);
    }
};

// multiply by the compile-time constant W_R^k (k is a literal once the loops are unrolled)
template <int R>
PB_HD float2 mul_twc(float2 a, int k) {
    k %= R;
    if (k == 0) return a;
    if (4 * k == R) return c_mul_mi(a);
    if (2 * k == R) return make_float2(-a.x, -a.y);
    if (4 * k == 3 * R) return c_mul_pi(a);
    return c_mul(a, twc<R>(k));
}

// DFT of size A*B as A transforms of size B, constant twiddles, B transforms of size A:
// input index m = n1 + A n2, output index k = B k1 + k2.
template <int A, int B>
struct DftAB {
    static PB_HD void run(float2 (&v)[A * B]) {
        constexpr int R = A * B;
        float2 y[A][B];
#pragma unroll
        for (int n1 = 0; n1 < A; ++n1) {
            float2 t[B];
#pragma unroll
            for (int n2 = 0; n2 < B; ++n2) t[n2] = v[n1 + A * n2];
            Dft<B>::run(t);
#pragma unroll
            for (int k2 = 0; k2 < B; ++k2) y[n1][k2] = mul_twc<R>(t[k2], n1 * k2);
        }
#pragma unroll
        for (int k2 = 0; k2 < B; ++k2) {
            float2 t[A];
#pragma unroll
            for (int n1 = 0; n1 < A; ++n1) t[n1] = y[n1][k2];
            Dft<A>::run(t);
#pragma unroll
            for (int k1 = 0; k1 < A; ++k1) v[B * k1 + k2] = t[k1];
        }
    }
};

template <> struct Dft<6> { static PB_HD void run(float2 (&v)[6]) { DftAB<2, 3>::run(v); } };
template <> struct Dft<9> { static PB_HD void run(float2 (&v)[9]) { DftAB<3, 3>::run(v); } };
template <> struct Dft<10> { static PB_HD void run(float2 (&v)[10]) { DftAB<2, 5>::run(v); } };
template <> struct Dft<12> { static PB_HD void run(float2 (&v)[12]) { DftAB<3, 4>::run(v); } };
template <> struct Dft<14> { static PB_HD void run(float2 (&v)[14]) { DftAB<2, 7>::run(v); } };
template <> struct Dft<15> { static PB_HD void run(float2 (&v)[15]) { DftAB<3, 5>::run(v); } };
template <> struct Dft<16> { static PB_HD void run(float2 (&v)[16]) { DftAB<4, 4>::run(v); } };
// large radices of the two-stage column plans (deconv_fft.cu k_fft_cols2): 64-72 data registers per thread
template <> struct Dft<18> { static PB_HD void run(float2 (&v)[18]) { DftAB<2, 9>::run(v); } };
template <> struct Dft<20> { static PB_HD void run(float2 (&v)[20]) { DftAB<4, 5>::run(v); } };
template <> struct Dft<24> { static PB_HD void run(float2 (&v)[24]) { DftAB<3, 8>::run(v); } };
template <> struct Dft<30> { static PB_HD void run(float2 (&v)[30]) { DftAB<5, 6>::run(v); } };
template <> struct Dft<32> { static PB_HD void run(float2 (&v)[32]) { DftAB<4, 8>::run(v); } };
template <> struct Dft<36> { static PB_HD void run(float2 (&v)[36]) { DftAB<4, 9>::run(v); } };

// Odd primes 11 and 13: direct O(R^2) sum with constant twiddles (rare sizes only).
template <int R>
struct DftDirect {
    static PB_HD void run(float2 (&v)[R]) {
        float2 o[R];
#pragma unroll
        for (int k = 0; k < R; ++k) {
            float2 acc = v[0];
#pragma unroll
            for (int m = 1; m < R; ++m) acc = c_add(acc, mul_twc<R>(v[m], m * k));
            o[k] = acc;
        }
#pragma unroll
        for (int k = 0; k < R; ++k) v[k] = o[k];
    }
};
template <> struct Dft<11> { static PB_HD void run(float2 (&v)[11]) { DftDirect<11>::run(v); } };
template <> struct Dft<13> { static PB_HD void run(float2 (&v)[13]) { DftDirect<13>::run(v); } };

// exact floor(a / d) for 0 <= a < 2^21, d >= 1, given inv = 1.0f / d
PB_HD int fast_div(int a, int d, float inv) {
    int q = (int)(((float)a + 0.5f) * inv);
    (void)d;
    return q;
}

// ---- one DIF stage: sub-length L, radix R, in place -----------------------------------------
//   v[m] = x[b + j + m M];  V = DFT_R(v);  x[b + j + q M] = V[q] W_L^{j q},   M = L / R
template <int R>
PB_HD void fft2_dif_stage(float2* x, int n, int stride, int nb, int L, const float2* __restrict__ stw, int tid,
                           int nthr) {
    const int M = L / R;
    const int bps = n / R;                 // butterflies per sequence
    const float inv_bps = 1.0f / (float)bps, inv_M = 1.0f / (float)M;
    const int total = nb * bps;
    for (int idx = tid; idx < total; idx += nthr) {
        const int f = fast_div(idx, bps, inv_bps);
        const int rem = idx - f * bps;
        const int blk = fast_div(rem, M, inv_M);
        const int j = rem - blk * M;
        float2* p = x + f * stride + blk * L + j;
        float2 v[R];
#pragma unroll
        for (int m = 0; m < R; ++m) v[m] = p[m * M];
        Dft<R>::run(v);
        p[0] = v[0];
        if (M == 1) {
#pragma unroll
            for (int q = 1; q < R; ++q) p[q] = v[q];
        } else {
#pragma unroll
            for (int q = 1; q < R; ++q) p[q * M] = c_mul(v[q], PB_LDG(stw + (q - 1) * M + j));
        }
    }
}

// ---- one DIT stage (the transpose of the DIF stage) -------------------------------------------
//   v[q] = x[b + j + q M] W_L^{j q};  V = DFT_R(v);  x[b + j + m M] = V[m]
template <int R>
PB_HD void fft2_dit_stage(float2* x, int n, int stride, int nb, int L, const float2* __restrict__ stw, int tid,
                           int nthr, const float* premul = nullptr, int premode = 1) {
    const int M = L / R;
    const int bps = n / R;
    const float inv_bps = 1.0f / (float)bps, inv_M = 1.0f / (float)M;
    const int total = nb * bps;
    for (int idx = tid; idx < total; idx += nthr) {
        const int f = fast_div(idx, bps, inv_bps);
        const int rem = idx - f * bps;
        const int blk = fast_div(rem, M, inv_M);
        const int j = rem - blk * M;
        float2* p = x + f * stride + blk * L + j;
        float2 v[R];
        v[0] = p[0];
        if (M == 1) {
#pragma unroll
            for (int q = 1; q < R; ++q) v[q] = p[q];
            if (premul) {
                // multiplier folded into the first inverse-direction stage, together with the
                // re/im swap of the inverse-by-forward trick:
                //   mode 1 (spectral derivative, one table for all sequences): y = i w z -> (w z.x, -w z.y)
                //   mode 2 (real transfer function, one table per sequence):   y = h z   -> (h z.y,  h z.x)
                const float* pm = premul + blk * L + (premode == 2 ? f * n : 0);
#pragma unroll
                for (int q = 0; q < R; ++q) {
                    const float w = premode == 2 ? pm[q] : PB_LDG(pm + q);
                    v[q] = premode == 2 ? make_float2(w * v[q].y, w * v[q].x) : make_float2(w * v[q].x, -w * v[q].y);
                }
            }
        } else {
#pragma unroll
            for (int q = 1; q < R; ++q) v[q] = c_mul(p[q * M], PB_LDG(stw + (q - 1) * M + j));
        }
        Dft<R>::run(v);
#pragma unroll
        for (int m = 0; m < R; ++m) p[m * M] = v[m];
    }
}

// Barrier between stages: the whole CTA cooperates on the sequences (WARP = false), or one warp
// owns them and only needs a warp-level barrier (WARP = true: warps of a CTA run decoupled).
template <bool WARP>
PB_HD void fft2_sync() {
#if defined(__CUDA_ARCH__)
    if (WARP) __syncwarp(); else __syncthreads();
#endif
}

#define PB_FFT2_DISPATCH(FN, R_, ...)                         \
    switch (R_) {                                             \
        case 2: FN<2>(__VA_ARGS__); break;                    \
        case 3: FN<3>(__VA_ARGS__); break;                    \
        case 4: FN<4>(__VA_ARGS__); break;                    \
        case 5: FN<5>(__VA_ARGS__); break;                    \
        case 6: FN<6>(__VA_ARGS__); break;                    \
        case 7: FN<7>(__VA_ARGS__); break;                    \
        case 8: FN<8>(__VA_ARGS__); break;                    \
        case 9: FN<9>(__VA_ARGS__); break;                    \
        case 10: FN<10>(__VA_ARGS__); break;                  \
        case 11: FN<11>(__VA_ARGS__); break;                  \
        case 12: FN<12>(__VA_ARGS__); break;                  \
        case 13: FN<13>(__VA_ARGS__); break;                  \
        case 14: FN<14>(__VA_ARGS__); break;                  \
        case 15: FN<15>(__VA_ARGS__); break;                  \
        default: FN<16>(__VA_ARGS__); break;                  \
    }

// Forward DFT, natural order in -> scrambled order out.  The caller has made its writes to x
// visible (barrier) before the call; a barrier has been executed after the last stage.
// On the host (unit tests) the caller loops tid over [0, nthr) per stage via fft2_*_stage.
template <bool WARP = false>
PB_HD void fft2_forward_dif(float2* x, int stride, int nb, const Fft2Plan& plan, const float2* __restrict__ tw,
                            int tid, int nthr) {
    int L = plan.n;
    for (int s = 0; s < plan.ns; ++s) {
        const int R = plan.radix[s];
        PB_FFT2_DISPATCH(fft2_dif_stage, R, x, plan.n, stride, nb, L, tw + plan.tw_off[s], tid, nthr);
        fft2_sync<WARP>();
        L /= R;
    }
}

// Forward-sign DFT, scrambled order in -> natural order out (inverse transform through the
// swap trick: IDFT(y) = swap(DFT(swap(y))) / n, the swaps are folded into the caller's code).
template <bool WARP = false>
PB_HD void fft2_forward_dit(float2* x, int stride, int nb, const Fft2Plan& plan, const float2* __restrict__ tw,
                            int tid, int nthr, const float* premul = nullptr, int premode = 1) {
    int L = 1;
    for (int s = plan.ns - 1; s >= 0; --s) {
        const int R = plan.radix[s];
        L *= R;
        PB_FFT2_DISPATCH(fft2_dit_stage, R, x, plan.n, stride, nb, L, tw + plan.tw_off[s], tid, nthr,
                         (s == plan.ns - 1) ? premul : nullptr, premode);
        fft2_sync<WARP>();
    }
}

// ---- fused middle: last DIF stage + pointwise multiplier + first DIT stage -------------------------
// Both stages have M = 1: they work on the same R contiguous slots and carry no twiddles, so when a
// forward transform is followed by a pointwise multiplication and the inverse-direction transform
// (spectral derivative, transfer function), one thread does  DFT_R -> multiply (+ re/im swap) -> DFT_R
// in registers: one shared-memory round trip and one barrier less per transform pair.
//   premode 1: y = i w z        -> (w z.x, -w z.y)   (one table for all sequences)
//   premode 2: y = h z, swapped -> (h z.y,  h z.x)   (one table per sequence)
template <int R>
PB_HD void fft2_mid_stage(float2* x, int n, int stride, int nb, int tid, int nthr, const float* premul, int premode) {
    const int bps = n / R;
    const float inv_bps = 1.0f / (float)bps;
    const int total = nb * bps;
    for (int idx = tid; idx < total; idx += nthr) {
        const int f = fast_div(idx, bps, inv_bps);
        const int blk = idx - f * bps;
        float2* p = x + f * stride + blk * R;
        float2 v[R];
#pragma unroll
        for (int m = 0; m < R; ++m) v[m] = p[m];
        Dft<R>::run(v);
        const float* pm = premul + blk * R + (premode == 2 ? f * n : 0);
#pragma unroll
        for (int q = 0; q < R; ++q) {
            const float w = premode == 2 ? pm[q] : PB_LDG(pm + q);
            v[q] = premode == 2 ? make_float2(w * v[q].y, w * v[q].x) : make_float2(w * v[q].x, -w * v[q].y);
        }
        Dft<R>::run(v);
#pragma unroll
        for (int m = 0; m < R; ++m) p[m] = v[m];
    }
}

// Forward DIF, multiply by `premul`, inverse-direction DIT, with the two innermost stages fused.
// Same result as fft2_forward_dif followed by fft2_forward_dit(premul); needs plan.ns >= 1.
template <bool WARP = false>
PB_HD void fft2_forward_mul_inverse(float2* x, int stride, int nb, const Fft2Plan& plan, const float2* __restrict__ tw,
                                    int tid, int nthr, const float* premul, int premode) {
    int L = plan.n;
    for (int s = 0; s < plan.ns - 1; ++s) {
        const int R = plan.radix[s];
        PB_FFT2_DISPATCH(fft2_dif_stage, R, x, plan.n, stride, nb, L, tw + plan.tw_off[s], tid, nthr);
        fft2_sync<WARP>();
        L /= R;
    }
    PB_FFT2_DISPATCH(fft2_mid_stage, plan.radix[plan.ns - 1], x, plan.n, stride, nb, tid, nthr, premul, premode);
    fft2_sync<WARP>();
    L = plan.radix[plan.ns - 1];
    for (int s = plan.ns - 2; s >= 0; --s) {
        const int R = plan.radix[s];
        L *= R;
        PB_FFT2_DISPATCH(fft2_dit_stage, R, x, plan.n, stride, nb, L, tw + plan.tw_off[s], tid, nthr, nullptr, 0);
        fft2_sync<WARP>();
    }
}

// ---- fused first / last stages ---------------------------------------------------------------
// The first DIF stage (sub-length n) reads every sample exactly once and the last DIT stage
// (sub-length n) writes every sample exactly once, in both cases at x[j + m M] with j running
// over consecutive threads.  Handing those two stages a source / sink functor lets a kernel feed
// the transform straight from global memory and write its result straight back, without a
// separate load or store pass through shared memory.  (Measured on B200 for the estimator's row
// kernel: 2.18 ms against 1.31 ms for the staged version with 128-bit loads -- the scalar accesses
// and the extra registers cost more than the two passes saved -- so the shipped kernels stage through
// shared memory; the fused stages stay available and unit-tested.)
//     float2 Src::operator()(int f, int i)            sample i of sequence f
//     void   Dst::operator()(int f, int i, float2 v)  result i of sequence f
template <int R, class Src>
PB_HD void fft2_dif_first(float2* x, int n, int stride, int nb, const float2* __restrict__ stw, int tid, int nthr,
                          Src& src) {
    const int M = n / R;                   // L = n: one block per sequence, j = butterfly index
    const float inv_M = 1.0f / (float)M;
    const int total = nb * M;
    for (int idx = tid; idx < total; idx += nthr) {
        const int f = fast_div(idx, M, inv_M);
        const int j = idx - f * M;
        float2* p = x + f * stride + j;
        float2 v[R];
#pragma unroll
        for (int m = 0; m < R; ++m) v[m] = src(f, j + m * M);
        Dft<R>::run(v);
        p[0] = v[0];
        if (M == 1) {
#pragma unroll
            for (int q = 1; q < R; ++q) p[q] = v[q];
        } else {
#pragma unroll
            for (int q = 1; q < R; ++q) p[q * M] = c_mul(v[q], PB_LDG(stw + (q - 1) * M + j));
        }
    }
}

template <int R, class Dst>
PB_HD void fft2_dit_last(const float2* x, int n, int stride, int nb, const float2* __restrict__ stw, int tid,
                         int nthr, Dst& dst) {
    const int M = n / R;
    const float inv_M = 1.0f / (float)M;
    const int total = nb * M;
    for (int idx = tid; idx < total; idx += nthr) {
        const int f = fast_div(idx, M, inv_M);
        const int j = idx - f * M;
        const float2* p = x + f * stride + j;
        float2 v[R];
        v[0] = p[0];
#pragma unroll
        for (int q = 1; q < R; ++q) v[q] = c_mul(p[q * M], PB_LDG(stw + (q - 1) * M + j));
        Dft<R>::run(v);
#pragma unroll
        for (int m = 0; m < R; ++m) dst(f, j + m * M, v[m]);
    }
}

#define PB_FFT2_DISPATCH2(FN, R_, T_, ...)                    \
    switch (R_) {                                             \
        case 2: FN<2, T_>(__VA_ARGS__); break;                \
        case 3: FN<3, T_>(__VA_ARGS__); break;                \
        case 4: FN<4, T_>(__VA_ARGS__); break;                \
        case 5: FN<5, T_>(__VA_ARGS__); break;                \
        case 6: FN<6, T_>(__VA_ARGS__); break;                \
        case 7: FN<7, T_>(__VA_ARGS__); break;                \
        case 8: FN<8, T_>(__VA_ARGS__); break;                \
        case 9: FN<9, T_>(__VA_ARGS__); break;                \
        case 10: FN<10, T_>(__VA_ARGS__); break;              \
        case 11: FN<11, T_>(__VA_ARGS__); break;              \
        case 12: FN<12, T_>(__VA_ARGS__); break;              \
        case 13: FN<13, T_>(__VA_ARGS__); break;              \
        case 14: FN<14, T_>(__VA_ARGS__); break;              \
        case 15: FN<15, T_>(__VA_ARGS__); break;              \
        default: FN<16, T_>(__VA_ARGS__); break;              \
    }

// Forward DIF transform fed by `src` (needs plan.ns >= 2: the fused stage must have M > 1 so that
// the multiplier-folding first inverse stage stays a separate one).  No barrier is needed before
// the call; a barrier has been executed after the last stage.
template <class Src, bool WARP = false>
PB_HD void fft2_forward_dif_from(float2* x, int stride, int nb, const Fft2Plan& plan, const float2* __restrict__ tw,
                                 int tid, int nthr, Src& src) {
    PB_FFT2_DISPATCH2(fft2_dif_first, plan.radix[0], Src, x, plan.n, stride, nb, tw + plan.tw_off[0], tid, nthr, src);
    fft2_sync<WARP>();
    int L = plan.n / plan.radix[0];
    for (int s = 1; s < plan.ns; ++s) {
        const int R = plan.radix[s];
        PB_FFT2_DISPATCH(fft2_dif_stage, R, x, plan.n, stride, nb, L, tw + plan.tw_off[s], tid, nthr);
        fft2_sync<WARP>();
        L /= R;
    }
}

// Inverse-direction (DIT) transform whose last stage hands its results to `dst` (plan.ns >= 2).
template <class Dst, bool WARP = false>
PB_HD void fft2_forward_dit_to(float2* x, int stride, int nb, const Fft2Plan& plan, const float2* __restrict__ tw,
                               int tid, int nthr, Dst& dst, const float* premul = nullptr, int premode = 1) {
    int L = 1;
    for (int s = plan.ns - 1; s >= 1; --s) {
        const int R = plan.radix[s];
        L *= R;
        PB_FFT2_DISPATCH(fft2_dit_stage, R, x, plan.n, stride, nb, L, tw + plan.tw_off[s], tid, nthr,
                         (s == plan.ns - 1) ? premul : nullptr, premode);
        fft2_sync<WARP>();
    }
    PB_FFT2_DISPATCH2(fft2_dit_last, plan.radix[0], Dst, x, plan.n, stride, nb, tw + plan.tw_off[0], tid, nthr, dst);
}

// Frequency index k held by slot p after the DIF transform:
//   p = q_1 M_1 + q_2 M_2 + ... + q_s,  M_i = n / (R_1 ... R_i);   k = q_1 + R_1 (q_2 + R_2 (q_3 + ...))
PB_HD int fft2_freq_of_slot(int p, const Fft2Plan& plan) {
    int M = plan.n, k = 0, w = 1;
    for (int s = 0; s < plan.ns; ++s) {
        M /= plan.radix[s];
        const int q = p / M;
        p -= q * M;
        k += q * w;
        w *= plan.radix[s];
    }
    return k;
}

// Slot that holds frequency k after the DIF transform (inverse of fft2_freq_of_slot).
PB_HD int fft2_slot_of_freq(int k, const Fft2Plan& plan) {
    int M = plan.n, p = 0;
    for (int s = 0; s < plan.ns; ++s) {
        M /= plan.radix[s];
        const int q = k % plan.radix[s];
        k /= plan.radix[s];
        p += q * M;
    }
    return p;
}

// Average shared-memory conflict degree of one in-place stage: a half warp (64-bit accesses are
// served 16 lanes at a time) touches x[blk L + j + m M] for 16 consecutive butterflies.
inline double fft2_stage_conflicts(int n, int R, int M, int L) {
    const int bps = n / R;
    double tot = 0;
    int cnt = 0;
    for (int t0 = 0; t0 < bps && t0 < 512; t0 += 16) {
        for (int m = 0; m < R; ++m) {
            int hits[16] = {0};
            int worst = 0;
            for (int t = t0; t < t0 + 16 && t < bps; ++t) {
                const int blk = t / M, j = t - blk * M;
                const int bank = (blk * L + j + m * M) & 15;
                if (++hits[bank] > worst) worst = hits[bank];
            }
            tot += worst;
            ++cnt;
        }
    }
    return cnt ? tot / cnt : 1.0;
}

// Host: factor n into radices <= 16 with the fewest stages, then order them so that the
// strided late stages (small M) conflict least in shared memory -- in practice an odd radix
// goes last.  Returns 0 on success, -1 if n has a prime factor > 13 or needs more than
// PB_FFT2_MAX_STAGES stages.  (Not cached here: callers plan once per API call at most; api.cu
// keeps a small cache.)
// fills tw_off / tw_total from n and the radices
inline void fft2_plan_offsets(Fft2Plan* plan) {
    int off = 0, L = plan->n;
    for (int s = 0; s < PB_FFT2_MAX_STAGES; ++s) plan->tw_off[s] = 0;
    for (int s = 0; s < plan->ns; ++s) {
        const int M = L / plan->radix[s];
        plan->tw_off[s] = off;
        if (M > 1) off += (plan->radix[s] - 1) * M;
        L = M;
    }
    plan->tw_total = off > 0 ? off : 1;
}

// Entry i of the stage-twiddle table (host or device), computed in float64 and rounded once.
PB_HD float2 fft2_stage_twiddle(int i, const Fft2Plan& plan) {
    int L = plan.n;
    for (int s = 0; s < plan.ns; ++s) {
        const int R = plan.radix[s], M = L / R;
        const int cnt = (M > 1) ? (R - 1) * M : 0;
        if (i >= plan.tw_off[s] && i < plan.tw_off[s] + cnt) {
            const int e = i - plan.tw_off[s];
            const int q = e / M + 1, j = e - (q - 1) * M;
            const long long t = ((long long)j * q) % L;
            double sn, cs;
#if defined(__CUDA_ARCH__)
            sincospi(-2.0 * (double)t / (double)L, &sn, &cs);
#else
            sn = sin(-2.0 * 3.14159265358979323846 * (double)t / (double)L);
            cs = cos(-2.0 * 3.14159265358979323846 * (double)t / (double)L);
#endif
            return make_float2((float)cs, (float)sn);
        }
        L = M;
    }
    return make_float2(1.0f, 0.0f);
}

inline int make_fft2_plan_uncached(int n, Fft2Plan* plan);
inline int make_fft2_plan(int n, Fft2Plan* plan) {
    // small cache: the search below costs ~0.1-1 ms and API calls repeat the same lengths
    static std::mutex mu;
    static std::unordered_map<int, std::pair<int, Fft2Plan>> cache;
    {
        std::lock_guard<std::mutex> lock(mu);
        auto it = cache.find(n);
        if (it != cache.end()) {
            *plan = it->second.second;
            return it->second.first;
        }
    }
    Fft2Plan p;
    const int rc = make_fft2_plan_uncached(n, &p);
    std::lock_guard<std::mutex> lock(mu);
    cache[n] = std::make_pair(rc, p);
    *plan = p;
    return rc;
}
inline int make_fft2_plan_uncached(int n, Fft2Plan* plan) {
    if (n < 1) return -1;
    plan->n = n;
    plan->ns = 0;
    fft2_plan_offsets(plan);
    if (n == 1) return 0;
    static const int cand[] = {16, 15, 14, 13, 12, 11, 10, 9, 8, 7, 6, 5, 4, 3, 2};
    struct Best {
        int ns;
        double score;
        int radix[PB_FFT2_MAX_STAGES];
    } best;
    best.ns = PB_FFT2_MAX_STAGES + 1;
    best.score = 1e30;
    int cur[PB_FFT2_MAX_STAGES];
    struct Rec {
        static void score_perms(int n, int* a, int k, int depth, Best& best) {
            if (k == depth) {
                double sc = 0;
                int L = n;
                for (int s = 0; s < depth; ++s) {
                    sc += fft2_stage_conflicts(n, a[s], L / a[s], L);
                    L /= a[s];
                }
                sc -= 1e-3 * a[0];       // ties: the largest radix first
                if (depth < best.ns || (depth == best.ns && sc < best.score)) {
                    best.ns = depth;
                    best.score = sc;
                    for (int i = 0; i < depth; ++i) best.radix[i] = a[i];
                }
                return;
            }
            for (int i = k; i < depth; ++i) {
                bool dup = false;
                for (int q = k; q < i; ++q) dup = dup || a[q] == a[i];
                if (dup) continue;
                int t = a[k]; a[k] = a[i]; a[i] = t;
                score_perms(n, a, k + 1, depth, best);
                t = a[k]; a[k] = a[i]; a[i] = t;
            }
        }
        static void go(int n, int m, int depth, int start, int* cur, Best& best) {
            if (m == 1) {
                if (depth <= best.ns) {
                    int a[PB_FFT2_MAX_STAGES];
                    for (int i = 0; i < depth; ++i) a[i] = cur[i];
                    score_perms(n, a, 0, depth, best);
                }
                return;
            }
            if (depth >= PB_FFT2_MAX_STAGES || depth + 1 > best.ns) return;
            for (int ci = start; ci < 15; ++ci) {       // non-increasing radices: each multiset once
                const int r = cand[ci];
                if (m % r) continue;
                cur[depth] = r;
                go(n, m / r, depth + 1, ci, cur, best);
            }
        }
    };
    Rec::go(n, n, 0, 0, cur, best);
    if (best.ns > PB_FFT2_MAX_STAGES) return -1;
    plan->ns = best.ns;
    for (int i = 0; i < best.ns; ++i) plan->radix[i] = best.radix[i];
    fft2_plan_offsets(plan);
    return 0;
}

// Sum over the stages of the average shared-memory conflict degree (>= ns; lower is better).
inline double fft2_plan_score(const Fft2Plan& plan) {
    double sc = 0;
    int L = plan.n;
    for (int s = 0; s < plan.ns; ++s) {
        sc += fft2_stage_conflicts(plan.n, plan.radix[s], L / plan.radix[s], L);
        L /= plan.radix[s];
    }
    return sc;
}

// Rough per-element cost of a plan in "shared-memory passes": conflict degree plus arithmetic
// (real operations per element of each register butterfly / 16); used to choose torus lengths.
inline double fft2_plan_cost(const Fft2Plan& plan) {
    static const double flops[17] = {0, 0, 2, 5.3, 4, 8, 9, 13, 6.5, 11, 10, 88, 9, 104, 15, 13, 9};
    double c = fft2_plan_score(plan);
    for (int s = 0; s < plan.ns; ++s) c += flops[plan.radix[s]] / 16.0;
    return c;
}

}  // namespace pb
