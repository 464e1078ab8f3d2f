// Blur estimation, second-generation kernels (fast path for lengths whose prime factors are
// <= 13; estimate.cu keeps the any-length path).
//
// Reference being replaced: blur_estimation.gaussian_blur_estimation up to the directional
// maxima (polyblur/blur_estimation.py:18-65, 96-134) and filters.fourier_gradients
// (polyblur/filters.py:159-186), which is a per-row / per-column 1-D spectral derivative
// (SURVEY.md A.2).
//
//   k_rows2<EST>  : one CTA owns 2*nb adjacent rows.  Reads the C channels of those rows with
//                   128-bit loads, forms the channel mean g (blur_estimation.py:36-37), tracks
//                   min / max (blur_estimation.py:106-108), writes g, and writes d g / d x.
//   k_cols2<EST>  : one CTA owns 2*nb adjacent columns of g.  Computes d g / d y on chip, reads
//                   d g / d x back and reduces max |cos(phi_j) gx - sin(phi_j) gy| over its
//                   pixels for the 7 angles (blur_estimation.py:122-134); nothing is written
//                   but 7 atomics per CTA.
//   GRAD variants : the plain filters.fourier_gradients drop-in (planes in, gx / gy out).
//
// Two real sequences share one complex transform (real / imaginary part): the derivative
// multiplier i*omega is Hermitian, so the two results separate by themselves.
//
// HBM bytes per pixel of one estimate (C = 3): rows 12 read + 8 written, columns 8 read = 28
// (algorithmic: 12).  g and gx are scratch that mostly lives in the 126 MB L2
// when the batch is processed a few images at a time.
#include "fft.cuh"
#include <type_traits>

#include "fft2.cuh"
#include "fft2_static.cuh"
#include "kernels.cuh"

namespace pb {

// cos / sin of torch.linspace(0, pi, 7) exactly as torch (float32) evaluates them
// (blur_estimation.py:127-129) are spelled as immediates inside k_cols2 (same bit patterns as
// estimate.cu).

// omega[p] = angular frequency (filters.py:175-181 rounding) of the bin held by slot p after
// the DIF transform of length plan.n.
__global__ void k_fft2_omega(float* __restrict__ omega, Fft2Plan plan) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < plan.n) omega[p] = bin_omega(fft2_freq_of_slot(p, plan), plan.n);
}

// stage-twiddle table of a plan (fft2.cuh), float64 sincospi rounded once
__global__ void k_fft2_stage_tw(float2* __restrict__ stw, Fft2Plan plan) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < plan.tw_total) stw[i] = fft2_stage_twiddle(i, plan);
}

// Mirror units of a plan (used by the FFT engine's row passes, deconv_fft.cu): the last DIF stage (M = 1) leaves
// the frequencies k = f_A + (n / R_last) q, q < R_last, in the R_last contiguous slots of block A, and their
// mirrors n - k all lie in ONE other block B = mirror(A) (the digit-wise negation with carry of A's index), at
// position R_last - 1 - q (block 0 mirrors onto itself at (R_last - q) mod R_last).  One thread that owns a unit
// {A, B} therefore holds every pair {Z[k], Z[n - k]} it needs to separate (forward) or rebuild (inverse) the
// spectra of the two real rows packed into one complex transform, in registers.
//   out[0] = (number of units, R_last, n / R_last, 0);  out[1 + u] = (A, B, f_A, 0), A <= B, ascending A.
__device__ void build_mirror_units(const TableJob& J) {
    __shared__ int s_warp[8];
    __shared__ int s_base;
    const Fft2Plan& plan = J.plan;
    const int n = plan.n, RL = plan.radix[plan.ns - 1], NB = n / RL;
    int4* out = static_cast<int4*>(J.out);
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    if (tid == 0) s_base = 0;
    __syncthreads();
    for (int base = 0; base < NB; base += 256) {
        const int a = base + tid;
        int flag = 0, B = 0, fA = 0;
        if (a < NB) {
            fA = fft2_freq_of_slot(a * RL, plan);
            B = fft2_slot_of_freq((n - fA) % n, plan) / RL;
            flag = a <= B;
        }
        const unsigned bal = __ballot_sync(0xffffffffu, flag);
        if (lane == 0) s_warp[w] = __popc(bal);
        __syncthreads();
        int off = s_base + __popc(bal & ((1u << lane) - 1u));
        for (int i = 0; i < w; ++i) off += s_warp[i];
        if (flag) out[1 + off] = make_int4(a, B, fA, 0);
        __syncthreads();
        if (tid == 0) {
            int t = 0;
            for (int i = 0; i < 8; ++i) t += s_warp[i];
            s_base += t;
        }
        __syncthreads();
    }
    if (tid == 0) out[0] = make_int4(s_base, RL, NB, 0);
}

// all tables of one API call in one launch: blockIdx.y = job
__global__ void __launch_bounds__(256) k_table_jobs(const TableJobs jobs) {
    const TableJob& J = jobs.job[blockIdx.y];
    if (J.kind == TJ_UNITS) {
        if (blockIdx.x == 0) build_mirror_units(J);
        return;
    }
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= J.n) return;
    switch (J.kind) {
        case TJ_TWIDDLES: {
            double s, c;
            sincospi(-2.0 * (double)i / (double)J.n, &s, &c);
            static_cast<float2*>(J.out)[i] = make_float2((float)c, (float)s);
            break;
        }
        case TJ_STAGE_TW: static_cast<float2*>(J.out)[i] = fft2_stage_twiddle(i, J.plan); break;
        case TJ_OMEGA: static_cast<float*>(J.out)[i] = bin_omega(fft2_freq_of_slot(i, J.plan), J.plan.n); break;
        default:
            if (J.out) static_cast<int*>(J.out)[i] = fft2_slot_of_freq(i, J.plan);
            if (J.out2) static_cast<int*>(J.out2)[i] = fft2_freq_of_slot(i, J.plan);
            break;
    }
}

void TableJobs::add(int kind, int n, void* out, void* out2, const Fft2Plan* plan) {
    if (count >= PB_MAX_TABLE_JOBS) return;       // cannot happen: an entry point adds at most 10
    TableJob& j = job[count++];
    j.kind = kind;
    j.n = n;
    j.out = out;
    j.out2 = out2;
    if (plan) j.plan = *plan; else j.plan.n = n, j.plan.ns = 0;
}

int launch_table_jobs(const TableJobs& jobs, cudaStream_t stream) {
    if (jobs.count == 0) return PB_OK;
    int nmax = 1;
    for (int i = 0; i < jobs.count; ++i) nmax = jobs.job[i].n > nmax ? jobs.job[i].n : nmax;
    ProfScope prof(PROF_SETUP, stream);
    k_table_jobs<<<dim3((nmax + 255) / 256, jobs.count), 256, 0, stream>>>(jobs);
    PB_LAUNCH_CHECK("k_table_jobs");
    return PB_OK;
}

// s / 3 correctly rounded: q = RN(s y), r = s - 3 q (exact in an FMA), q' = RN(q + r y), y = RN(1/3)
__device__ __forceinline__ float pb_div3(float s) {
    const float y = 0x1.555556p-2f;
    const float q = __fmul_rn(s, y);
    const float r = __fmaf_rn(-3.0f, q, s);
    return __fmaf_rn(r, y, q);
}

#ifndef R2_THREADS
#define R2_THREADS 256
#endif
// shared-memory budget of a row tile, threads and column pairs of a column tile (tuning hooks)
#ifndef PB_R2_SMEM
#define PB_R2_SMEM (64 * 1024)
#endif
#ifndef PB_R2_LOAD_U
#define PB_R2_LOAD_U 2
#endif
#ifndef PB_C2_THREADS
#define PB_C2_THREADS 256
#endif
#ifndef PB_C2_NB
#define PB_C2_NB 8
#endif
// resident CTAs per SM the register allocation aims for (3 x 256 threads at <= 85 registers;
// measured: 2 CTAs at 128 registers is 2x slower, 4 CTAs do not fit the shared memory)
#ifndef PB_R2_MINB
#define PB_R2_MINB 3
#endif

template <bool EST, class SP>
__global__ void __launch_bounds__(R2_THREADS, PB_R2_MINB)
k_rows2(const float* __restrict__ img, float* __restrict__ gray, float* __restrict__ gx,
        unsigned* __restrict__ stats, int C, int H, int W, int nb, Fft2Plan plan,
        const float2* __restrict__ tw, const float* __restrict__ omega, const float* __restrict__ qrange) {
    extern __shared__ __align__(16) float2 sm2[];
    const int tid = threadIdx.x;
    const int y0 = blockIdx.x * 2 * nb;
    const int im = blockIdx.y;
    const size_t plane = (size_t)H * W;
    const float* src = img + (size_t)im * (EST ? C : 1) * plane;
    const float fC = (float)C;
    float lmin = INFINITY, lmax = -INFINITY;
    // q > 0: the input is the gray plane and is range-normalised with the quantiles while loading:
    // clamp_((x - lo) / (hi - lo), 0, 1)  (blur_estimation.py:92-93, 102-105)
    const bool qn = EST && qrange != nullptr;
    const float qlo = qn ? qrange[2 * im] : 0.f;
    const float qden = qn ? __fsub_rn(qrange[2 * im + 1], qlo) : 1.f;
#define PB_QNORM(v) fminf(fmaxf(__fdiv_rn(__fsub_rn((v), qlo), qden), 0.0f), 1.0f)
    // channel mean = sum / C, correctly rounded like the reference's (blur_estimation.py:36-37).  C = 3: Markstein's
    // FMA correction of s * RN(1/3) -- bit-identical to the IEEE division for every float (checked exhaustively,
    // tools/ubench/div3_exhaustive.c) at 3 instructions instead of ~12; C = 2, 4: exact scalings; else the division.
    const int cmode = (C == 3) ? 3 : (C == 2 || C == 4) ? 2 : (C == 1 ? 1 : 0);
    const float cscale = 1.0f / fC;
#define PB_CMEAN(s) (cmode == 3 ? pb_div3(s) : cmode == 2 ? __fmul_rn((s), cscale) : cmode == 1 ? (s) : __fdiv_rn((s), fC))

    if ((W & 3) == 0) {
        const int w4 = W >> 2;
        const float inv_w4 = 1.0f / (float)w4;
        // two row pairs per trip, every 128-bit load of the trip (2 rows x up to 3 channels x 2)
        // issued before the first use
        constexpr int CU = 3;                                   // channels unrolled for loads in flight
        constexpr int LU = PB_R2_LOAD_U;                        // row-pair items per trip
        for (int base = tid; base < nb * w4; base += LU * R2_THREADS) {
            float4 t[LU][2][CU];
            int pp[LU], xx[LU];
            bool ok[LU][2];
#pragma unroll
            for (int u = 0; u < LU; ++u) {
                const int idx = base + u * R2_THREADS;
                const bool live = idx < nb * w4;
                pp[u] = live ? fast_div(idx, w4, inv_w4) : 0;
                xx[u] = live ? (idx - pp[u] * w4) << 2 : 0;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int y = y0 + 2 * pp[u] + h;
                    ok[u][h] = live && y < H;
                    const float* q = src + (size_t)y * W + xx[u];
#pragma unroll
                    for (int c = 0; c < CU; ++c) {
                        t[u][h][c] = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (ok[u][h] && c < (EST ? C : 1))
                            t[u][h][c] = __ldg(reinterpret_cast<const float4*>(q + (size_t)c * plane));
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < LU; ++u) {
                if (base + u * R2_THREADS >= nb * w4) continue;
                float4 g[2];
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    g[h] = t[u][h][0];
                    if (EST && ok[u][h]) {
                        const int y = y0 + 2 * pp[u] + h;
#pragma unroll
                        for (int c = 1; c < CU; ++c) {
                            if (c < C) {
                                g[h].x = __fadd_rn(g[h].x, t[u][h][c].x);
                                g[h].y = __fadd_rn(g[h].y, t[u][h][c].y);
                                g[h].z = __fadd_rn(g[h].z, t[u][h][c].z);
                                g[h].w = __fadd_rn(g[h].w, t[u][h][c].w);
                            }
                        }
                        for (int c = CU; c < C; ++c) {      // more than 3 channels: plain loop
                            const float4 v = __ldg(reinterpret_cast<const float4*>(src + (size_t)c * plane + (size_t)y * W + xx[u]));
                            g[h].x = __fadd_rn(g[h].x, v.x);
                            g[h].y = __fadd_rn(g[h].y, v.y);
                            g[h].z = __fadd_rn(g[h].z, v.z);
                            g[h].w = __fadd_rn(g[h].w, v.w);
                        }
                        g[h].x = PB_CMEAN(g[h].x);
                        g[h].y = PB_CMEAN(g[h].y);
                        g[h].z = PB_CMEAN(g[h].z);
                        g[h].w = PB_CMEAN(g[h].w);
                        if (qn) g[h] = make_float4(PB_QNORM(g[h].x), PB_QNORM(g[h].y), PB_QNORM(g[h].z), PB_QNORM(g[h].w));
                        *reinterpret_cast<float4*>(gray + (size_t)im * plane + (size_t)y * W + xx[u]) = g[h];
                        lmin = fminf(lmin, fminf(fminf(g[h].x, g[h].y), fminf(g[h].z, g[h].w)));
                        lmax = fmaxf(lmax, fmaxf(fmaxf(g[h].x, g[h].y), fmaxf(g[h].z, g[h].w)));
                    }
                }
                float4* d = reinterpret_cast<float4*>(sm2 + (size_t)pp[u] * W + xx[u]);
                d[0] = make_float4(g[0].x, g[1].x, g[0].y, g[1].y);
                d[1] = make_float4(g[0].z, g[1].z, g[0].w, g[1].w);
            }
        }
    } else {
        const float inv_w = 1.0f / (float)W;
        for (int idx = tid; idx < nb * W; idx += R2_THREADS) {
            const int p = fast_div(idx, W, inv_w);
            const int x = idx - p * W;
            float g[2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int y = y0 + 2 * p + h;
                g[h] = 0.f;
                if (y < H) {
                    const float* q = src + (size_t)y * W + x;
                    g[h] = __ldg(q);
                    if (EST) {
                        for (int c = 1; c < C; ++c) g[h] = __fadd_rn(g[h], __ldg(q + (size_t)c * plane));
                        g[h] = PB_CMEAN(g[h]);
                        if (qn) g[h] = PB_QNORM(g[h]);
                        gray[(size_t)im * plane + (size_t)y * W + x] = g[h];
                        lmin = fminf(lmin, g[h]);
                        lmax = fmaxf(lmax, g[h]);
                    }
                }
            }
            sm2[(size_t)p * W + x] = make_float2(g[0], g[1]);
        }
    }
#undef PB_QNORM
#undef PB_CMEAN
    __syncthreads();
    // forward, multiply by i omega, inverse -- the two innermost stages fused in registers; SP = a
    // compile-time plan for the standard widths (fft2_static.cuh), else the run-time core
    if constexpr (std::is_same<SP, NoStaticPlan>::value)
        fft2_forward_mul_inverse(sm2, W, nb, plan, tw, tid, R2_THREADS, omega, 1);
    else
        s_forward_mul_inverse<SP, 1>(sm2, W, nb, tw, tid, R2_THREADS, omega);
    // inverse by forward transform of the swapped data: d(real part) = r.y / n, d(imag part) = r.x / n
    const float inv = 1.0f / (float)W;
    float* dst = gx + (size_t)im * plane;
    if ((W & 3) == 0) {
        const int w4 = W >> 2;
        const float inv_w4 = 1.0f / (float)w4;
        for (int idx = tid; idx < nb * w4; idx += R2_THREADS) {
            const int p = fast_div(idx, w4, inv_w4);
            const int x = (idx - p * w4) << 2;
            const float4* s = reinterpret_cast<const float4*>(sm2 + (size_t)p * W + x);
            const float4 a = s[0], b = s[1];
            const int y = y0 + 2 * p;
            if (y < H)
                *reinterpret_cast<float4*>(dst + (size_t)y * W + x) =
                    make_float4(a.y * inv, a.w * inv, b.y * inv, b.w * inv);
            if (y + 1 < H)
                *reinterpret_cast<float4*>(dst + (size_t)(y + 1) * W + x) =
                    make_float4(a.x * inv, a.z * inv, b.x * inv, b.z * inv);
        }
    } else {
        const float inv_w = 1.0f / (float)W;
        for (int idx = tid; idx < nb * W; idx += R2_THREADS) {
            const int p = fast_div(idx, W, inv_w);
            const int x = idx - p * W;
            const float2 z = sm2[(size_t)p * W + x];
            const int y = y0 + 2 * p;
            if (y < H) dst[(size_t)y * W + x] = z.y * inv;
            if (y + 1 < H) dst[(size_t)(y + 1) * W + x] = z.x * inv;
        }
    }
    if (EST) {
        lmin = warp_min(lmin);
        lmax = warp_max(lmax);
        if ((tid & 31) == 0) {
            atomicMin(&stats[im * PB_STATS_STRIDE + 0], f2ord(lmin));
            atomicMax(&stats[im * PB_STATS_STRIDE + 1], f2ord(lmax));
        }
    }
}

template <bool EST, int THREADS, class SP>
__global__ void __launch_bounds__(THREADS, (THREADS == PB_C2_THREADS ? PB_R2_MINB : 1))
k_cols2(const float* __restrict__ plane_in, const float* __restrict__ gx, float* __restrict__ gy,
        unsigned* __restrict__ stats, int H, int W, int nb, int stride, Fft2Plan plan,
        const float2* __restrict__ tw, const float* __restrict__ omega, int discard_saturation,
        const float* __restrict__ mask_src) {
    extern __shared__ __align__(16) float2 sm2[];
    __shared__ float red[THREADS / 32][8];
    const int tid = threadIdx.x;
    const int x0 = blockIdx.x * 2 * nb;
    const int im = blockIdx.y;
    const size_t plane = (size_t)H * W;
    const float* src = plane_in + (size_t)im * plane;
    const float inv_nb = 1.0f / (float)nb;
    const bool vec2 = (W & 1) == 0;

    for (int base = tid; base < H * nb; base += 4 * THREADS) {
        float2 g[4];
        int off[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int idx = base + u * THREADS;
            g[u] = make_float2(0.f, 0.f);
            off[u] = -1;
            if (idx < H * nb) {
                const int y = fast_div(idx, nb, inv_nb);
                const int p = idx - y * nb;
                const int x = x0 + 2 * p;
                if (vec2) {
                    if (x < W) g[u] = __ldcs(reinterpret_cast<const float2*>(src + (size_t)y * W + x));   // read once: evict-first
                } else {
                    if (x < W) g[u].x = __ldg(src + (size_t)y * W + x);
                    if (x + 1 < W) g[u].y = __ldg(src + (size_t)y * W + x + 1);
                }
                off[u] = p * stride + y;
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (off[u] >= 0) sm2[off[u]] = g[u];
    }
    __syncthreads();
    if constexpr (std::is_same<SP, NoStaticPlan>::value)
        fft2_forward_mul_inverse(sm2, stride, nb, plan, tw, tid, THREADS, omega, 1);
    else
        s_forward_mul_inverse<SP, 1>(sm2, stride, nb, tw, tid, THREADS, omega);
    const float inv = 1.0f / (float)H;

    if (!EST) {
        float* dst = gy + (size_t)im * plane;
        for (int idx = tid; idx < H * nb; idx += THREADS) {
            const int y = fast_div(idx, nb, inv_nb);
            const int p = idx - y * nb;
            const int x = x0 + 2 * p;
            const float2 z = sm2[(size_t)p * stride + y];
            const float2 o = make_float2(z.y * inv, z.x * inv);
            if (vec2) {
                if (x < W) *reinterpret_cast<float2*>(dst + (size_t)y * W + x) = o;
            } else {
                if (x < W) dst[(size_t)y * W + x] = o.x;
                if (x + 1 < W) dst[(size_t)y * W + x + 1] = o.y;
            }
        }
        return;
    }

    // cos / sin of the 7 angles as immediates (same bit patterns as c_cos7 / c_sin7 above)
    const float cs7[7] = {0x1.000000p+0f, 0x1.bb67aep-1f, 0x1.fffffep-2f, -0x1.777a5cp-25f,
                          -0x1.000002p-1f, -0x1.bb67aep-1f, -0x1.000000p+0f};
    const float sn7[7] = {0x0.0p+0f, 0x1.000000p-1f, 0x1.bb67aep-1f, 0x1.000000p+0f,
                          0x1.bb67aep-1f, 0x1.000002p-1f, -0x1.777a5cp-24f};
    float m[7];
#pragma unroll
    for (int j = 0; j < 7; ++j) m[j] = 0.0f;
    const float* gxp = gx + (size_t)im * plane;
    // saturation mask: un-normalised gray > 0.99 (blur_estimation.py:59, 83-88)
    const float* msk = (mask_src ? mask_src : plane_in) + (size_t)im * plane;
    // Four pixel pairs per trip, the d/dx (and mask) loads of all four in flight before the first use.  Pixels
    // that do not count (beyond the right edge, saturated) enter as gx = gy = 0: |0| never raises a maximum that
    // starts at 0, so the inner loop has no branches.  Angle 0 has cos = 1, sin = 0: the value is gx itself.
    for (int base = tid; base < H * nb; base += 4 * THREADS) {
        float2 zz[4], gxx[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int idx = base + u * THREADS;
            gxx[u] = make_float2(0.f, 0.f);
            zz[u] = make_float2(0.f, 0.f);
            if (idx < H * nb) {
                const int y = fast_div(idx, nb, inv_nb);
                const int p = idx - y * nb;
                const int x = x0 + 2 * p;
                if (x < W) {
                    const size_t o = (size_t)y * W + x;
                    float2 gr = make_float2(0.f, 0.f);
                    if (vec2) {
                        gxx[u] = __ldcs(reinterpret_cast<const float2*>(gxp + o));    // read once: evict-first (-3.6 % measured)
                        if (discard_saturation) gr = __ldg(reinterpret_cast<const float2*>(msk + o));
                    } else {
                        gxx[u].x = __ldg(gxp + o);
                        if (x + 1 < W) gxx[u].y = __ldg(gxp + o + 1);
                        if (discard_saturation) {
                            gr.x = __ldg(msk + o);
                            if (x + 1 < W) gr.y = __ldg(msk + o + 1);
                        }
                    }
                    zz[u] = sm2[(size_t)p * stride + y];
                    if (x + 1 >= W) zz[u].x = 0.f;               // .x carries d/dy of the pair's second column
                    if (discard_saturation) {
                        if (gr.x > 0.99f) { gxx[u].x = 0.f; zz[u].y = 0.f; }
                        if (gr.y > 0.99f) { gxx[u].y = 0.f; zz[u].x = 0.f; }
                    }
                }
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const float gyv[2] = {zz[u].y * inv, zz[u].x * inv};
            const float gxv[2] = {gxx[u].x, gxx[u].y};
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                m[0] = fmaxf(m[0], fabsf(gxv[h]));
#pragma unroll
                for (int j = 1; j < 7; ++j) {
                    const float v = __fsub_rn(__fmul_rn(cs7[j], gxv[h]), __fmul_rn(sn7[j], gyv[h]));
                    m[j] = fmaxf(m[j], fabsf(v));
                }
            }
        }
    }
    const int warp = tid >> 5, lane = tid & 31;
#pragma unroll
    for (int j = 0; j < 7; ++j) {
        const float v = warp_max(m[j]);
        if (lane == 0) red[warp][j] = v;
    }
    __syncthreads();
    if (tid < 7) {
        float v = 0.0f;
        for (int w = 0; w < THREADS / 32; ++w) v = fmaxf(v, red[w][tid]);
        atomicMax(&stats[im * PB_STATS_STRIDE + 2 + tid], __float_as_uint(v));
    }
}

// ---- host side ------------------------------------------------------------------------------

int launch_fft2_stage_tw(float2* stw, const Fft2Plan& plan, cudaStream_t stream) {
    ProfScope prof(PROF_SETUP, stream);
    k_fft2_stage_tw<<<(plan.tw_total + 255) / 256, 256, 0, stream>>>(stw, plan);
    PB_LAUNCH_CHECK("k_fft2_stage_tw");
    return PB_OK;
}

int launch_fft2_omega(float* omega, const Fft2Plan& plan, cudaStream_t stream) {
    ProfScope prof(PROF_SETUP, stream);
    k_fft2_omega<<<(plan.n + 255) / 256, 256, 0, stream>>>(omega, plan);
    PB_LAUNCH_CHECK("k_fft2_omega");
    return PB_OK;
}

template <typename KernelT>
static int set_smem2(KernelT kern, size_t bytes) {
    if (bytes > PB_SMEM_MAX - 2048) {
        set_error("FFT length needs %zu bytes of shared memory (> %d)", bytes, PB_SMEM_MAX - 2048);
        return PB_ERR_UNSUPPORTED;
    }
    PB_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    return PB_OK;
}

// sequences (pairs) per CTA: as many as fit `budget` bytes, at most max_nb, at least 1
static int pairs_for(int n, int max_nb, size_t budget) {
    int nb = (int)(budget / ((size_t)(n + 4) * sizeof(float2)));
    if (nb < 1) nb = 1;
    if (nb > max_nb) nb = max_nb;
    return nb;
}

bool fft2_supported(int H, int W) {
    Fft2Plan p;
    if (H < 2 || W < 2) return false;      // a length-1 transform has no stage to fold omega into
    if (make_fft2_plan(H, &p) || make_fft2_plan(W, &p)) return false;
    // one pair of rows / columns must fit the shared memory of a CTA
    const size_t lim = PB_SMEM_MAX - 2048;
    return (size_t)(W + 1) * sizeof(float2) <= lim && (size_t)(H + 1) * sizeof(float2) <= lim;
}

int launch_rows2(bool est, const float* img, float* gray, float* gx, unsigned* stats, int nimg, int C,
                 int H, int W, const Fft2Plan& planW, const float2* twW, const float* omegaW,
                 const float* qrange, cudaStream_t stream) {
    if (est && !qrange) {
        const int r3 = launch_rows3(img, gray, gx, stats, nimg, C, H, W, planW, twW, omegaW, stream);
        if (r3 != 1) return r3;
    }
    int nb = pairs_for(W, 8, PB_R2_SMEM);
    const int pairs_total = (H + 1) / 2;
    if (nb > pairs_total) nb = pairs_total;
    const size_t smem = (size_t)nb * W * sizeof(float2);
    dim3 grid((pairs_total + nb - 1) / nb, nimg);
    ProfScope prof(PROF_ROWS, stream);
    int rc;
#define PB_LAUNCH_ROWS2(E, SP)                                                                            \
    do {                                                                                                  \
        if ((rc = set_smem2(k_rows2<E, SP>, smem))) return rc;                                            \
        k_rows2<E, SP><<<grid, R2_THREADS, smem, stream>>>(img, gray, gx, stats, (E) ? C : 1, H, W, nb, planW, twW, \
                                                           omegaW, (E) ? qrange : nullptr);               \
    } while (0)
#define PB_LAUNCH_ROWS2_SP(SP)                                  \
    do {                                                        \
        if (est) PB_LAUNCH_ROWS2(true, SP); else PB_LAUNCH_ROWS2(false, SP); \
    } while (0)
    if (PlanW1920::matches(planW)) PB_LAUNCH_ROWS2_SP(PlanW1920);
    else if (PlanW3840::matches(planW)) PB_LAUNCH_ROWS2_SP(PlanW3840);
    else PB_LAUNCH_ROWS2_SP(NoStaticPlan);
#undef PB_LAUNCH_ROWS2_SP
#undef PB_LAUNCH_ROWS2
    PB_LAUNCH_CHECK("k_rows2");
    return PB_OK;
}

int launch_cols2(bool est, const float* plane_in, const float* gx, float* gy, unsigned* stats, int nimg,
                 int H, int W, const Fft2Plan& planH, const float2* twH, const float* omegaH,
                 int discard_saturation, const float* mask_src, cudaStream_t stream) {
    if (est) {
        const int r3 = launch_cols3(plane_in, gx, stats, nimg, H, W, planH, twH, omegaH, discard_saturation, mask_src, stream);
        if (r3 != 1) return r3;
    }
    // 8 pairs = 16 columns = 64-byte row segments; fall back to fewer when H is very long
    int nb = pairs_for(H, PB_C2_NB, 200 * 1024);
    // three resident CTAs of 4+ pairs (32-byte row segments) beat one big CTA of 8 (measured at H = 2160:
    // occupancy matters more than the segment length)
    const int nb3 = pairs_for(H, PB_C2_NB, 72 * 1024);
    if (nb3 >= 4) nb = nb3;
    const int pairs_total = (W + 1) / 2;
    if (nb > pairs_total) nb = pairs_total;
    // float2 per column pair: >= H and = 2 (mod 4), so that the 8 pairs x 2 rows a half-warp touches in
    // the load and reduce passes (address = pair * stride + row) fall into 16 different bank pairs
    const int stride = H + ((2 - H) & 3);
    const size_t smem = (size_t)nb * stride * sizeof(float2);
    dim3 grid((pairs_total + nb - 1) / nb, nimg);
    ProfScope prof(PROF_COLS, stream);
    int rc;
    const bool big = smem > 72 * 1024;
#define PB_LAUNCH_COLS2(E, T, SP)                                                                          \
    do {                                                                                                   \
        if ((rc = set_smem2(k_cols2<E, T, SP>, smem))) return rc;                                          \
        k_cols2<E, T, SP><<<grid, T, smem, stream>>>(plane_in, gx, gy, stats, H, W, nb, stride, planH, twH, \
                                                     omegaH, discard_saturation, mask_src);                \
    } while (0)
#define PB_LAUNCH_COLS2_SP(T, SP)                                               \
    do {                                                                        \
        if (est) PB_LAUNCH_COLS2(true, T, SP); else PB_LAUNCH_COLS2(false, T, SP); \
    } while (0)
    if (!big && PlanH1080::matches(planH)) PB_LAUNCH_COLS2_SP(PB_C2_THREADS, PlanH1080);
    else if (!big && PlanH2160::matches(planH)) PB_LAUNCH_COLS2_SP(PB_C2_THREADS, PlanH2160);
    else if (big) PB_LAUNCH_COLS2_SP(512, NoStaticPlan);
    else PB_LAUNCH_COLS2_SP(PB_C2_THREADS, NoStaticPlan);
#undef PB_LAUNCH_COLS2_SP
#undef PB_LAUNCH_COLS2
    PB_LAUNCH_CHECK("k_cols2");
    return PB_OK;
}

}  // namespace pb
