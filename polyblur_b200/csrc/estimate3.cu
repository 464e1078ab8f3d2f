// Blur estimation, third-generation kernels: the estimator's two passes with the first / last FFT
// stage working straight from / to global memory (the scheme of the FFT engine's second-generation
// row passes, deconv_fft.cu), so a row or column pair crosses shared memory once per INNER stage
// only -- no load pass, no store pass, no reduce pass:
//
//   k_rows3 : global RGB --(channel mean, min / max, stage 0 DIF in registers)--> smem
//             --(inner stages, i omega folded into the middle one)--> smem
//             --(stage 0 DIT in registers)--> d g / d x;  g is written while it is loaded.
//   k_cols3 : global g --(stage 0 DIF in registers)--> smem --(inner stages)--> smem
//             --(stage 0 DIT in registers)--> d g / d y, reduced on the spot with d g / d x (read
//             at the same pixels, issued before the butterfly) into the 7 directional maxima.
//
// Same arithmetic as estimate2.cu (same butterflies, same tables, same order of the channel
// sum): the directional maxima agree with the second generation to the last bit or two.
// Compile-time plans only (full-HD and 4K sides); everything else stays on estimate2.cu.
//
// Reference being replaced: blur_estimation.gaussian_blur_estimation up to the directional maxima
// (polyblur/blur_estimation.py:18-65, 96-134), filters.fourier_gradients (polyblur/filters.py:159-186).
#include <cstdlib>
#include <type_traits>

#include "fft2_static.cuh"
#include "kernels.cuh"

namespace pb {

// s / 3 correctly rounded (see estimate2.cu: Markstein's FMA correction, checked exhaustively)
static __device__ __forceinline__ float div3_rn(float s) {
    const float y = 0x1.555556p-2f;
    const float q = __fmul_rn(s, y);
    const float r = __fmaf_rn(-3.0f, q, s);
    return __fmaf_rn(r, y, q);
}
static __device__ __forceinline__ float gray3(float a, float b, float c) {
    return div3_rn(__fadd_rn(__fadd_rn(a, b), c));        // (r + g + b) / 3 in the reference's order
}


// ---------------------------------------------------------------------------------------------
// rows: CTA = nb row pairs (rows y0 .. y0 + 2 nb - 1) of one image, C = 3.
// A thread owns WIDE adjacent columns j .. j + WIDE - 1 of a pair: WIDE butterflies side by side, so
// global accesses are 8 bytes and shared-memory accesses 16 bytes wide when WIDE = 2.
// ---------------------------------------------------------------------------------------------
template <class SP, int WIDE, int THREADS, int MINB, bool PADL>
__global__ void __launch_bounds__(THREADS, MINB)
k_rows3(const float* __restrict__ img, float* __restrict__ gray, float* __restrict__ gx,
        unsigned* __restrict__ stats, int H, int nb, const float2* __restrict__ tw,
        const float* __restrict__ omega, int rev) {
    constexpr int W = SP::n, NS = SP::ns, R0 = SP::R(0), M0 = W / R0, JM = M0 / WIDE;
    // padded stage-1 blocks (fft2_static.cuh: LayoutPad1) for the three-stage plans; RS = float2 per row pair
    using LY = typename std::conditional<(PADL && NS == 3), PadFor<SP>, LayoutFlat>::type;
    constexpr int M0P = M0 + LY::PAD, RS = W + R0 * LY::PAD;
    constexpr bool V4 = WIDE == 2 && (M0P % 2 == 0);        // 16-byte shared-memory accesses need even block starts
    static_assert(M0 % WIDE == 0 && (WIDE == 1 || WIDE == 2), "column groups must tile the first stage");
    static_assert(NS >= 3, "needs an inner stage besides the fused middle one");
    extern __shared__ __align__(16) float2 sm2[];
    const int tid = threadIdx.x;
    // rev: the grid walks the batch backwards (last image, last rows first).  The kernel that produced the iterate
    // (P3 / the narrow engine) walked it forwards, so the ~100 MB it wrote last are still in the 126 MB L2 when
    // this kernel starts with them; it ends at image 0, where the column pass (forwards) picks g and d g / d x up.
    const int y0 = (rev ? gridDim.x - 1 - blockIdx.x : blockIdx.x) * 2 * nb;
    const int im = rev ? gridDim.y - 1 - blockIdx.y : blockIdx.y;
    const size_t plane = (size_t)H * W;
    const float* src = img + (size_t)im * 3 * plane;
    float* gdst = gray + (size_t)im * plane;
    const float2* stw0 = tw + SP::tw_off(0);
    float lmin = INFINITY, lmax = -INFINITY;

    // ---- stage 0 (DIF), fed from global memory -------------------------------------------------
    for (int idx = tid; idx < nb * JM; idx += THREADS) {
        const int f = idx / JM;
        const int j = (idx - f * JM) * WIDE;
        const int y = y0 + 2 * f;
        if (y >= H) continue;
        const bool okb = y + 1 < H;
        const float* pa = src + (size_t)y * W + j;
        const float* pb = okb ? pa + W : pa;
        float* ga = gdst + (size_t)y * W + j;
        float2 v[WIDE][R0];
#pragma unroll
        for (int m = 0; m < R0; ++m) {
            if constexpr (WIDE == 2) {
                const float2 a0 = __ldg(reinterpret_cast<const float2*>(pa + m * M0));
                const float2 a1 = __ldg(reinterpret_cast<const float2*>(pa + m * M0 + plane));
                const float2 a2 = __ldg(reinterpret_cast<const float2*>(pa + m * M0 + 2 * plane));
                const float2 b0 = __ldg(reinterpret_cast<const float2*>(pb + m * M0));
                const float2 b1 = __ldg(reinterpret_cast<const float2*>(pb + m * M0 + plane));
                const float2 b2 = __ldg(reinterpret_cast<const float2*>(pb + m * M0 + 2 * plane));
                const float2 g0 = make_float2(gray3(a0.x, a1.x, a2.x), gray3(a0.y, a1.y, a2.y));
                float2 g1 = make_float2(gray3(b0.x, b1.x, b2.x), gray3(b0.y, b1.y, b2.y));
                *reinterpret_cast<float2*>(ga + m * M0) = g0;
                lmin = fminf(lmin, fminf(g0.x, g0.y));
                lmax = fmaxf(lmax, fmaxf(g0.x, g0.y));
                if (okb) {
                    *reinterpret_cast<float2*>(ga + W + m * M0) = g1;
                    lmin = fminf(lmin, fminf(g1.x, g1.y));
                    lmax = fmaxf(lmax, fmaxf(g1.x, g1.y));
                } else {
                    g1 = make_float2(0.f, 0.f);
                }
                v[0][m] = make_float2(g0.x, g1.x);
                v[1][m] = make_float2(g0.y, g1.y);
            } else {
                const float g0 = gray3(__ldg(pa + m * M0), __ldg(pa + m * M0 + plane), __ldg(pa + m * M0 + 2 * plane));
                float g1 = gray3(__ldg(pb + m * M0), __ldg(pb + m * M0 + plane), __ldg(pb + m * M0 + 2 * plane));
                ga[m * M0] = g0;
                lmin = fminf(lmin, g0);
                lmax = fmaxf(lmax, g0);
                if (okb) {
                    ga[W + m * M0] = g1;
                    lmin = fminf(lmin, g1);
                    lmax = fmaxf(lmax, g1);
                } else {
                    g1 = 0.f;
                }
                v[0][m] = make_float2(g0, g1);
            }
        }
#pragma unroll
        for (int w = 0; w < WIDE; ++w) Dft<R0>::run(v[w]);
        float2* p = sm2 + (size_t)f * RS + j;
        if constexpr (WIDE == 2) {
            *reinterpret_cast<float4*>(p) = make_float4(v[0][0].x, v[0][0].y, v[1][0].x, v[1][0].y);
#pragma unroll
            for (int q = 1; q < R0; ++q) {
                const float4 t = __ldg(reinterpret_cast<const float4*>(stw0 + (q - 1) * M0 + j));
                const float2 r0 = c_mul(v[0][q], make_float2(t.x, t.y));
                const float2 r1 = c_mul(v[1][q], make_float2(t.z, t.w));
                if constexpr (V4) {
                    *reinterpret_cast<float4*>(p + q * M0P) = make_float4(r0.x, r0.y, r1.x, r1.y);
                } else {
                    p[q * M0P] = r0;
                    p[q * M0P + 1] = r1;
                }
            }
        } else {
            p[0] = v[0][0];
#pragma unroll
            for (int q = 1; q < R0; ++q) p[q * M0P] = c_mul(v[0][q], __ldg(stw0 + (q - 1) * M0 + j));
        }
    }
    __syncthreads();

    // ---- inner stages: DIF 1 .. NS-2, [DIF NS-1, i omega, DIT NS-1] in registers, DIT NS-2 .. 1 ---
    SDifRun<SP, 1, NS - 2, false, LY>::run(sm2, RS, nb, tw, tid, THREADS);
    s_mid_stage<SP::R(NS - 1), W, 1, LY>(sm2, RS, nb, tid, THREADS, omega);
    __syncthreads();
    SDitRun<SP, NS - 2, NS - 2, false, LY>::run(sm2, RS, nb, tw, tid, THREADS);

    // ---- stage 0 (DIT), drained to global memory: r = DFT(swap(.)): row a = r.y / n, row b = r.x / n
    const float inv = 1.0f / (float)W;
    float* gxd = gx + (size_t)im * plane;
    for (int idx = tid; idx < nb * JM; idx += THREADS) {
        const int f = idx / JM;
        const int j = (idx - f * JM) * WIDE;
        const int y = y0 + 2 * f;
        if (y >= H) continue;
        const bool okb = y + 1 < H;
        const float2* p = sm2 + (size_t)f * RS + j;
        float2 v[WIDE][R0];
        if constexpr (WIDE == 2) {
            const float4 z0 = *reinterpret_cast<const float4*>(p);
            v[0][0] = make_float2(z0.x, z0.y);
            v[1][0] = make_float2(z0.z, z0.w);
#pragma unroll
            for (int q = 1; q < R0; ++q) {
                float4 z;
                if constexpr (V4) {
                    z = *reinterpret_cast<const float4*>(p + q * M0P);
                } else {
                    const float2 za = p[q * M0P], zb = p[q * M0P + 1];
                    z = make_float4(za.x, za.y, zb.x, zb.y);
                }
                const float4 t = __ldg(reinterpret_cast<const float4*>(stw0 + (q - 1) * M0 + j));
                v[0][q] = c_mul(make_float2(z.x, z.y), make_float2(t.x, t.y));
                v[1][q] = c_mul(make_float2(z.z, z.w), make_float2(t.z, t.w));
            }
        } else {
            v[0][0] = p[0];
#pragma unroll
            for (int q = 1; q < R0; ++q) v[0][q] = c_mul(p[q * M0P], __ldg(stw0 + (q - 1) * M0 + j));
        }
#pragma unroll
        for (int w = 0; w < WIDE; ++w) Dft<R0>::run(v[w]);
        float* da = gxd + (size_t)y * W + j;
#pragma unroll
        for (int m = 0; m < R0; ++m) {
            if constexpr (WIDE == 2) {
                *reinterpret_cast<float2*>(da + m * M0) = make_float2(v[0][m].y * inv, v[1][m].y * inv);
                if (okb) *reinterpret_cast<float2*>(da + W + m * M0) = make_float2(v[0][m].x * inv, v[1][m].x * inv);
            } else {
                da[m * M0] = v[0][m].y * inv;
                if (okb) da[W + m * M0] = v[0][m].x * inv;
            }
        }
    }
    lmin = warp_min(lmin);
    lmax = warp_max(lmax);
    if ((tid & 31) == 0) {
        atomicMin(&stats[im * PB_STATS_STRIDE + 0], f2ord(lmin));
        atomicMax(&stats[im * PB_STATS_STRIDE + 1], f2ord(lmax));
    }
}

// ---------------------------------------------------------------------------------------------
// columns: CTA = NB column pairs (2 NB adjacent columns) of one gray plane; consecutive threads take
// consecutive pairs of one row (8 NB contiguous bytes), then the next row.
// ---------------------------------------------------------------------------------------------
template <class SP, int NB, int THREADS, int MINB, bool PADL>
__global__ void __launch_bounds__(THREADS, MINB)
k_cols3(const float* __restrict__ g, const float* __restrict__ gx, unsigned* __restrict__ stats, int W, int stride,
        const float2* __restrict__ tw, const float* __restrict__ omega, int discard_saturation,
        const float* __restrict__ mask_src) {
    constexpr int H = SP::n, NS = SP::ns, R0 = SP::R(0), M0 = H / R0;
    using LY = typename std::conditional<(PADL && NS == 3), PadFor<SP>, LayoutFlat>::type;
    constexpr int M0P = M0 + LY::PAD;
    static_assert(NS >= 3, "needs an inner stage besides the fused middle one");
    static_assert((NB & (NB - 1)) == 0, "pairs per CTA: a power of two");
    extern __shared__ __align__(16) float2 sm2[];
    __shared__ float red[THREADS / 32][8];
    const int tid = threadIdx.x;
    const int x0 = blockIdx.x * 2 * NB;
    const int im = blockIdx.y;
    const size_t plane = (size_t)H * W;
    const float* src = g + (size_t)im * plane;
    const float2* stw0 = tw + SP::tw_off(0);

    // ---- stage 0 (DIF), fed from global memory -------------------------------------------------
    for (int idx = tid; idx < NB * M0; idx += THREADS) {
        const int j = idx / NB, f = idx & (NB - 1);
        const int x = x0 + 2 * f;
        float2 v[R0];
        if (x < W) {
            const float* ps = src + (size_t)j * W + x;
#pragma unroll
            for (int m = 0; m < R0; ++m) v[m] = __ldcs(reinterpret_cast<const float2*>(ps + (size_t)m * M0 * W));
        } else {
#pragma unroll
            for (int m = 0; m < R0; ++m) v[m] = make_float2(0.f, 0.f);
        }
        Dft<R0>::run(v);
        float2* p = sm2 + (size_t)f * stride + j;
        p[0] = v[0];
#pragma unroll
        for (int q = 1; q < R0; ++q) p[q * M0P] = c_mul(v[q], __ldg(stw0 + (q - 1) * M0 + j));
    }
    __syncthreads();

    SDifRun<SP, 1, NS - 2, false, LY>::run(sm2, stride, NB, tw, tid, THREADS);
    s_mid_stage<SP::R(NS - 1), H, 1, LY>(sm2, stride, NB, tid, THREADS, omega);
    __syncthreads();
    SDitRun<SP, NS - 2, NS - 2, false, LY>::run(sm2, stride, NB, tw, tid, THREADS);

    // ---- stage 0 (DIT) in registers + the 7 directional maxima (blur_estimation.py:122-134) ------
    // cos / sin of torch.linspace(0, pi, 7) as torch (float32) evaluates them (same bit patterns as estimate2.cu)
    const float cs7[7] = {0x1.000000p+0f, 0x1.bb67aep-1f, 0x1.fffffep-2f, -0x1.777a5cp-25f,
                          -0x1.000002p-1f, -0x1.bb67aep-1f, -0x1.000000p+0f};
    const float sn7[7] = {0x0.0p+0f, 0x1.000000p-1f, 0x1.bb67aep-1f, 0x1.000000p+0f,
                          0x1.bb67aep-1f, 0x1.000002p-1f, -0x1.777a5cp-24f};
    float mx[7];
#pragma unroll
    for (int a = 0; a < 7; ++a) mx[a] = 0.0f;
    const float inv = 1.0f / (float)H;
    const float* gxp = gx + (size_t)im * plane;
    const float* msk = (mask_src ? mask_src : g) + (size_t)im * plane;   // un-normalised gray > 0.99 is saturated
    for (int idx = tid; idx < NB * M0; idx += THREADS) {
        const int j = idx / NB, f = idx & (NB - 1);
        const int x = x0 + 2 * f;
        if (x >= W) continue;
        const size_t o = (size_t)j * W + x;
        // d g / d x of the same pixels: all loads in flight before the butterfly when the registers allow it
        constexpr bool PRE = R0 <= 10 || (THREADS * MINB <= 576);   // the registers of a second butterfly must be there
        float2 gxv[PRE ? R0 : 1];
        if constexpr (PRE) {
#pragma unroll
            for (int m = 0; m < R0; ++m) gxv[m] = __ldcs(reinterpret_cast<const float2*>(gxp + o + (size_t)m * M0 * W));
        }
        const float2* p = sm2 + (size_t)f * stride + j;
        float2 v[R0];
        v[0] = p[0];
#pragma unroll
        for (int q = 1; q < R0; ++q) v[q] = c_mul(p[q * M0P], __ldg(stw0 + (q - 1) * M0 + j));
        Dft<R0>::run(v);
#pragma unroll
        for (int m = 0; m < R0; ++m) {
            float gyv[2] = {v[m].y * inv, v[m].x * inv};       // .y: first column of the pair, .x: second
            float2 gxm;
            if constexpr (PRE) gxm = gxv[m];
            else gxm = __ldcs(reinterpret_cast<const float2*>(gxp + o + (size_t)m * M0 * W));
            float gxw[2] = {gxm.x, gxm.y};
            if (discard_saturation) {
                const float2 gr = __ldg(reinterpret_cast<const float2*>(msk + o + (size_t)m * M0 * W));
                if (gr.x > 0.99f) { gxw[0] = 0.f; gyv[0] = 0.f; }
                if (gr.y > 0.99f) { gxw[1] = 0.f; gyv[1] = 0.f; }
            }
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                mx[0] = fmaxf(mx[0], fabsf(gxw[h]));
#pragma unroll
                for (int a = 1; a < 7; ++a) {
                    const float d = __fsub_rn(__fmul_rn(cs7[a], gxw[h]), __fmul_rn(sn7[a], gyv[h]));
                    mx[a] = fmaxf(mx[a], fabsf(d));
                }
            }
        }
    }
    const int warp = tid >> 5, lane = tid & 31;
#pragma unroll
    for (int a = 0; a < 7; ++a) {
        const float r = warp_max(mx[a]);
        if (lane == 0) red[warp][a] = r;
    }
    __syncthreads();
    if (tid < 7) {
        float r = 0.0f;
        for (int w = 0; w < THREADS / 32; ++w) r = fmaxf(r, red[w][tid]);
        atomicMax(&stats[im * PB_STATS_STRIDE + 2 + tid], __float_as_uint(r));
    }
}

// ---- host side ------------------------------------------------------------------------------
static int env3(const char* name, int dflt) {
    const char* v = getenv(name);
    return v ? atoi(v) : dflt;
}
static bool est_gen3() {
    static const int on = env3("PB_EST_GEN", 3);
    return on >= 3;
}

// the single-image 12000 x 9000 configuration (BASELINE C4): the run-time planner's radix orders
using PlanW12000 = StaticPlan<12000, 10, 5, 16, 15>;
using PlanH9000 = StaticPlan<9000, 15, 15, 8, 5>;

template <typename KernelT>
static int set_smem3(KernelT kern, size_t bytes) {
    PB_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    return PB_OK;
}

// returns PB_OK when launched, 1 when this generation does not cover the case (the caller falls back)
int launch_rows3(const float* img, float* gray, float* gx, unsigned* stats, int nimg, int C, int H, int W,
                 const Fft2Plan& planW, const float2* twW, const float* omegaW, cudaStream_t stream) {
    if (!est_gen3() || C != 3) return 1;
    static const int rev3 = env3("PB_REVERSE", 1);
    const int pairs_total = (H + 1) / 2;
    int rc;
    // nb row pairs per CTA: what fits BUDGET bytes of shared memory (at least one)
#define PB_ROWS3(SP, WIDE, THREADS, MINB, BUDGET, PADL)                                                        \
    do {                                                                                                       \
        int nb = (BUDGET) / (int)(sizeof(float2) * SP::n);                                                     \
        if (nb < 1) nb = 1;                                                                                    \
        if (nb > pairs_total) nb = pairs_total;                                                                \
        const size_t len = SP::n + ((PADL) && SP::ns == 3 ? SP::R(0) * PadFor<SP>::PAD : 0);                   \
        const size_t smem = (size_t)nb * len * sizeof(float2);                                                 \
        if ((rc = set_smem3(k_rows3<SP, WIDE, THREADS, MINB, PADL>, smem))) return rc;                         \
        ProfScope prof(PROF_ROWS, stream);                                                                     \
        k_rows3<SP, WIDE, THREADS, MINB, PADL><<<dim3((pairs_total + nb - 1) / nb, nimg), THREADS, smem, stream>>>( \
            img, gray, gx, stats, H, nb, twW, omegaW, rev3);                                                   \
        PB_LAUNCH_CHECK("k_rows3");                                                                            \
        return PB_OK;                                                                                          \
    } while (0)
    // padded stage-1 blocks (conflict-free middle stage): 1 = the 4K lengths, 2 = the 1080p lengths too (measured
    // neutral there: 0.81 / 0.80 ms per step either way; 4K rows 0.86 -> 0.83, columns 1.35 -> 1.32)
    static const int pad3 = env3("PB_E3_PAD", 1);
    // 1920: two CTAs per SM at 128 registers (the 48 64-bit loads of a butterfly pair in flight together) measured
    // 0.77 against 0.80 ms per step at three CTAs of 80; 3840 (16-point first stage, one column per thread): equal
    static const int rminb = env3("PB_E3_RMINB", 2);
    if (PlanW1920::matches(planW)) {
        if (rminb == 2) PB_ROWS3(PlanW1920, 2, 256, 2, 64 * 1024, false);
        if (pad3 >= 2) PB_ROWS3(PlanW1920, 2, 256, 3, 64 * 1024, true);
        PB_ROWS3(PlanW1920, 2, 256, 3, 64 * 1024, false);
    }
    if (PlanW3840::matches(planW)) {
        if (pad3) PB_ROWS3(PlanW3840, 1, 256, 3, 64 * 1024, true);
        PB_ROWS3(PlanW3840, 1, 256, 3, 64 * 1024, false);
    }
    // one pair (94 KB), two CTAs per SM (384 threads per CTA measured the same as 256)
    if (PlanW12000::matches(planW)) PB_ROWS3(PlanW12000, 2, 256, 2, 100 * 1024, false);
#undef PB_ROWS3
    return 1;
}

int launch_cols3(const float* g, const float* gx, unsigned* stats, int nimg, int H, int W, const Fft2Plan& planH,
                 const float2* twH, const float* omegaH, int discard_saturation, const float* mask_src,
                 cudaStream_t stream) {
    if (!est_gen3() || (W & 1)) return 1;
    const int pairs_total = W / 2;
    int rc;
#define PB_COLS3(SP, NB, THREADS, MINB, PADL)                                                                  \
    do {                                                                                                       \
        const int len = SP::n + ((PADL) && SP::ns == 3 ? SP::R(0) * PadFor<SP>::PAD : 0);                      \
        const int stride = len + ((2 - len) & 3);          /* = 2 (mod 4) float2: see launch_cols2 */           \
        const size_t smem = (size_t)(NB) * stride * sizeof(float2);                                            \
        if ((rc = set_smem3(k_cols3<SP, NB, THREADS, MINB, PADL>, smem))) return rc;                           \
        ProfScope prof(PROF_COLS, stream);                                                                     \
        k_cols3<SP, NB, THREADS, MINB, PADL><<<dim3((pairs_total + (NB) - 1) / (NB), nimg), THREADS, smem, stream>>>(\
            g, gx, stats, W, stride, twH, omegaH, discard_saturation, mask_src);                               \
        PB_LAUNCH_CHECK("k_cols3");                                                                            \
        return PB_OK;                                                                                          \
    } while (0)
    // (192 threads, which divide the butterfly counts of the 1080 / 2160 stages evenly where 256 leave a quarter of a
    // trip idle, measured the same as 256: the trips are not what binds)
    static const int pad3 = env3("PB_E3_PAD", 1);
    if (PlanH1080::matches(planH)) {
        if (pad3 >= 2) PB_COLS3(PlanH1080, 8, 256, 3, true);
        PB_COLS3(PlanH1080, 8, 256, 3, false);
    }
    if (PlanH2160::matches(planH)) {
        // 15-point first / last stage: the d g / d x loads of a butterfly are all in flight before it only when a thread may
        // use ~110 registers.  Measured (8 x 4K, ms per step): 256 threads x 3 CTAs at 80 registers (loads issued late,
        // 108 bytes of spills) 1.32, 192 x 3 at 96 registers 1.19, 256 x 2 at 128 registers 0.91.
        static const int t2160 = env3("PB_E3_T2160", 512);
        if (t2160 == 192) PB_COLS3(PlanH2160, 4, 192, 3, true);
        if (t2160 == 512) PB_COLS3(PlanH2160, 4, 256, 2, true);
        if (pad3) PB_COLS3(PlanH2160, 4, 256, 3, true);
        PB_COLS3(PlanH2160, 4, 256, 3, false);
    }
    // one pair of 9000 rows is 70 KB: two pairs in one CTA of 512 threads (16-byte row segments).  Measured and left out:
    // one pair in each of two resident CTAs of 256 threads (8-byte segments; C4 step 5.31 against 5.09 ms), and a
    // persistent variant of this kernel that runs the last stage of an item and the first stage of the CTA's next
    // item in one loop (same thread, same shared-memory slots, so no barrier and the next item's loads are in flight
    // during the reduce): C2 0.88 against 0.82 ms per step, C4 4.95 against 5.07 ms -- the loads are not what binds.
    if (PlanH9000::matches(planH)) PB_COLS3(PlanH9000, 2, 512, 1, false);
#undef PB_COLS3
    return 1;
}

}  // namespace pb
