// Compile-time specialisation of the fft2.cuh stages for the lengths Polyblur meets most (full-HD
// and 4K rows / columns and their FFT-engine tori).  Same algorithm, same tables, same results as
// the run-time core: only the radices, strides and loop bounds become template constants, so the
// butterfly addresses and twiddle offsets are immediates, the index divisions turn into
// multiply-shifts and every loop has a known trip count.  A kernel instantiated with a StaticPlan
// must be launched with the run-time plan of the same radices (the host checks it): the twiddle,
// omega and permutation tables are still generated from the run-time plan.
#pragma once
#include "fft2.cuh"

namespace pb {

template <int N_, int... RS>
struct StaticPlan {
    static constexpr int n = N_;
    static constexpr int ns = sizeof...(RS);
    static constexpr PB_HDC int R(int s) {
        const int r[] = {RS...};
        return r[s];
    }
    // sub-length on entry to DIF stage s
    static constexpr PB_HDC int L(int s) {
        int l = N_;
        for (int i = 0; i < s; ++i) l /= R(i);
        return l;
    }
    // offset of stage s in the stage-twiddle table (same rule as fft2_plan_offsets)
    static constexpr PB_HDC int tw_off(int s) {
        int off = 0, l = N_;
        for (int i = 0; i < s; ++i) {
            const int m = l / R(i);
            if (m > 1) off += (R(i) - 1) * m;
            l = m;
        }
        return off;
    }
    static bool matches(const Fft2Plan& p) {
        if (p.n != N_ || p.ns != ns) return false;
        for (int s = 0; s < ns; ++s)
            if (p.radix[s] != R(s)) return false;
        return true;
    }
};

struct NoStaticPlan {};     // kernels instantiated with this use the run-time core

// ---- shared-memory layout of one sequence ------------------------------------------------------------------
// A stage with M < 16 makes the 16 lanes of a half warp straddle blocks: butterfly t = (blk, j) touches
// blk L + j + m M, and unless L = M (mod 16) two lanes of the half warp meet in one bank pair -- every such access
// then costs two wavefronts (ncu: 30 % of the shared-memory wavefronts of the row / column kernels were excess).
// In a three-stage plan that stage is stage 1 (M = the last radix).  LayoutPad1 pads the R0 blocks of length
// L1 = n / R0 that stage 0 produces by PAD float2 so that L1 + PAD = M1 (mod 16): slot t of stage 1 then falls into
// bank pair t mod 16 for every butterfly input, conflict-free; natural slot i lives at i + (i / L1) PAD.
struct LayoutFlat {
    static constexpr int L1 = 1 << 30, PAD = 0;
    static PB_HDC int off(int i) { return i; }
};
template <int L1_, int PAD_>
struct LayoutPad1 {
    static constexpr int L1 = L1_, PAD = PAD_;
    static PB_HDC int off(int i) { return i + (i / L1_) * PAD_; }
};
// the padded layout of a plan with >= 3 stages, and the float2 one sequence occupies in it
template <class P>
using PadFor = LayoutPad1<P::L(1), ((P::L(2) - P::L(1)) % 16 + 16) % 16>;
template <class P, class LY>
constexpr int padded_len() { return P::n + P::R(0) * LY::PAD; }

template <int R, int M, int N, class LY = LayoutFlat>
PB_HD void s_dif_stage(float2* x, int stride, int nb, const float2* __restrict__ stw, int tid, int nthr) {
    static_assert(R * M <= LY::L1, "a padded layout serves the stages inside the padded blocks only");
    constexpr int L = R * M, bps = N / R;
    const int total = nb * bps;
    for (int idx = tid; idx < total; idx += nthr) {
        const int f = idx / bps;
        const int rem = idx - f * bps;
        const int blk = rem / M;
        const int j = rem - blk * M;
        float2* p = x + f * stride + LY::off(blk * L + j);
        float2 v[R];
#pragma unroll
        for (int m = 0; m < R; ++m) v[m] = p[m * M];
        Dft<R>::run(v);
        p[0] = v[0];
#pragma unroll
        for (int q = 1; q < R; ++q) p[q * M] = (M == 1) ? v[q] : c_mul(v[q], PB_LDG(stw + (q - 1) * M + j));
    }
}

template <int R, int M, int N, class LY = LayoutFlat>
PB_HD void s_dit_stage(float2* x, int stride, int nb, const float2* __restrict__ stw, int tid, int nthr) {
    static_assert(R * M <= LY::L1, "a padded layout serves the stages inside the padded blocks only");
    constexpr int L = R * M, bps = N / R;
    const int total = nb * bps;
    for (int idx = tid; idx < total; idx += nthr) {
        const int f = idx / bps;
        const int rem = idx - f * bps;
        const int blk = rem / M;
        const int j = rem - blk * M;
        float2* p = x + f * stride + LY::off(blk * L + j);
        float2 v[R];
        v[0] = p[0];
#pragma unroll
        for (int q = 1; q < R; ++q) v[q] = (M == 1) ? p[q] : c_mul(p[q * M], PB_LDG(stw + (q - 1) * M + j));
        Dft<R>::run(v);
#pragma unroll
        for (int m = 0; m < R; ++m) p[m * M] = v[m];
    }
}

// last DIF stage + pointwise multiplier + first DIT stage (see fft2_mid_stage)
template <int R, int N, int PREMODE, class LY = LayoutFlat>
PB_HD void s_mid_stage(float2* x, int stride, int nb, int tid, int nthr, const float* premul) {
    constexpr int bps = N / R;
    const int total = nb * bps;
    for (int idx = tid; idx < total; idx += nthr) {
        const int f = idx / bps;
        const int blk = idx - f * bps;
        float2* p = x + f * stride + LY::off(blk * R);
        float2 v[R];
#pragma unroll
        for (int m = 0; m < R; ++m) v[m] = p[m];
        Dft<R>::run(v);
        const float* pm = premul + blk * R + (PREMODE == 2 ? f * N : 0);
#pragma unroll
        for (int q = 0; q < R; ++q) {
            const float w = PREMODE == 2 ? pm[q] : PB_LDG(pm + q);
            v[q] = PREMODE == 2 ? make_float2(w * v[q].y, w * v[q].x) : make_float2(w * v[q].x, -w * v[q].y);
        }
        Dft<R>::run(v);
#pragma unroll
        for (int m = 0; m < R; ++m) p[m] = v[m];
    }
}

// ---- fused first / last stages (see fft2_dif_first / fft2_dit_last in fft2.cuh) ---------------------
// The first DIF stage (sub-length N) reads every sample exactly once and the last DIT stage writes every
// sample exactly once, at x[j + m M] with j running over consecutive threads: a functor feeds / drains
// them straight from / to global memory (coalesced over j), so a kernel needs no staging pass through
// shared memory and no separate output pass.
//     Src::load<R, M>(int f, int j, float2 (&v)[R])         v[m] = sample j + m M of sequence f
//     Dst::store<R, M>(int f, int j, const float2 (&v)[R])  v[m] = result j + m M of sequence f
template <int R, int M, int N, class Src, class LY = LayoutFlat>
PB_HD void s_dif_first(float2* x, int stride, int nb, const float2* __restrict__ stw, int tid, int nthr, Src& src) {
    static_assert(R * M == N, "first stage works on the whole sequence");
    const int total = nb * M;
    for (int idx = tid; idx < total; idx += nthr) {
        const int f = idx / M;
        const int j = idx - f * M;
        float2 v[R];
        src.template load<R, M>(f, j, v);
        Dft<R>::run(v);
        float2* p = x + f * stride + j;
        p[0] = v[0];
#pragma unroll
        for (int q = 1; q < R; ++q) p[q * (M + LY::PAD)] = (M == 1) ? v[q] : c_mul(v[q], PB_LDG(stw + (q - 1) * M + j));
    }
}

template <int R, int M, int N, class Dst, class LY = LayoutFlat>
PB_HD void s_dit_last(const float2* x, int stride, int nb, const float2* __restrict__ stw, int tid, int nthr, Dst& dst) {
    static_assert(R * M == N, "last stage works on the whole sequence");
    const int total = nb * M;
    for (int idx = tid; idx < total; idx += nthr) {
        const int f = idx / M;
        const int j = idx - f * M;
        const float2* p = x + f * stride + j;
        float2 v[R];
        v[0] = p[0];
#pragma unroll
        for (int q = 1; q < R; ++q) v[q] = (M == 1) ? p[q] : c_mul(p[q * (M + LY::PAD)], PB_LDG(stw + (q - 1) * M + j));
        Dft<R>::run(v);
        dst.template store<R, M>(f, j, v);
    }
}

// ---- drivers (compile-time recursion over the stages) ---------------------------------------
template <class P, int S, int COUNT, bool WARP, class LY = LayoutFlat>
struct SDifRun {     // DIF stages S, S+1, ... (COUNT of them)
    static PB_HD void run(float2* x, int stride, int nb, const float2* __restrict__ tw, int tid, int nthr) {
        constexpr int R = P::R(S), L = P::L(S);
        s_dif_stage<R, L / R, P::n, LY>(x, stride, nb, tw + P::tw_off(S), tid, nthr);
        fft2_sync<WARP>();
        SDifRun<P, S + 1, COUNT - 1, WARP, LY>::run(x, stride, nb, tw, tid, nthr);
    }
};
template <class P, int S, bool WARP, class LY>
struct SDifRun<P, S, 0, WARP, LY> {
    static PB_HD void run(float2*, int, int, const float2* __restrict__, int, int) {}
};

template <class P, int S, int COUNT, bool WARP, class LY = LayoutFlat>
struct SDitRun {     // DIT stages S, S-1, ... (COUNT of them)
    static PB_HD void run(float2* x, int stride, int nb, const float2* __restrict__ tw, int tid, int nthr) {
        constexpr int R = P::R(S), L = P::L(S);
        s_dit_stage<R, L / R, P::n, LY>(x, stride, nb, tw + P::tw_off(S), tid, nthr);
        fft2_sync<WARP>();
        SDitRun<P, S - 1, COUNT - 1, WARP, LY>::run(x, stride, nb, tw, tid, nthr);
    }
};
template <class P, int S, bool WARP, class LY>
struct SDitRun<P, S, 0, WARP, LY> {
    static PB_HD void run(float2*, int, int, const float2* __restrict__, int, int) {}
};

// forward DIF (all stages)
template <class P, bool WARP = false>
PB_HD void s_forward_dif(float2* x, int stride, int nb, const float2* __restrict__ tw, int tid, int nthr) {
    SDifRun<P, 0, P::ns, WARP>::run(x, stride, nb, tw, tid, nthr);
}
// inverse-direction DIT (all stages, no multiplier)
template <class P, bool WARP = false>
PB_HD void s_forward_dit(float2* x, int stride, int nb, const float2* __restrict__ tw, int tid, int nthr) {
    SDitRun<P, P::ns - 1, P::ns, WARP>::run(x, stride, nb, tw, tid, nthr);
}
// forward, pointwise multiplier, inverse direction, innermost stages fused (fft2_forward_mul_inverse)
template <class P, int PREMODE, bool WARP = false>
PB_HD void s_forward_mul_inverse(float2* x, int stride, int nb, const float2* __restrict__ tw, int tid, int nthr,
                                 const float* premul) {
    SDifRun<P, 0, P::ns - 1, WARP>::run(x, stride, nb, tw, tid, nthr);
    s_mid_stage<P::R(P::ns - 1), P::n, PREMODE>(x, stride, nb, tid, nthr, premul);
    fft2_sync<WARP>();
    SDitRun<P, (P::ns >= 2 ? P::ns - 2 : 0), P::ns - 1, WARP>::run(x, stride, nb, tw, tid, nthr);
}

// The lengths with a compile-time path (radix order = what make_fft2_plan returns for them).
using PlanW1920 = StaticPlan<1920, 8, 16, 15>;
using PlanH1080 = StaticPlan<1080, 8, 15, 9>;
using PlanX2016 = StaticPlan<2016, 16, 14, 9>;
using PlanY1152 = StaticPlan<1152, 8, 16, 9>;
using PlanW3840 = StaticPlan<3840, 16, 16, 15>;
using PlanH2160 = StaticPlan<2160, 15, 16, 9>;
using PlanX4000 = StaticPlan<4000, 10, 10, 8, 5>;       // what the run-time planner (radices <= 16) gives: four stages
using PlanX4000b = StaticPlan<4000, 20, 20, 10>;        // three stages with 20-point butterflies (second-generation row passes)
using PlanY2304 = StaticPlan<2304, 16, 16, 9>;
// the torus of the single 12000 x 9000 image (BASELINE C4): the run-time planner's orders
using PlanX12096 = StaticPlan<12096, 16, 9, 12, 7>;
using PlanY9216 = StaticPlan<9216, 16, 4, 16, 9>;
// three stages for the second-generation row passes of that torus: a 36-point first stage (128 registers per thread, two
// CTAs per SM are all the 95 KB sequences allow anyway), 14-point mirror units last
using PlanX12096b = StaticPlan<12096, 36, 24, 14>;

}  // namespace pb
