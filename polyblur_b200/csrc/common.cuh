// Shared helpers of the polyblur_b200 CUDA kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/polyblur_b200.h"

#define PB_PAD 12                 // kernel.shape[-1]//2 for the 25x25 support (utils.py:48-61)
#define PB_KS 25
#define PB_KS2 625
#define PB_MAX_STAGES 32
#define PB_SMEM_MAX 232448        // 227 KB opt-in dynamic shared memory per CTA on sm_100
#define PB_NUM_SMS 148

namespace pb {

// ---- error plumbing (api.cu owns the storage) -------------------------------------------
void set_error(const char* fmt, ...);
int check_cuda(cudaError_t e, const char* what);
#define PB_CUDA_TRY(expr)                                        \
    do {                                                         \
        int _rc = ::pb::check_cuda((expr), #expr);               \
        if (_rc != PB_OK) return _rc;                            \
    } while (0)
#define PB_LAUNCH_CHECK(name) PB_CUDA_TRY(cudaPeekAtLastError())

// ---- FFT plan ---------------------------------------------------------------------------
struct FftPlan {
    int n;
    int ns;
    int radix[PB_MAX_STAGES];
};
int make_fft_plan(int n, FftPlan* plan);   // host; returns PB_OK / PB_ERR_ARG

// ---- per-image blur parameters, written by k_params, read by the deconvolution ----------
//   One record per image, all 4-byte fields (see estimate.cu / deconv.cu).
struct ImgKernel {
    float k[PB_KS2];          // normalised 25x25 taps, row-major [dy+12][dx+12]
    int   lo[PB_KS];          // per row dy: first kept dx (+12), 25 when the row is empty
    int   hi[PB_KS];          // per row dy: last kept dx (+12), -1 when the row is empty
    int   radius;             // max(|dy|,|dx|) over kept taps
    int   ntaps;              // number of kept taps
    int   ksize;              // odd support actually used (<= 25); taps outside are zero
    int   engine;             // PB_ENGINE_SPATIAL / PB_ENGINE_FFT chosen for this image
    float theta, sigma, rho;  // radians
    int   rx, ry;             // max |dx|, max |dy| over kept taps
    int   cls;                // deconvolution engine class (PB_CLS_*), chosen on the device
    int   pad_[2];
};

// ---- ordered-int encoding so that atomicMin/atomicMax work on any float -----------------
__device__ __forceinline__ unsigned f2ord(float f) {
    unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(unsigned u) {
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// positive modulo
__device__ __forceinline__ int pmod(int a, int n) {
    int r = a % n;
    return r < 0 ? r + n : r;
}

// Gather map of the replicate-padded torus (utils.py:48-53 + the circular
// FFT product of deblurring.py:141-169): padded coordinate (any integer) -> source index.
__device__ __forceinline__ int torus_src(int padded, int n, int pad) {
    int m = pmod(padded, n + 2 * pad) - pad;
    return min(max(m, 0), n - 1);
}

// Where a deconvolution engine reads its input from.  Default: the (H, W) image itself, replicate
// padded by the kernel half-size on the fly (pad < 0 = take ksize / 2 of the image's kernel record).
// Edgetaper mode: an explicitly padded (Hin, Win) = (H + 2P, W + 2P) plane that already holds the
// tapered padded image -- the torus is then pure wrap-around (pad = 0) and image pixel (y, x)
// lives at (y + off, x + off).
struct SrcGeom {
    int Hin, Win;
    int off;
    int pad;
    int clamp_out;     // clamp the result to [0,1] (off when halo masking post-processes it)
};
__device__ __forceinline__ int geom_src(int coord, int n_in, int off, int pad) {
    return torus_src(coord + off + pad, n_in, pad);
}

// One-pass image / spectrum traffic of the FFT engine's row passes: with PB_STREAM_HINTS these loads and stores
// use the streaming (evict-first) cache operators.  Measured +1.5 % (slower) there, so off by default; the
// column kernel of the estimator, whose shared memory leaves only 28 KB of L1, uses __ldcs directly (-3.6 %).
#ifdef PB_STREAM_HINTS
#define PB_LD_STREAM(p) __ldcs(p)
#define PB_ST_STREAM(p, v) __stcs((p), (v))
#else
#define PB_LD_STREAM(p) __ldg(p)
#define PB_ST_STREAM(p, v) (*(p) = (v))
#endif

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

}  // namespace pb
