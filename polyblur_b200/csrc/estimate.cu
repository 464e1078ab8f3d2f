// Blur estimation kernels: gray + min/max + column spectral derivative (k_cols), row spectral
// derivative + 7 directional maxima (k_rows), and the per-image parameter kernel (k_params:
// Keys interpolation -> argmin -> sigma, rho -> 25x25 taps).
//
// Reference being replaced: blur_estimation.gaussian_blur_estimation
// (polyblur/blur_estimation.py:18-79) and filters.fourier_gradients (polyblur/filters.py:159-186).
//
// HBM traffic per pixel of one estimate (C = 3): k_cols reads 12 B (the iterate) and writes
// 8 B (gray, gy); k_rows reads 8 B (gray, gy).  gray/gy are scratch that stays L2 resident
// when the batch is processed a few images at a time (api.cu chunks the launches).
#include "fft.cuh"
#include "kernels.cuh"

namespace pb {

// cos / sin of torch.linspace(0, pi, 7) exactly as torch (float32) evaluates them
// (blur_estimation.py:127-129); bit patterns recorded from torch 2.11 CPU.
__constant__ float c_cos7[7] = {0x1.000000p+0f, 0x1.bb67aep-1f, 0x1.fffffep-2f, -0x1.777a5cp-25f,
                                -0x1.000002p-1f, -0x1.bb67aep-1f, -0x1.000000p+0f};
__constant__ float c_sin7[7] = {0x0.0p+0f, 0x1.000000p-1f, 0x1.bb67aep-1f, 0x1.000000p+0f,
                                0x1.bb67aep-1f, 0x1.000002p-1f, -0x1.777a5cp-24f};
// (30,7) Keys interpolation weights travel as a kernel argument (840 bytes): no host->device copy in
// the call, so the whole pb_polyblur_f32 enqueue can be captured into a CUDA graph.
struct KeysW {
    float w[30 * 7];
};

__global__ void k_twiddles(float2* __restrict__ tw, int n) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) {
        double s, c;
        sincospi(-2.0 * (double)k / (double)n, &s, &c);
        tw[k] = make_float2((float)c, (float)s);
    }
}

__global__ void k_init_stats(unsigned* __restrict__ stats, int B) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < B * PB_STATS_STRIDE) {
        int j = i % PB_STATS_STRIDE;
        unsigned v = 0u;                       // mags: +0.0f
        if (j == 0) v = 0xffffffffu;           // running min (ordered encoding)
        stats[i] = v;                          // j == 1: running max starts at 0 = below everything
    }
}

// ---------------------------------------------------------------------------------------------
// k_cols: one CTA owns a strip of 2*nb adjacent columns of one image (all H rows).
//   EST : img is (B,C,H,W); forms the channel mean g (blur_estimation.py:36-37), stores it to
//         `gray`, tracks min/max (blur_estimation.py:106-108) and writes d g / d y to `gy`.
//   GRAD: img is a stack of planes (B*C,H,W); writes d plane / d y to `gy` only.
// Two real columns share one complex transform (real/imaginary parts).
// ---------------------------------------------------------------------------------------------
template <bool EST>
__global__ void __launch_bounds__(PB_FFT_THREADS)
k_cols(const float* __restrict__ img, float* __restrict__ gray, float* __restrict__ gy,
       unsigned* __restrict__ stats, int C, int H, int W, int nb, FftPlan plan,
       const float2* __restrict__ tw) {
    extern __shared__ float2 smem[];
    float2* a = smem;
    float2* b = smem + (size_t)nb * H;
    const int cw = 2 * nb;
    const int x0 = blockIdx.x * cw;
    const int im = blockIdx.y;
    const size_t plane = (size_t)H * W;
    const float* src = img + (size_t)im * (EST ? C : 1) * plane;
    float* af = reinterpret_cast<float*>(a);

    float lmin = INFINITY, lmax = -INFINITY;
    const float invC = (float)C;
    for (int idx = threadIdx.x; idx < H * cw; idx += blockDim.x) {
        const int y = idx / cw;
        const int xx = idx - y * cw;
        const int x = x0 + xx;
        float g = 0.0f;
        if (x < W) {
            const float* p = src + (size_t)y * W + x;
            if (EST) {
                g = p[0];
                for (int c = 1; c < C; ++c) g = __fadd_rn(g, p[(size_t)c * plane]);
                if (C > 1) g = __fdiv_rn(g, invC);
                gray[(size_t)im * plane + (size_t)y * W + x] = g;
                lmin = fminf(lmin, g);
                lmax = fmaxf(lmax, g);
            } else {
                g = p[0];
            }
        }
        af[((size_t)(xx >> 1) * H + y) * 2 + (xx & 1)] = g;
    }
    __syncthreads();
    fft_derivative(a, b, nb, plan, tw);
    const float inv = 1.0f / (float)H;
    for (int idx = threadIdx.x; idx < H * cw; idx += blockDim.x) {
        const int y = idx / cw;
        const int xx = idx - y * cw;
        const int x = x0 + xx;
        if (x < W) {
            float2 z = a[(size_t)(xx >> 1) * H + y];
            float v = (xx & 1) ? -z.y : z.x;
            gy[(size_t)im * plane + (size_t)y * W + x] = v * inv;
        }
    }
    if (EST) {
        lmin = warp_min(lmin);
        lmax = warp_max(lmax);
        if ((threadIdx.x & 31) == 0) {
            atomicMin(&stats[im * PB_STATS_STRIDE + 0], f2ord(lmin));
            atomicMax(&stats[im * PB_STATS_STRIDE + 1], f2ord(lmax));
        }
    }
}

// ---------------------------------------------------------------------------------------------
// k_rows: one CTA owns 2*nb adjacent rows of one image / plane.
//   EST : reads gray rows, computes d g / d x, reads gy and reduces
//         max |cos(phi_j) gx - sin(phi_j) gy| for the 7 angles (blur_estimation.py:122-134)
//         into stats (un-normalised: k_params divides by max - min, which is what
//         normalising before the derivative does for q = 0).
//   GRAD: reads plane rows, writes d plane / d x to gx.
// ---------------------------------------------------------------------------------------------
template <bool EST>
__global__ void __launch_bounds__(PB_FFT_THREADS)
k_rows(const float* __restrict__ plane_in, const float* __restrict__ gy, float* __restrict__ gx,
       unsigned* __restrict__ stats, int H, int W, int nb, FftPlan plan,
       const float2* __restrict__ tw, int discard_saturation) {
    extern __shared__ float2 smem[];
    __shared__ float red[PB_FFT_THREADS / 32][8];
    float2* a = smem;
    float2* b = smem + (size_t)nb * W;
    const int rh = 2 * nb;
    const int y0 = blockIdx.x * rh;
    const int im = blockIdx.y;
    const size_t plane = (size_t)H * W;
    const float* src = plane_in + (size_t)im * plane;
    float* af = reinterpret_cast<float*>(a);

    for (int idx = threadIdx.x; idx < rh * W; idx += blockDim.x) {
        const int rr = idx / W;
        const int x = idx - rr * W;
        const int y = y0 + rr;
        float g = (y < H) ? src[(size_t)y * W + x] : 0.0f;
        af[((size_t)(rr >> 1) * W + x) * 2 + (rr & 1)] = g;
    }
    __syncthreads();
    fft_derivative(a, b, nb, plan, tw);
    const float inv = 1.0f / (float)W;

    if (!EST) {
        for (int idx = threadIdx.x; idx < rh * W; idx += blockDim.x) {
            const int rr = idx / W;
            const int x = idx - rr * W;
            const int y = y0 + rr;
            if (y < H) {
                float2 z = a[(size_t)(rr >> 1) * W + x];
                gx[(size_t)im * plane + (size_t)y * W + x] = ((rr & 1) ? -z.y : z.x) * inv;
            }
        }
        return;
    }

    float m[7];
#pragma unroll
    for (int j = 0; j < 7; ++j) m[j] = 0.0f;
    for (int idx = threadIdx.x; idx < rh * W; idx += blockDim.x) {
        const int rr = idx / W;
        const int x = idx - rr * W;
        const int y = y0 + rr;
        if (y >= H) continue;
        const size_t off = (size_t)im * plane + (size_t)y * W + x;
        // get_saturation_mask (blur_estimation.py:83-88): un-normalised gray > 0.99
        if (discard_saturation && plane_in[off] > 0.99f) continue;
        float2 z = a[(size_t)(rr >> 1) * W + x];
        const float gxv = ((rr & 1) ? -z.y : z.x) * inv;
        const float gyv = gy[off];
#pragma unroll
        for (int j = 0; j < 7; ++j) {
            float v = __fsub_rn(__fmul_rn(c_cos7[j], gxv), __fmul_rn(c_sin7[j], gyv));
            m[j] = fmaxf(m[j], fabsf(v));
        }
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int j = 0; j < 7; ++j) {
        float v = warp_max(m[j]);
        if (lane == 0) red[warp][j] = v;
    }
    __syncthreads();
    if (threadIdx.x < 7) {
        float v = 0.0f;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) v = fmaxf(v, red[w][threadIdx.x]);
        atomicMax(&stats[im * PB_STATS_STRIDE + 2 + threadIdx.x], __float_as_uint(v));
    }
}

// ---------------------------------------------------------------------------------------------
// k_params: one CTA per image.  Thread 0 follows blur_estimation.py:138-185 (Keys
// interpolation 7 -> 30, first argmin, orthogonal lookup, affine model, clamp, sqrt), then
// the CTA builds the normalised 25x25 taps (blur_estimation.py:189-232), thresholds them for
// the spatial engine and records per-row tap ranges.
//   mode 0: parameters come from stats (estimation);  mode 1: theta/sigma/rho given
//   (pb_make_kernel_f32);  mode 2: explicit taps given in `kin` (pb_deconv_f32).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
k_params(const unsigned* __restrict__ stats, ImgKernel* __restrict__ kern, float* __restrict__ est,
         const float* __restrict__ th_in, const float* __restrict__ sg_in, const float* __restrict__ rh_in,
         const float* __restrict__ kin, float* __restrict__ kout, int mode, int ksize, float cc, float bb,
         float tap_thr, int engine_req, int fft_radius_min, int* __restrict__ cls_buf, int list_stride,
         KeysW keys) {
    __shared__ float s_par[3];
    __shared__ float s_red[4];
    __shared__ float s_k[PB_KS2];
    __shared__ int s_cnt, s_rad, s_rx, s_ry;
    const int im = blockIdx.x;
    const int tid = threadIdx.x;
    ImgKernel* K = kern + im;
    if (tid == 0) {
        s_cnt = 0;
        s_rad = 0;
        s_rx = 0;
        s_ry = 0;
        float theta = 0.f, sigma = 0.f, rho = 0.f;
        if (mode == 0) {
            const unsigned* st = stats + im * PB_STATS_STRIDE;
            const float mn = ord2f(st[0]), mx = ord2f(st[1]);
            const float range = __fsub_rn(mx, mn);
            float mags[7], interp[30];
            for (int j = 0; j < 7; ++j) mags[j] = __fdiv_rn(__uint_as_float(st[2 + j]), range);
            int imin = 0;
            for (int i = 0; i < 30; ++i) {
                float acc = 0.0f;
                for (int j = 0; j < 7; ++j) acc = fmaf(keys.w[i * 7 + j], mags[j], acc);
                interp[i] = acc;
                if (acc < interp[imin]) imin = i;          // first minimum (torch.argmin)
            }
            const int deg = 6 * imin;
            const int iort = ((deg + 90) % 180) / 6;
            const float mnrm = interp[imin], mort = interp[iort];
            float v = __fsub_rn(__fdiv_rn(cc, __fadd_rn(__fmul_rn(mnrm, mnrm), 1e-8f)), bb);
            sigma = __fsqrt_rn(fminf(fmaxf(v, 0.09f), 16.0f));
            v = __fsub_rn(__fdiv_rn(cc, __fadd_rn(__fmul_rn(mort, mort), 1e-8f)), bb);
            rho = __fsqrt_rn(fminf(fmaxf(v, 0.09f), 16.0f));
            theta = __fdiv_rn(__fmul_rn((float)deg, 3.14159274101257324f), 180.0f);
            if (est) {
                float* e = est + (size_t)im * PB_EST_STRIDE;
                for (int j = 0; j < 7; ++j) e[j] = mags[j];
                e[7] = (float)deg;
                e[8] = sigma;
                e[9] = rho;
                e[10] = mnrm;
                e[11] = mort;
            }
        } else if (mode == 1) {
            theta = th_in[im];
            sigma = sg_in[im];
            rho = rh_in[im];
        }
        s_par[0] = theta;
        s_par[1] = sigma;
        s_par[2] = rho;
    }
    __syncthreads();
    const int half = (ksize - 1) / 2;
    float local = 0.0f;
    if (mode != 2) {
        const float theta = s_par[0], sigma = s_par[1], rho = s_par[2];
        // compute_gaussian_filter_parameters (blur_estimation.py:189-208), fp32 op by op;
        // cos/sin/exp are evaluated in fp64 and rounded (the correctly rounded fp32 value).
        const float th = -theta;
        const float c = (float)cos((double)th), s = (float)sin((double)th);
        const float c2 = __fmul_rn(c, c), s2 = __fmul_rn(s, s), sc = __fmul_rn(s, c);
        const float il1 = __fdiv_rn(1.0f, __fmul_rn(sigma, sigma));
        const float il2 = __fdiv_rn(1.0f, __fmul_rn(rho, rho));
        const float a00 = __fadd_rn(__fmul_rn(c2, il1), __fmul_rn(s2, il2));
        const float a01 = __fmul_rn(sc, __fsub_rn(il1, il2));
        const float a11 = __fadd_rn(__fmul_rn(c2, il2), __fmul_rn(s2, il1));
        for (int i = tid; i < PB_KS2; i += blockDim.x) {
            const int dy = i / PB_KS - PB_PAD, dx = i % PB_KS - PB_PAD;
            float kv = 0.0f;
            if (abs(dy) <= half && abs(dx) <= half) {
                const float X = (float)dx, Y = (float)dy;
                const float u = __fadd_rn(__fmul_rn(X, a00), __fmul_rn(Y, a01));
                const float v = __fadd_rn(__fmul_rn(X, a01), __fmul_rn(Y, a11));
                const float q = __fadd_rn(__fmul_rn(u, X), __fmul_rn(v, Y));
                kv = (float)exp((double)__fmul_rn(-0.5f, q));
            }
            s_k[i] = kv;
            local += kv;
        }
        local = warp_sum(local);
        if ((tid & 31) == 0) s_red[tid >> 5] = local;
        __syncthreads();
        const float total = (s_red[0] + s_red[1]) + (s_red[2] + s_red[3]);
        for (int i = tid; i < PB_KS2; i += blockDim.x) s_k[i] = __fdiv_rn(s_k[i], total);
    } else {
        // explicit taps: (ksize x ksize) embedded in the centre of the 25x25 grid, rotated by 180 degrees:
        // the engines (and k_et_pass) evaluate sum_d k[d] x[p + d], the reference's p2o / fft2 product
        // (filters.py:255-273, deblurring.py:141-169) is the convolution sum_d K[d] x[p - d], so k[d] = K[-d].
        // (A Gaussian from mode 0 / 1 is point-symmetric bit for bit, so nothing changes for it.)
        for (int i = tid; i < PB_KS2; i += blockDim.x) {
            const int dy = i / PB_KS - PB_PAD, dx = i % PB_KS - PB_PAD;
            float kv = 0.0f;
            if (abs(dy) <= half && abs(dx) <= half)
                kv = kin[(size_t)im * ksize * ksize + (half - dy) * ksize + (half - dx)];
            s_k[i] = kv;
        }
    }
    __syncthreads();
    // largest magnitude tap (the centre for a Gaussian)
    float kmax = 0.0f;
    for (int i = tid; i < PB_KS2; i += blockDim.x) kmax = fmaxf(kmax, fabsf(s_k[i]));
    kmax = warp_max(kmax);
    __syncthreads();
    if ((tid & 31) == 0) s_red[tid >> 5] = kmax;
    __syncthreads();
    kmax = fmaxf(fmaxf(s_red[0], s_red[1]), fmaxf(s_red[2], s_red[3]));
    const float thr = tap_thr * kmax;
    for (int i = tid; i < PB_KS2; i += blockDim.x) K->k[i] = s_k[i];
    // The FFT engine builds a real kernel spectrum from the rows dy >= 0 (deconv_fft.cu), which is only the
    // spectrum of a point-symmetric kernel: explicit taps with K[-d] != K[d] (motion blur, shifted Gaussians)
    // are kept on the spatial engines, whatever engine was asked for.
    int asym = 0;
    if (mode == 2) {
        const float tol = kmax * 2.4e-7f;
        for (int i = tid; i < PB_KS2 / 2; i += blockDim.x) asym |= fabsf(s_k[i] - s_k[PB_KS2 - 1 - i]) > tol;
    }
    asym = __syncthreads_or(asym);
    if (kout) {
        for (int i = tid; i < ksize * ksize; i += blockDim.x) {
            const int yy = i / ksize, xx = i % ksize;
            kout[(size_t)im * ksize * ksize + i] = s_k[(yy - half + PB_PAD) * PB_KS + (xx - half + PB_PAD)];
        }
    }
    if (tid < PB_KS) {
        int lo = PB_KS, hi = -1, cnt = 0, rad = 0;
        for (int x = 0; x < PB_KS; ++x) {
            const float v = fabsf(s_k[tid * PB_KS + x]);
            if (v > 0.0f && v >= thr) {
                if (lo == PB_KS) lo = x;
                hi = x;
                ++cnt;
                rad = max(rad, max(abs(x - PB_PAD), abs(tid - PB_PAD)));
            }
        }
        // the stencil applies every tap in [lo, hi]; interior gaps cannot occur for a Gaussian
        K->lo[tid] = lo;
        K->hi[tid] = hi;
        atomicAdd(&s_cnt, hi >= lo ? hi - lo + 1 : 0);
        atomicMax(&s_rad, rad);
        if (hi >= lo) {
            atomicMax(&s_ry, abs(tid - PB_PAD));
            atomicMax(&s_rx, max(abs(lo - PB_PAD), abs(hi - PB_PAD)));
        }
        (void)cnt;
    }
    __syncthreads();
    if (tid == 0) {
        K->radius = s_rad;
        K->ntaps = s_cnt;
        K->ksize = ksize;
        // engine class: the FFT engine when asked for (or, on AUTO, for wide kernels), else the
        // narrowest spatial engine that holds every kept tap
        int cls;
        if (engine_req == PB_ENGINE_FFT && !asym) cls = PB_CLS_FFT;
        else if (s_rx <= 1 && s_ry <= 1) cls = PB_CLS_N11;
        else if (s_rx <= 2 && s_ry <= 2) cls = PB_CLS_N22;
        else if (engine_req == PB_ENGINE_AUTO && s_rad >= fft_radius_min && !asym) cls = PB_CLS_FFT;
        else cls = (s_rad <= 4) ? PB_CLS_TILED4 : PB_CLS_TILED;
        K->engine = (cls == PB_CLS_FFT) ? PB_ENGINE_FFT : PB_ENGINE_SPATIAL;
        K->rx = s_rx;
        K->ry = s_ry;
        K->cls = cls;
        if (cls_buf) {
            const int slot = atomicAdd(&cls_buf[cls], 1);
            cls_buf[PB_CLS_COUNT_STRIDE + cls * list_stride + slot] = im;
        }
        K->theta = s_par[0];
        K->sigma = s_par[1];
        K->rho = s_par[2];
        K->pad_[0] = K->pad_[1] = 0;
    }
}

// ---- host side ------------------------------------------------------------------------------

static void host_keys_weights(float* w) {
    // blur_estimation.py:138-148 with x = long(linspace(0,180,7))/30, x_new = long(arange(0,180,6))/30
    volatile float tmp;
    for (int i = 0; i < 30; ++i) {
        float xn = (float)(6 * i) / 30.0f;
        float row[7];
        float sum = 0.0f;
        for (int j = 0; j < 7; ++j) {
            float x = (float)(30 * j) / 30.0f;
            float d = fabsf(xn - x);
            float v = 0.0f;
            if (d < 1.0f) {
                tmp = 1.5f * d; tmp = tmp - 2.5f; tmp = tmp * d; tmp = tmp * d; tmp = tmp + 1.0f;
                v = tmp;
            } else if (d < 2.0f) {
                tmp = -0.5f * d; tmp = tmp + 2.5f; tmp = tmp * d; tmp = tmp - 4.0f; tmp = tmp * d; tmp = tmp + 2.0f;
                v = tmp;
            }
            row[j] = v;
            tmp = sum + v;
            sum = tmp;
        }
        tmp = sum + 1e-5f;
        float den = tmp;
        for (int j = 0; j < 7; ++j) {
            tmp = row[j] / den;
            w[i * 7 + j] = tmp;
        }
    }
}

void keys_weights_host(float* out210) { host_keys_weights(out210); }

int upload_constants(cudaStream_t) { return PB_OK; }   // nothing to upload any more (see KeysW)

int launch_twiddles(float2* tw, int n, cudaStream_t stream) {
    ProfScope prof(PROF_SETUP, stream);
    k_twiddles<<<(n + 255) / 256, 256, 0, stream>>>(tw, n);
    PB_LAUNCH_CHECK("k_twiddles");
    return PB_OK;
}

int launch_init_stats(unsigned* stats, int B, cudaStream_t stream) {
    int n = B * PB_STATS_STRIDE;
    ProfScope prof(PROF_SETUP, stream);
    k_init_stats<<<(n + 255) / 256, 256, 0, stream>>>(stats, B);
    PB_LAUNCH_CHECK("k_init_stats");
    return PB_OK;
}

// number of packed transforms per CTA for a length-n FFT: as many as fit ~96 KB, at most 8
int fft_batch_for(int n, int max_nb) {
    size_t per = (size_t)n * sizeof(float2) * 2;
    int nb = (int)((96 * 1024) / per);
    if (nb < 1) nb = 1;
    if (nb > max_nb) nb = max_nb;
    return nb;
}

template <typename KernelT>
static int set_smem(KernelT kern, size_t bytes) {
    if (bytes > PB_SMEM_MAX - 1024) {
        set_error("FFT length needs %zu bytes of shared memory (> %d): image side too large for the on-chip FFT",
                  bytes, PB_SMEM_MAX - 1024);
        return PB_ERR_UNSUPPORTED;
    }
    if (bytes > 48 * 1024)
        PB_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    return PB_OK;
}

int launch_cols(bool est, const float* img, float* gray, float* gy, unsigned* stats, int nimg, int C,
                int H, int W, const FftPlan& planH, const float2* twH, cudaStream_t stream) {
    const int nb = fft_batch_for(H, 8);
    const size_t smem = (size_t)nb * H * sizeof(float2) * 2;
    dim3 grid((W + 2 * nb - 1) / (2 * nb), nimg);
    ProfScope prof(PROF_COLS, stream);
    if (est) {
        int rc = set_smem(k_cols<true>, smem);
        if (rc) return rc;
        k_cols<true><<<grid, PB_FFT_THREADS, smem, stream>>>(img, gray, gy, stats, C, H, W, nb, planH, twH);
    } else {
        int rc = set_smem(k_cols<false>, smem);
        if (rc) return rc;
        k_cols<false><<<grid, PB_FFT_THREADS, smem, stream>>>(img, gray, gy, stats, C, H, W, nb, planH, twH);
    }
    PB_LAUNCH_CHECK("k_cols");
    return PB_OK;
}

int launch_rows(bool est, const float* plane_in, const float* gy, float* gx, unsigned* stats, int nimg,
                int H, int W, const FftPlan& planW, const float2* twW, int discard_saturation,
                cudaStream_t stream) {
    const int nb = fft_batch_for(W, 4);
    const size_t smem = (size_t)nb * W * sizeof(float2) * 2;
    dim3 grid((H + 2 * nb - 1) / (2 * nb), nimg);
    ProfScope prof(PROF_ROWS, stream);
    if (est) {
        int rc = set_smem(k_rows<true>, smem);
        if (rc) return rc;
        k_rows<true><<<grid, PB_FFT_THREADS, smem, stream>>>(plane_in, gy, gx, stats, H, W, nb, planW, twW,
                                                             discard_saturation);
    } else {
        int rc = set_smem(k_rows<false>, smem);
        if (rc) return rc;
        k_rows<false><<<grid, PB_FFT_THREADS, smem, stream>>>(plane_in, gy, gx, stats, H, W, nb, planW, twW,
                                                              discard_saturation);
    }
    PB_LAUNCH_CHECK("k_rows");
    return PB_OK;
}

int launch_params(const unsigned* stats, ImgKernel* kern, float* est, const float* th, const float* sg,
                  const float* rh, const float* kin, float* kout, int mode, int B, int ksize, float cc,
                  float bb, float tap_thr, int engine_req, int fft_radius_min, int* cls, cudaStream_t stream) {
    ProfScope prof(PROF_PARAMS, stream);
    if (cls) PB_CUDA_TRY(cudaMemsetAsync(cls, 0, PB_CLS_COUNT_STRIDE * sizeof(int), stream));
    static KeysW keys;
    static bool keys_ready = false;
    if (!keys_ready) {
        host_keys_weights(keys.w);
        keys_ready = true;
    }
    k_params<<<B, 128, 0, stream>>>(stats, kern, est, th, sg, rh, kin, kout, mode, ksize, cc, bb, tap_thr,
                                    engine_req, fft_radius_min, cls, B, keys);
    PB_LAUNCH_CHECK("k_params");
    return PB_OK;
}

}  // namespace pb
