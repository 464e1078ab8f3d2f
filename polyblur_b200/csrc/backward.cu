// Backward pass through the blur estimator: what torch.autograd computes over
// gaussian_blur_estimation (polyblur/blur_estimation.py:18-79) and over the kernel argument of
// inverse_filtering_rank3 (polyblur/deblurring.py:211-239), restated as explicit kernels.
//
// Forward trace (k_bw_gray .. k_bw_finish): the channel mean g, its min / max and how many pixels
// attain them (torch.amin / amax spread their gradient evenly over ties), the normalised image, its
// spectral gradients (the ordinary gradient kernels) and, for each of the 7 angles, the pixel where
// |cos gx - sin gy| is largest together with the sign there.
//
// Gradient of the loss with respect to the 25 x 25 kernel (k_bw_kernel_grad): with Q(K) = dP/dK the
// derivative of the deconvolution polynomial, V = Q(K) (*) pad(x) and y~ the masked upstream gradient,
//     K~[d] = sum_c sum_p y~[c, p] V[c, p + pad - d]       (p over the image, never wraps).
// The chain K~ -> sigma~, rho~ -> the 7 maxima is a handful of scalars per image and stays on the host
// side of the C ABI (polyblur_b200/autograd.py); it comes back as m~ (B, 7).
//
// Estimator VJP (k_bw_scatter .. k_bw_norm_apply): m~ goes to the arg-max pixels with the sign and the
// angle's cos / -sin, the spectral derivative is antisymmetric (D^T = -D), so the gradient with respect
// to the normalised image is -(Dx sgx + Dy sgy) computed by the same gradient kernels, and the range
// normalisation (g - mn) / (mx - mn) is differentiated including its min / max terms.
#include "kernels.cuh"

namespace pb {

// cos / sin of torch.linspace(0, pi, 7) in float32 (same bit patterns as estimate2.cu)
__device__ __constant__ float c_bw_cos7[7] = {0x1.000000p+0f, 0x1.bb67aep-1f, 0x1.fffffep-2f, -0x1.777a5cp-25f,
                                              -0x1.000002p-1f, -0x1.bb67aep-1f, -0x1.000000p+0f};
__device__ __constant__ float c_bw_sin7[7] = {0x0.0p+0f, 0x1.000000p-1f, 0x1.bb67aep-1f, 0x1.000000p+0f,
                                              0x1.bb67aep-1f, 0x1.000002p-1f, -0x1.777a5cp-24f};

__device__ __forceinline__ float bw_gray_of(const float* __restrict__ img, size_t plane, int C, size_t o) {
    float g = img[o];
    for (int c = 1; c < C; ++c) g = __fadd_rn(g, img[(size_t)c * plane + o]);
    return C > 1 ? __fdiv_rn(g, (float)C) : g;
}

__global__ void k_bw_init(unsigned* __restrict__ stats, unsigned long long* __restrict__ keys, int B) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < B) {
        stats[4 * i] = 0xffffffffu;          // running min (ordered encoding)
        stats[4 * i + 1] = 0u;               // running max
        stats[4 * i + 2] = 0u;               // pixels that attain the min
        stats[4 * i + 3] = 0u;               // ... the max
    }
    if (i < B * 7) keys[i] = 0ull;
}

// g = channel mean (blur_estimation.py:36-37), running min / max per image
__global__ void k_bw_gray(const float* __restrict__ img, float* __restrict__ g, unsigned* __restrict__ stats, int C,
                          size_t plane) {
    const int b = blockIdx.y;
    const float* src = img + (size_t)b * C * plane;
    float lmin = INFINITY, lmax = -INFINITY;
    for (size_t o = (size_t)blockIdx.x * blockDim.x + threadIdx.x; o < plane; o += (size_t)gridDim.x * blockDim.x) {
        const float v = bw_gray_of(src, plane, C, o);
        g[(size_t)b * plane + o] = v;
        lmin = fminf(lmin, v);
        lmax = fmaxf(lmax, v);
    }
    lmin = warp_min(lmin);
    lmax = warp_max(lmax);
    if ((threadIdx.x & 31) == 0) {
        atomicMin(&stats[4 * b], f2ord(lmin));
        atomicMax(&stats[4 * b + 1], f2ord(lmax));
    }
}

// clamp_((g - mn) / (mx - mn), 0, 1)  (blur_estimation.py:92-93, 106-109)
__global__ void k_bw_normalize(const float* __restrict__ g, float* __restrict__ gn, const unsigned* __restrict__ stats,
                               size_t plane) {
    const int b = blockIdx.y;
    const float mn = ord2f(stats[4 * b]), mx = ord2f(stats[4 * b + 1]);
    const float den = __fsub_rn(mx, mn);
    for (size_t o = (size_t)blockIdx.x * blockDim.x + threadIdx.x; o < plane; o += (size_t)gridDim.x * blockDim.x) {
        const float v = __fdiv_rn(__fsub_rn(g[(size_t)b * plane + o], mn), den);
        gn[(size_t)b * plane + o] = fminf(fmaxf(v, 0.0f), 1.0f);
    }
}

// per angle: the largest |cos gx - sin gy| and where (first pixel on ties) as one 64-bit key
// discard: pixels whose un-normalised gray value is > 0.99 have their gradients zeroed (get_saturation_mask +
// compute_gradients, blur_estimation.py:83-88, 112-119)
__global__ void k_bw_dirmax(const float* __restrict__ gx, const float* __restrict__ gy, const float* __restrict__ g,
                            unsigned long long* __restrict__ keys, unsigned* __restrict__ stats, size_t plane,
                            int discard) {
    const int b = blockIdx.y;
    const float mn = ord2f(stats[4 * b]), mx = ord2f(stats[4 * b + 1]);
    unsigned nmin = 0, nmax = 0;
    unsigned long long best[7];
#pragma unroll
    for (int j = 0; j < 7; ++j) best[j] = 0ull;
    for (size_t o = (size_t)blockIdx.x * blockDim.x + threadIdx.x; o < plane; o += (size_t)gridDim.x * blockDim.x) {
        float x = gx[(size_t)b * plane + o], y = gy[(size_t)b * plane + o];
        const float gv = g[(size_t)b * plane + o];
        nmin += (gv == mn);
        nmax += (gv == mx);
        if (discard && gv > 0.99f) x = y = 0.0f;
#pragma unroll
        for (int j = 0; j < 7; ++j) {
            const float v = fabsf(__fsub_rn(__fmul_rn(c_bw_cos7[j], x), __fmul_rn(c_bw_sin7[j], y)));
            const unsigned long long k =
                ((unsigned long long)__float_as_uint(v) << 32) | (unsigned long long)(0xffffffffu - (unsigned)o);
            if (k > best[j]) best[j] = k;
        }
    }
#pragma unroll
    for (int j = 0; j < 7; ++j) {
        unsigned long long k = best[j];
        for (int s = 16; s > 0; s >>= 1) {
            const unsigned long long other = __shfl_xor_sync(0xffffffffu, k, s);
            if (other > k) k = other;
        }
        if ((threadIdx.x & 31) == 0) atomicMax(&keys[b * 7 + j], k);
    }
    nmin = __reduce_add_sync(0xffffffffu, nmin);
    nmax = __reduce_add_sync(0xffffffffu, nmax);
    if ((threadIdx.x & 31) == 0) {
        if (nmin) atomicAdd(&stats[4 * b + 2], nmin);
        if (nmax) atomicAdd(&stats[4 * b + 3], nmax);
    }
}

// one warp per image: decode the arg-max pixels and the sign there
//   trace_f[b][0..6] maxima, [7..13] sign, [14] min, [15] max, [16] #min, [17] #max;  trace_pos[b][0..6] pixel
__global__ void k_bw_finish(const float* __restrict__ gx, const float* __restrict__ gy, const float* __restrict__ g,
                            const unsigned long long* __restrict__ keys, const unsigned* __restrict__ stats,
                            float* __restrict__ trace_f, int* __restrict__ trace_pos, size_t plane, int discard) {
    const int b = blockIdx.x;
    float* tf = trace_f + (size_t)b * PB_BW_TRACE_STRIDE;
    if (threadIdx.x < 7) {
        const int j = threadIdx.x;
        const unsigned long long k = keys[b * 7 + j];
        const unsigned o = 0xffffffffu - (unsigned)(k & 0xffffffffull);
        float x = gx[(size_t)b * plane + o], y = gy[(size_t)b * plane + o];
        if (discard && g[(size_t)b * plane + o] > 0.99f) x = y = 0.0f;     // (every unmasked pixel has a zero gradient)
        const float v = __fsub_rn(__fmul_rn(c_bw_cos7[j], x), __fmul_rn(c_bw_sin7[j], y));
        tf[j] = fabsf(v);
        tf[7 + j] = (v > 0.0f) ? 1.0f : (v < 0.0f ? -1.0f : 0.0f);      // d|v|/dv, 0 at v = 0 like torch.abs
        trace_pos[b * 8 + j] = (int)o;
    }
    if (threadIdx.x == 0) {
        tf[14] = ord2f(stats[4 * b]);
        tf[15] = ord2f(stats[4 * b + 1]);
        tf[16] = (float)stats[4 * b + 2];
        tf[17] = (float)stats[4 * b + 3];
        trace_pos[b * 8 + 7] = 0;
    }
}

// K~[b][d] += sum over one 32 x 32 tile of one plane of y~[p] V[p + pad - d].
// A thread owns a 5 x 5 block of kernel offsets and 4 rows of the tile and walks along x with a sliding
// window of V per offset row in registers: 6 shared-memory loads per 25 FMAs.  The 8 row groups are summed
// by warp shuffles, one global atomic per offset and tile.
#define BW_KG_TILE 32
#define BW_KG_OB 5
__global__ void __launch_bounds__(256)
k_bw_kernel_grad(const float* __restrict__ gout, const float* __restrict__ preclamp, const float* __restrict__ V,
                 float* __restrict__ kbar, int C, int H, int W, int ks) {
    extern __shared__ float smk[];
    const int pad = ks >> 1;
    const int VT = BW_KG_TILE + 2 * pad;             // V tile edge
    const int VS = VT + 1;                           // row stride of the V tile
    constexpr int GS = BW_KG_TILE + 1;               // row stride of the gradient tile
    float* sg = smk;                                 // [32][33]
    float* sv = sg + BW_KG_TILE * GS;                // [VT][VT + 1]
    const int pl = blockIdx.z, b = pl / C;
    const int x0 = blockIdx.x * BW_KG_TILE, y0 = blockIdx.y * BW_KG_TILE;
    const int Hp = H + 2 * pad, Wp = W + 2 * pad;
    const size_t plane = (size_t)H * W;
    for (int i = threadIdx.x; i < BW_KG_TILE * BW_KG_TILE; i += blockDim.x) {
        const int ty = i / BW_KG_TILE, tx = i - ty * BW_KG_TILE;
        const int y = y0 + ty, x = x0 + tx;
        float v = 0.0f;
        if (y < H && x < W) {
            const size_t o = (size_t)pl * plane + (size_t)y * W + x;
            v = gout[o];
            if (preclamp) {
                const float u = preclamp[o];
                if (!(u >= 0.0f && u <= 1.0f)) v = 0.0f;
            }
        }
        sg[ty * GS + tx] = v;
    }
    // V tile: padded coordinates [y0, y0 + 32 + 2 pad) x [x0, ...); pixel p and kernel index (iy, ix) meet at
    // tile position (ty + 2 pad - iy, tx + 2 pad - ix)
    for (int i = threadIdx.x; i < VT * VT; i += blockDim.x) {
        const int ty = i / VT, tx = i - ty * VT;
        const int yp = y0 + ty, xp = x0 + tx;
        sv[ty * VS + tx] = (yp < Hp && xp < Wp) ? V[(size_t)pl * Hp * Wp + (size_t)yp * Wp + xp] : 0.0f;
    }
    __syncthreads();
    const int nob = (ks + BW_KG_OB - 1) / BW_KG_OB;      // offset blocks per axis
    const int ob = threadIdx.x >> 3, rg = threadIdx.x & 7;
    const int iy0 = (ob / nob) * BW_KG_OB, ix0 = (ob - (ob / nob) * nob) * BW_KG_OB;
    float acc[BW_KG_OB][BW_KG_OB];
#pragma unroll
    for (int a = 0; a < BW_KG_OB; ++a)
#pragma unroll
        for (int c = 0; c < BW_KG_OB; ++c) acc[a][c] = 0.0f;
    if (ob < nob * nob) {
        // offsets past the kernel edge are computed on clamped addresses and dropped at the end
        const int colbase = 2 * pad - min(ix0, ks - 1);
        for (int r = 0; r < 4; ++r) {
            const int ty = 4 * rg + r;
            const float* vr[BW_KG_OB];
            float w[BW_KG_OB][BW_KG_OB];
#pragma unroll
            for (int a = 0; a < BW_KG_OB; ++a) {
                const int iy = min(iy0 + a, ks - 1);
                vr[a] = sv + (ty + 2 * pad - iy) * VS + colbase;
#pragma unroll
                for (int c = 1; c < BW_KG_OB; ++c) w[a][c] = (colbase - c >= 0) ? vr[a][-c] : 0.0f;
            }
            const float* gr = sg + ty * GS;
#pragma unroll          // fully: the sliding window then lives in renamed registers, no moves
            for (int tx = 0; tx < BW_KG_TILE; ++tx) {
                const float sgv = gr[tx];
#pragma unroll
                for (int a = 0; a < BW_KG_OB; ++a) {
                    w[a][0] = vr[a][tx];
#pragma unroll
                    for (int c = 0; c < BW_KG_OB; ++c) acc[a][c] = fmaf(sgv, w[a][c], acc[a][c]);
#pragma unroll
                    for (int c = BW_KG_OB - 1; c > 0; --c) w[a][c] = w[a][c - 1];
                }
            }
        }
    }
    // the 8 row groups of an offset block sit in adjacent lanes: butterfly sum, lane 0 of the group adds to global
#pragma unroll
    for (int a = 0; a < BW_KG_OB; ++a)
#pragma unroll
        for (int c = 0; c < BW_KG_OB; ++c) {
            float v = acc[a][c];
            v += __shfl_xor_sync(0xffffffffu, v, 1);
            v += __shfl_xor_sync(0xffffffffu, v, 2);
            v += __shfl_xor_sync(0xffffffffu, v, 4);
            if (rg == 0 && ob < nob * nob && iy0 + a < ks && ix0 + c < ks)
                atomicAdd(&kbar[(size_t)b * ks * ks + (iy0 + a) * ks + ix0 + c], v);
        }
}

// m~ -> sparse gradient planes at the arg-max pixels
__global__ void k_bw_scatter(const float* __restrict__ mbar, const float* __restrict__ trace_f,
                             const int* __restrict__ trace_pos, float* __restrict__ sgx, float* __restrict__ sgy,
                             int B, size_t plane) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * 7) return;
    const int b = i / 7, j = i - b * 7;
    const float w = mbar[i] * trace_f[(size_t)b * PB_BW_TRACE_STRIDE + 7 + j];
    const size_t o = (size_t)b * plane + (size_t)trace_pos[b * 8 + j];
    atomicAdd(&sgx[o], w * c_bw_cos7[j]);
    atomicAdd(&sgy[o], -w * c_bw_sin7[j]);
}

// sums of gn~ and gn~ g per image, gn~ = -(dx + dy)
__global__ void k_bw_norm_reduce(const float* __restrict__ dx, const float* __restrict__ dy,
                                 const float* __restrict__ img, double* __restrict__ sums, int C, size_t plane) {
    const int b = blockIdx.y;
    const float* src = img + (size_t)b * C * plane;
    double s1 = 0.0, s2 = 0.0;
    for (size_t o = (size_t)blockIdx.x * blockDim.x + threadIdx.x; o < plane; o += (size_t)gridDim.x * blockDim.x) {
        const float gb = -(dx[(size_t)b * plane + o] + dy[(size_t)b * plane + o]);
        if (gb != 0.0f) {
            s1 += (double)gb;
            s2 += (double)gb * (double)bw_gray_of(src, plane, C, o);
        }
    }
    for (int s = 16; s > 0; s >>= 1) {
        s1 += __shfl_xor_sync(0xffffffffu, s1, s);
        s2 += __shfl_xor_sync(0xffffffffu, s2, s);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(&sums[2 * b], s1);
        atomicAdd(&sums[2 * b + 1], s2);
    }
}

// g~ = gn~ / r + [g = mn] mn~ / #min + [g = mx] mx~ / #max, r = mx - mn; every channel gets g~ / C
__global__ void k_bw_norm_apply(const float* __restrict__ dx, const float* __restrict__ dy,
                                const float* __restrict__ img, const float* __restrict__ trace_f,
                                const double* __restrict__ sums, float* __restrict__ gin, int C, size_t plane) {
    const int b = blockIdx.y;
    const float* tf = trace_f + (size_t)b * PB_BW_TRACE_STRIDE;
    const float mn = tf[14], mx = tf[15];
    const double r = (double)mx - (double)mn;
    const double s1 = sums[2 * b], s2 = sums[2 * b + 1];
    // d/dmn [(g - mn) / (mx - mn)] = (g - mx) / r^2,  d/dmx = -(g - mn) / r^2
    const float mnbar = (float)((s2 - (double)mx * s1) / (r * r) / (double)tf[16]);
    const float mxbar = (float)(-(s2 - (double)mn * s1) / (r * r) / (double)tf[17]);
    const float inv_r = (float)(1.0 / r), inv_c = 1.0f / (float)C;
    const float* src = img + (size_t)b * C * plane;
    float* dst = gin + (size_t)b * C * plane;
    for (size_t o = (size_t)blockIdx.x * blockDim.x + threadIdx.x; o < plane; o += (size_t)gridDim.x * blockDim.x) {
        const float gv = bw_gray_of(src, plane, C, o);
        float gb = -(dx[(size_t)b * plane + o] + dy[(size_t)b * plane + o]) * inv_r;
        if (gv == mn) gb += mnbar;
        if (gv == mx) gb += mxbar;
        if (gb != 0.0f) {
            gb *= inv_c;
            for (int c = 0; c < C; ++c) dst[(size_t)c * plane + o] += gb;
        }
    }
}

// ---- host side ------------------------------------------------------------------------------
static dim3 bw_grid(size_t plane, int B) {
    size_t nb = (plane + 255) / 256;
    if (nb > 1184) nb = 1184;
    return dim3((unsigned)nb, (unsigned)B);
}

int launch_bw_trace(const float* img, float* g, float* gn, unsigned* stats, unsigned long long* keys, int B, int C,
                    int H, int W, cudaStream_t stream) {
    const size_t plane = (size_t)H * W;
    ProfScope prof(PROF_OTHER, stream);
    k_bw_init<<<(B * 7 + 255) / 256, 256, 0, stream>>>(stats, keys, B);
    k_bw_gray<<<bw_grid(plane, B), 256, 0, stream>>>(img, g, stats, C, plane);
    k_bw_normalize<<<bw_grid(plane, B), 256, 0, stream>>>(g, gn, stats, plane);
    PB_LAUNCH_CHECK("k_bw_gray / k_bw_normalize");
    return PB_OK;
}

int launch_bw_dirmax(const float* gx, const float* gy, const float* g, unsigned long long* keys, unsigned* stats,
                     float* trace_f, int* trace_pos, int B, int H, int W, int discard_saturation, cudaStream_t stream) {
    const size_t plane = (size_t)H * W;
    ProfScope prof(PROF_OTHER, stream);
    k_bw_dirmax<<<bw_grid(plane, B), 256, 0, stream>>>(gx, gy, g, keys, stats, plane, discard_saturation);
    k_bw_finish<<<B, 32, 0, stream>>>(gx, gy, g, keys, stats, trace_f, trace_pos, plane, discard_saturation);
    PB_LAUNCH_CHECK("k_bw_dirmax / k_bw_finish");
    return PB_OK;
}

int launch_bw_kernel_grad(const float* gout, const float* preclamp, const float* V, float* kbar, int B, int C, int H,
                          int W, int ks, cudaStream_t stream) {
    if (B * C > 65535) {
        set_error("B*C = %d exceeds the grid z limit", B * C);
        return PB_ERR_ARG;
    }
    const int pad = ks >> 1, VT = BW_KG_TILE + 2 * pad;
    const size_t smem = (size_t)(BW_KG_TILE * (BW_KG_TILE + 1) + VT * (VT + 1)) * sizeof(float);
    dim3 grid((W + BW_KG_TILE - 1) / BW_KG_TILE, (H + BW_KG_TILE - 1) / BW_KG_TILE, B * C);
    ProfScope prof(PROF_OTHER, stream);
    PB_CUDA_TRY(cudaMemsetAsync(kbar, 0, (size_t)B * ks * ks * sizeof(float), stream));
    k_bw_kernel_grad<<<grid, 256, smem, stream>>>(gout, preclamp, V, kbar, C, H, W, ks);
    PB_LAUNCH_CHECK("k_bw_kernel_grad");
    return PB_OK;
}

int launch_bw_scatter(const float* mbar, const float* trace_f, const int* trace_pos, float* sgx, float* sgy, int B,
                      int H, int W, cudaStream_t stream) {
    const size_t plane = (size_t)H * W;
    ProfScope prof(PROF_OTHER, stream);
    PB_CUDA_TRY(cudaMemsetAsync(sgx, 0, (size_t)B * plane * sizeof(float), stream));
    PB_CUDA_TRY(cudaMemsetAsync(sgy, 0, (size_t)B * plane * sizeof(float), stream));
    k_bw_scatter<<<(B * 7 + 127) / 128, 128, 0, stream>>>(mbar, trace_f, trace_pos, sgx, sgy, B, plane);
    PB_LAUNCH_CHECK("k_bw_scatter");
    return PB_OK;
}

int launch_bw_norm(const float* dx, const float* dy, const float* img, const float* trace_f, double* sums, float* gin,
                   int B, int C, int H, int W, cudaStream_t stream) {
    const size_t plane = (size_t)H * W;
    ProfScope prof(PROF_OTHER, stream);
    PB_CUDA_TRY(cudaMemsetAsync(sums, 0, (size_t)B * 2 * sizeof(double), stream));
    k_bw_norm_reduce<<<bw_grid(plane, B), 256, 0, stream>>>(dx, dy, img, sums, C, plane);
    k_bw_norm_apply<<<bw_grid(plane, B), 256, 0, stream>>>(dx, dy, img, trace_f, sums, gin, C, plane);
    PB_LAUNCH_CHECK("k_bw_norm_reduce / k_bw_norm_apply");
    return PB_OK;
}

}  // namespace pb
