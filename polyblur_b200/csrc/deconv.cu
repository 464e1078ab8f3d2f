// Spatial deconvolution engine: the degree-3 polynomial of the blur applied as three stacked
// sparse stencils (Horner), fused in shared memory.
//
// Reference being replaced: deblurring.inverse_filtering_rank3 with default flags
// (polyblur/deblurring.py:211-239) = utils.pad_with_kernel (utils.py:48-53) ->
// compute_polynomial_fft (deblurring.py:141-169) -> utils.crop_with_kernel (utils.py:56-61) ->
// clamp.  The FFT product there is a circular correlation on the replicate-padded torus of
// size (H+2P) x (W+2P); SURVEY.md A.6 shows it equals
//     o = a3 p;  o = K (*) o + a2 p;  o = K (*) o + a1 p;  o = K (*) o + b p
// with one gather map  src = clamp((coord mod (n + 2P)) - P, 0, n - 1).
//
// One CTA produces a 64 x 64 output tile of one channel: it gathers the tile plus a halo of
// 3 r (r = radius of the taps kept for this image, <= 12) through the torus map, then runs the
// three stencils shrinking the halo by r each time.  Only taps above the relative threshold
// are applied (per-row ranges from k_params), so a near-delta kernel costs ~9 FMA per stencil.
// HBM traffic: 4 B read (+ halo re-reads that hit L2) and 4 B written per pixel-channel.
#include "kernels.cuh"

namespace pb {

#define DT_W 64
#define DT_H 64
#define DC_THREADS 512
#define DC_MAXEXT (DT_W + 6 * PB_PAD)   // 136

struct DeconvSmem {
    float k[640];
    int lo[32];
    int hi[32];
    int srcx[DC_MAXEXT + 8];
    int srcy[DC_MAXEXT + 8];
};

template <int HX, bool FINAL>
__device__ __forceinline__ void conv_stage(const float* __restrict__ in, int in_stride,
                                           float* __restrict__ outb, int out_stride, int RH, int RW,
                                           int hy, const DeconvSmem& S, const float* __restrict__ P,
                                           int p_stride, int poff_y, int poff_x, float cK, float cP,
                                           float* __restrict__ gout, int gy0, int gx0, int H, int W,
                                           int clamp_out) {
    const int nsx = RW >> 2;
    const int nstrips = RH * nsx;
    for (int s = threadIdx.x; s < nstrips; s += blockDim.x) {
        const int yo = s / nsx;
        const int xo = (s - yo * nsx) << 2;
        float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
        for (int dyi = PB_PAD - hy; dyi <= PB_PAD + hy; ++dyi) {
            const int lo = S.lo[dyi], hi = S.hi[dyi];
            if (lo > hi) continue;
            const float* row = in + (yo + hy + dyi - PB_PAD) * in_stride + xo;
            float seg[2 * HX + 4];
#pragma unroll
            for (int q = 0; q < (2 * HX + 4) / 4; ++q) {
                float4 t = *reinterpret_cast<const float4*>(row + 4 * q);
                seg[4 * q + 0] = t.x;
                seg[4 * q + 1] = t.y;
                seg[4 * q + 2] = t.z;
                seg[4 * q + 3] = t.w;
            }
            const float* wrow = S.k + dyi * PB_KS + PB_PAD - HX;
            // two-level summation (row partial sums, then rows): keeps the rounding noise of a
            // 625-tap stencil near that of the reference's fp32 FFT product
            float r0 = 0.f, r1 = 0.f, r2 = 0.f, r3 = 0.f;
#pragma unroll
            for (int t = 0; t <= 2 * HX; ++t) {
                const int ti = PB_PAD - HX + t;
                if (ti >= lo && ti <= hi) {
                    const float w = wrow[t];
                    r0 = fmaf(w, seg[t + 0], r0);
                    r1 = fmaf(w, seg[t + 1], r1);
                    r2 = fmaf(w, seg[t + 2], r2);
                    r3 = fmaf(w, seg[t + 3], r3);
                }
            }
            acc0 += r0;
            acc1 += r1;
            acc2 += r2;
            acc3 += r3;
        }
        const float4 pv = *reinterpret_cast<const float4*>(P + (yo + poff_y) * p_stride + xo + poff_x);
        float4 r;
        r.x = fmaf(cK, acc0, cP * pv.x);
        r.y = fmaf(cK, acc1, cP * pv.y);
        r.z = fmaf(cK, acc2, cP * pv.z);
        r.w = fmaf(cK, acc3, cP * pv.w);
        if (FINAL) {
            const int y = gy0 + yo;
            if (y < H) {
                float* g = gout + (size_t)y * W + gx0 + xo;
                const float v[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (gx0 + xo + j < W) g[j] = clamp_out ? fminf(fmaxf(v[j], 0.0f), 1.0f) : v[j];
            }
        } else {
            *reinterpret_cast<float4*>(outb + yo * out_stride + xo) = r;
        }
    }
}

template <int HX>
__device__ __forceinline__ void horner_tile(float* P, float* O1, float* O2, int hy, const DeconvSmem& S,
                                            float a3, float a2, float a1, float b0, float* gout,
                                            int gy0, int gx0, int H, int W, int clamp_out) {
    const int PW = DT_W + 6 * HX, W1 = DT_W + 4 * HX, W2 = DT_W + 2 * HX;
    conv_stage<HX, false>(P, PW, O1, W1, DT_H + 4 * hy, W1, hy, S, P, PW, hy, HX, a3, a2,
                          nullptr, 0, 0, 0, 0, 0);
    __syncthreads();
    conv_stage<HX, false>(O1, W1, O2, W2, DT_H + 2 * hy, W2, hy, S, P, PW, 2 * hy, 2 * HX, 1.0f, a1,
                          nullptr, 0, 0, 0, 0, 0);
    __syncthreads();
    conv_stage<HX, true>(O2, W2, nullptr, 0, DT_H, DT_W, hy, S, P, PW, 3 * hy, 3 * HX, 1.0f, b0,
                         gout, gy0, gx0, H, W, clamp_out);
}

__global__ void __launch_bounds__(DC_THREADS, 2)
k_deconv_spatial(const float* __restrict__ img, float* __restrict__ out,
                 const ImgKernel* __restrict__ kern, const int* __restrict__ list,
                 const int* __restrict__ count, int C, int H, int W,
                 float a3, float a2, float a1, float b0, SrcGeom G) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    DeconvSmem& S = *reinterpret_cast<DeconvSmem*>(smem_raw);
    float* bufs = reinterpret_cast<float*>(smem_raw + sizeof(DeconvSmem));
    const int tilesX = (W + DT_W - 1) / DT_W, tilesY = (H + DT_H - 1) / DT_H;
    const int per_plane = tilesX * tilesY;
    const int per_img = C * per_plane;
    const int total = count[0] * per_img;      // images of the tiled class (list filled by k_params)

    for (int w = blockIdx.x; w < total; w += gridDim.x) {
        const int slot = w / per_img;
        int rr = w - slot * per_img;
        const int c = rr / per_plane;
        rr -= c * per_plane;
        const int tyi = rr / tilesX;
        const int txi = rr - tyi * tilesX;
        const int im = list[slot];
        const ImgKernel* K = kern + im;
        const int r = K->radius;
        const int pad = G.pad >= 0 ? G.pad : K->ksize / 2;
        const int hy = r;
        int HX = (r + 3) & ~3;
        if (HX == 0) HX = 4;

        // taps of D = K - I (centre tap minus one): see poly_coeffs_d in api.cu
        for (int i = threadIdx.x; i < 640; i += blockDim.x)
            S.k[i] = (i < PB_KS2) ? K->k[i] - (i == PB_PAD * PB_KS + PB_PAD ? 1.0f : 0.0f) : 0.0f;
        if (threadIdx.x < 32) {
            S.lo[threadIdx.x] = (threadIdx.x < PB_KS) ? K->lo[threadIdx.x] : PB_KS;
            S.hi[threadIdx.x] = (threadIdx.x < PB_KS) ? K->hi[threadIdx.x] : -1;
        }
        const int gx0 = txi * DT_W, gy0 = tyi * DT_H;
        const int PW = DT_W + 6 * HX, PH = DT_H + 6 * hy;
        for (int i = threadIdx.x; i < PW; i += blockDim.x) S.srcx[i] = geom_src(gx0 + i - 3 * HX, G.Win, G.off, pad);
        for (int i = threadIdx.x; i < PH; i += blockDim.x) S.srcy[i] = geom_src(gy0 + i - 3 * hy, G.Hin, G.off, pad);
        __syncthreads();

        float* P = bufs;
        float* O1 = P + PH * PW;
        float* O2 = O1 + (DT_H + 4 * hy) * (DT_W + 4 * HX);
        const size_t pl = ((size_t)im * C + c) * H * W;
        const float* plane = img + ((size_t)im * C + c) * (size_t)G.Hin * G.Win;
        for (int i = threadIdx.x; i < PH * PW; i += blockDim.x) {
            const int ly = i / PW, lx = i - ly * PW;
            P[i] = __ldg(plane + (size_t)S.srcy[ly] * G.Win + S.srcx[lx]);
        }
        __syncthreads();
        float* gout = out + pl;
        switch (HX) {
            case 4: horner_tile<4>(P, O1, O2, hy, S, a3, a2, a1, b0, gout, gy0, gx0, H, W, G.clamp_out); break;
            case 8: horner_tile<8>(P, O1, O2, hy, S, a3, a2, a1, b0, gout, gy0, gx0, H, W, G.clamp_out); break;
            default: horner_tile<12>(P, O1, O2, hy, S, a3, a2, a1, b0, gout, gy0, gx0, H, W, G.clamp_out); break;
        }
        __syncthreads();
    }
}

int launch_deconv_spatial(const float* img, float* out, const ImgKernel* kern, const int* list, const int* count,
                          int B, int C, int H, int W, float a3, float a2, float a1, float b0, const SrcGeom& G,
                          int max_radius, cudaStream_t stream) {
    // shared memory for the largest halo of this class (radius <= 4: 77 KB, two CTAs per SM)
    const int r = max_radius < PB_PAD ? max_radius : PB_PAD;
    const size_t smem_of_r = sizeof(DeconvSmem) +
                             sizeof(float) * ((size_t)(DT_W + 6 * r) * (DT_H + 6 * r) + (size_t)(DT_W + 4 * r) * (DT_H + 4 * r) +
                                              (size_t)(DT_W + 2 * r) * (DT_H + 2 * r));
    const int rmax = PB_PAD;
    const size_t smem_max = sizeof(DeconvSmem) +
                            sizeof(float) * ((size_t)(DT_W + 6 * rmax) * (DT_H + 6 * rmax) +
                                             (size_t)(DT_W + 4 * rmax) * (DT_H + 4 * rmax) +
                                             (size_t)(DT_W + 2 * rmax) * (DT_H + 2 * rmax));
    const size_t smem = smem_of_r;
    PB_CUDA_TRY(cudaFuncSetAttribute(k_deconv_spatial, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max));
    const long long items = (long long)B * C * ((W + DT_W - 1) / DT_W) * ((H + DT_H - 1) / DT_H);
    const int grid = (int)(items < 4LL * PB_NUM_SMS ? items : 4LL * PB_NUM_SMS);   // persistent, 1 CTA / SM resident
    ProfScope prof(PROF_DECONV_SPATIAL, stream);
    k_deconv_spatial<<<grid, DC_THREADS, smem, stream>>>(img, out, kern, list, count, C, H, W, a3, a2, a1, b0, G);
    PB_LAUNCH_CHECK("k_deconv_spatial");
    return PB_OK;
}

}  // namespace pb
