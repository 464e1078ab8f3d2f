// Blur-independent deconvolution engine: the polynomial of the blur applied in the Fourier
// domain on chip, for images whose estimated kernel is too wide for the stencil engines.
//
// Reference being replaced: deblurring.inverse_filtering_rank3 with default flags
// (polyblur/deblurring.py:211-239): utils.pad_with_kernel (utils.py:48-53) ->
// compute_polynomial_fft (deblurring.py:141-169: fft2, p2o, Horner in the frequency domain,
// ifft2) -> utils.crop_with_kernel -> clamp.
//
// The reference transforms the replicate-padded image on a torus of (H+2P) x (W+2P), sizes with
// awkward prime factors (1104 = 2^4*3*23).  Only output pixels inside the crop are kept and the
// composite filter reaches 3P, so the same numbers come out of ANY torus that holds the periodic
// extension of the padded image over [-2P, n+4P) without wrap-around overlap:
// NY x NX = the next lengths >= n + 6P that factor into three or four radices <= 16.
//
//   P1 k_fft_rows_fwd : gathers two extended rows of one plane through the torus map into one
//                       complex sequence, DIF transform along x (fft2.cuh), separates the two
//                       Hermitian half spectra and writes them transposed: Z[plane][kx][j].
//                       (kx = 0 holds DC and Nyquist, both real, as one complex number.)
//   P2 k_fft_cols     : per image and block of columns kx: bulk-copies the contiguous columns
//                       into shared memory (cp.async.bulk + mbarrier), DIF along y, multiplies by
//                       H = ((a3 K^ + a2) K^ + a1) K^ + b with K^(ky,kx) evaluated from the 25x25
//                       taps (once per column block, reused by the C planes), DIT back, bulk-store.
//   P3 k_fft_rows_inv : rebuilds the full spectrum of a row pair, DIT transform along x, crops,
//                       clamps to [0,1] and writes the output rows.
//
// HBM bytes per pixel-channel: P1 4 read + ~4.4 written, P2 ~4.4 + ~4.4, P3 ~4.4 read + 4
// written = ~26 B against 8 B algorithmic; independent of the blur.
#include "kernels.cuh"

namespace pb {

// ---- PTX wrappers: mbarrier + 1-D bulk (TMA) copies ---------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE;\n"
        "bra WAIT_LOOP;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* dst, const void* src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- tables -------------------------------------------------------------------------------
// slot[k] = position of frequency k after the DIF transform; freq[p] = its inverse.
__global__ void k_fft2_perm(int* __restrict__ slot_of_freq, int* __restrict__ freq_of_slot, Fft2Plan plan) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < plan.n) {
        if (slot_of_freq) slot_of_freq[i] = fft2_slot_of_freq(i, plan);
        if (freq_of_slot) freq_of_slot[i] = fft2_freq_of_slot(i, plan);
    }
}

#define FFTD_THREADS 256

// extended coordinate (any torus of length >= n + 6 pad) -> source index, or -1 for the zero fill
__device__ __forceinline__ int ext_src(int i, int n, int pad) {
    if (i >= n + 6 * pad) return -1;
    const int X = i - 3 * pad;                   // image coordinate of this extended sample
    if (X >= 0 && X < n) return X;
    return torus_src(X + pad, n, pad);           // padded coordinate = image coordinate + pad
}

// ---------------------------------------------------------------------------------------------
// P1: rows forward.  Work item = (slot in the FFT class list, channel, block of nb row pairs).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(FFTD_THREADS)
k_fft_rows_fwd(const float* __restrict__ img, float2* __restrict__ Z, const ImgKernel* __restrict__ kern,
               const int* __restrict__ list, const int* __restrict__ count, int C, int H, int W, int NX, int NY,
               int nb, Fft2Plan planX, const float2* __restrict__ twX, const int* __restrict__ slotX) {
    extern __shared__ __align__(16) float2 smf[];
    const int tid = threadIdx.x;
    const int blocks_per_plane = (NY / 2 + nb - 1) / nb;
    const int per_img = C * blocks_per_plane;
    const int total = count[0] * per_img;
    const size_t plane = (size_t)H * W;
    const int half = NX >> 1;
    const float inv_nx = 1.0f / (float)NX, inv_half = 1.0f / (float)half;

    for (int w = blockIdx.x; w < total; w += gridDim.x) {
        const int slot = w / per_img;
        int r = w - slot * per_img;
        const int c = r / blocks_per_plane;
        const int rb = r - c * blocks_per_plane;
        const int im = list[slot];
        const int pad = kern[im].ksize >> 1;
        const float* src = img + ((size_t)im * C + c) * plane;
        const int j0 = rb * 2 * nb;

        for (int idx = tid; idx < nb * NX; idx += FFTD_THREADS) {
            const int p = fast_div(idx, NX, inv_nx);
            const int i = idx - p * NX;
            const int sx = ext_src(i, W, pad);
            float2 v = make_float2(0.f, 0.f);
            if (sx >= 0) {
                const int ja = j0 + 2 * p;
                const int sa = (ja < NY) ? ext_src(ja, H, pad) : -1;
                const int sb = (ja + 1 < NY) ? ext_src(ja + 1, H, pad) : -1;
                if (sa >= 0) v.x = __ldg(src + (size_t)sa * W + sx);
                if (sb >= 0) v.y = __ldg(src + (size_t)sb * W + sx);
            }
            smf[(size_t)p * NX + i] = v;
        }
        __syncthreads();
        fft2_forward_dif(smf, NX, nb, planX, twX, tid, FFTD_THREADS);
        // separate the two real rows: Xa[k] = (Z[k] + conj Z[-k]) / 2, Xb[k] = (Z[k] - conj Z[-k]) / (2i)
        float2* Zp = Z + ((size_t)slot * C + c) * half * NY;
        for (int idx = tid; idx < nb * half; idx += FFTD_THREADS) {
            const int kx = fast_div(idx, nb, 1.0f / (float)nb);
            const int p = idx - kx * nb;
            const int ja = j0 + 2 * p;
            if (ja >= NY) continue;
            const float2* row = smf + (size_t)p * NX;
            float2 xa, xb;
            if (kx == 0) {
                const float2 z0 = row[__ldg(slotX)];
                const float2 zn = row[__ldg(slotX + half)];
                xa = make_float2(z0.x, zn.x);
                xb = make_float2(z0.y, zn.y);
            } else {
                const float2 z1 = row[__ldg(slotX + kx)];
                const float2 z2 = row[__ldg(slotX + NX - kx)];
                xa = make_float2(0.5f * (z1.x + z2.x), 0.5f * (z1.y - z2.y));
                xb = make_float2(0.5f * (z1.y + z2.y), 0.5f * (z2.x - z1.x));
            }
            float2* d = Zp + (size_t)kx * NY + ja;
            if (ja + 1 < NY) {
                *reinterpret_cast<float4*>(d) = make_float4(xa.x, xa.y, xb.x, xb.y);
            } else {
                d[0] = xa;
            }
        }
        __syncthreads();
        (void)inv_half;
    }
}

// ---------------------------------------------------------------------------------------------
// P2: columns.  Work item = (slot, block of CB columns); loops over the C planes of the image.
//   shared memory: data[CB][NY] float2 | Hs[CB][NY] float | Hn[NY] float (Nyquist, block 0)
//                  | R[CB+1][13] float2 | mbarrier
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(FFTD_THREADS)
k_fft_cols(float2* __restrict__ Z, const ImgKernel* __restrict__ kern, const int* __restrict__ list,
           const int* __restrict__ count, int C, int NX, int NY, int CB, Fft2Plan planY,
           const float2* __restrict__ twX, const float2* __restrict__ twY, const int* __restrict__ freqY,
           const int* __restrict__ slotY, float a3, float a2, float a1, float b0) {
    extern __shared__ __align__(16) unsigned char smraw[];
    float2* data = reinterpret_cast<float2*>(smraw);
    float* Hs = reinterpret_cast<float*>(data + (size_t)CB * NY);
    float* Hn = Hs + (size_t)CB * NY;
    float2* Rk = reinterpret_cast<float2*>(Hn + NY);                  // [(CB + 1)][13]
    uint64_t* bar = reinterpret_cast<uint64_t*>(Rk + (CB + 1) * 13 + 1);
    bar = reinterpret_cast<uint64_t*>(((uintptr_t)bar + 7) & ~(uintptr_t)7);
    const int tid = threadIdx.x;
    const int half = NX >> 1;
    const int nblk = (half + CB - 1) / CB;
    const int total = count[0] * nblk;
    const float scale = 1.0f / ((float)NX * (float)NY);
    const float inv_ny = 1.0f / (float)NY;
    uint32_t phase = 0;
    if (tid == 0) mbar_init(bar, 1);
    __syncthreads();

    for (int w = blockIdx.x; w < total; w += gridDim.x) {
        const int slot = w / nblk;
        const int cb = w - slot * nblk;
        const int im = list[slot];
        const ImgKernel* K = kern + im;
        const int kx0 = cb * CB;
        const int ncol = min(CB, half - kx0);

        // R[col][dy] = sum_dx K[dy][dx] exp(-2 pi i kx dx / NX), dy = 0..12 (R[-dy] = conj R[dy]);
        // entry CB is the Nyquist column kx = NX / 2 (needed by the block that holds kx = 0)
        for (int idx = tid; idx < (CB + 1) * 13; idx += FFTD_THREADS) {
            const int col = idx / 13, dy = idx - col * 13;
            int kx = (col < CB) ? kx0 + col : half;
            float2 acc = make_float2(0.f, 0.f);
            if (col < ncol || (col == CB && cb == 0)) {
                for (int dx = -PB_PAD; dx <= PB_PAD; ++dx) {
                    const float kv = __ldg(&K->k[(dy + PB_PAD) * PB_KS + dx + PB_PAD]);
                    int t = (int)(((long long)kx * (dx + NX)) % NX);
                    const float2 e = __ldg(twX + t);
                    acc.x = fmaf(kv, e.x, acc.x);
                    acc.y = fmaf(kv, e.y, acc.y);
                }
            }
            Rk[idx] = acc;
        }
        __syncthreads();
        // Hs[col][slot] = scale * P(K^(ky(slot), kx))
        for (int idx = tid; idx < (ncol + (cb == 0 ? 1 : 0)) * NY; idx += FFTD_THREADS) {
            int col = fast_div(idx, NY, inv_ny);
            const int s = idx - col * NY;
            const bool nyq = (col == ncol);
            if (nyq) col = CB;
            const int ky = __ldg(freqY + s);
            const float2* R = Rk + col * 13;
            float kh = R[0].x;
            int t = 0;
#pragma unroll 4
            for (int dy = 1; dy <= PB_PAD; ++dy) {
                t += ky;
                if (t >= NY) t -= NY;
                const float2 e = __ldg(twY + t);            // (cos, -sin)(2 pi ky dy / NY)
                kh = fmaf(2.0f * R[dy].x, e.x, kh);
                kh = fmaf(-2.0f * R[dy].y, e.y, kh);
            }
            const float h = fmaf(fmaf(fmaf(a3, kh, a2), kh, a1), kh, b0) * scale;
            if (nyq) Hn[s] = h; else Hs[(size_t)col * NY + s] = h;
        }
        __syncthreads();

        for (int c = 0; c < C; ++c) {
            float2* Zc = Z + (((size_t)slot * C + c) * half + kx0) * NY;
            const uint32_t bytes = (uint32_t)((size_t)ncol * NY * sizeof(float2));
            if (tid == 0) {
                fence_async_smem();
                mbar_expect_tx(bar, bytes);
                for (int col = 0; col < ncol; ++col)
                    bulk_g2s(data + (size_t)col * NY, Zc + (size_t)col * NY, (uint32_t)(NY * sizeof(float2)), bar);
            }
            mbar_wait(bar, phase);
            phase ^= 1;
            fft2_forward_dif(data, NY, ncol, planY, twY, tid, FFTD_THREADS);
            // multiply by H and swap re/im (inverse by forward transform)
            if (cb == 0) {
                // column 0 carries DC (real part) and Nyquist (imaginary part) of two real spectra:
                // Y[k] = (H0 + Hn)/2 Z[k] + (H0 - Hn)/2 conj Z[-k]; the pair {k, -k} goes to one thread
                for (int ky = tid; ky <= NY / 2; ky += FFTD_THREADS) {
                    const int s1 = __ldg(slotY + ky);
                    const int s2 = __ldg(slotY + (NY - ky) % NY);
                    const float2 z1 = data[s1], z2 = data[s2];
                    const float A1 = 0.5f * (Hs[s1] + Hn[s1]), B1 = 0.5f * (Hs[s1] - Hn[s1]);
                    const float A2 = 0.5f * (Hs[s2] + Hn[s2]), B2 = 0.5f * (Hs[s2] - Hn[s2]);
                    const float2 y1 = make_float2(A1 * z1.x + B1 * z2.x, A1 * z1.y - B1 * z2.y);
                    const float2 y2 = make_float2(A2 * z2.x + B2 * z1.x, A2 * z2.y - B2 * z1.y);
                    data[s1] = make_float2(y1.y, y1.x);
                    if (s2 != s1) data[s2] = make_float2(y2.y, y2.x);
                }
                for (int idx = NY + tid; idx < ncol * NY; idx += FFTD_THREADS) {
                    const float h = Hs[idx];
                    const float2 z = data[idx];
                    data[idx] = make_float2(h * z.y, h * z.x);
                }
            } else {
                for (int idx = tid; idx < ncol * NY; idx += FFTD_THREADS) {
                    const float h = Hs[idx];
                    const float2 z = data[idx];
                    data[idx] = make_float2(h * z.y, h * z.x);
                }
            }
            __syncthreads();
            fft2_forward_dit(data, NY, ncol, planY, twY, tid, FFTD_THREADS);
            // result r = DFT(swap(Y)); the inverse transform is swap(r): swap back while storing
            for (int idx = tid; idx < ncol * NY; idx += FFTD_THREADS) {
                const float2 z = data[idx];
                data[idx] = make_float2(z.y, z.x);
            }
            fence_async_smem();
            __syncthreads();
            if (tid == 0) {
                for (int col = 0; col < ncol; ++col)
                    bulk_s2g(Zc + (size_t)col * NY, data + (size_t)col * NY, (uint32_t)(NY * sizeof(float2)));
                bulk_commit();
                bulk_wait_all();
            }
            __syncthreads();
        }
    }
}

// ---------------------------------------------------------------------------------------------
// P3: rows inverse.  Work item = (slot, channel, block of nb row pairs that hold output rows).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(FFTD_THREADS)
k_fft_rows_inv(const float2* __restrict__ Z, float* __restrict__ out, const ImgKernel* __restrict__ kern,
               const int* __restrict__ list, const int* __restrict__ count, int C, int H, int W, int NX, int NY,
               int nb, Fft2Plan planX, const float2* __restrict__ twX, const int* __restrict__ slotX) {
    extern __shared__ __align__(16) float2 smf[];
    const int tid = threadIdx.x;
    const int blocks_per_plane = (NY / 2 + nb - 1) / nb;
    const int per_img = C * blocks_per_plane;
    const int total = count[0] * per_img;
    const size_t plane = (size_t)H * W;
    const int half = NX >> 1;
    const float inv_w = 1.0f / (float)W;

    for (int w = blockIdx.x; w < total; w += gridDim.x) {
        const int slot = w / per_img;
        int r = w - slot * per_img;
        const int c = r / blocks_per_plane;
        const int rb = r - c * blocks_per_plane;
        const int im = list[slot];
        const int pad = kern[im].ksize >> 1;
        const int j0 = rb * 2 * nb;
        // rows of the extended image that are output rows: [3 pad, H + 3 pad)
        if (j0 + 2 * nb <= 3 * pad || j0 >= H + 3 * pad) continue;
        const float2* Zp = Z + ((size_t)slot * C + c) * half * NY;
        for (int idx = tid; idx < nb * half; idx += FFTD_THREADS) {
            const int kx = fast_div(idx, nb, 1.0f / (float)nb);
            const int p = idx - kx * nb;
            const int ja = j0 + 2 * p;
            float2 xa = make_float2(0.f, 0.f), xb = make_float2(0.f, 0.f);
            if (ja + 1 < NY) {
                const float4 v = __ldg(reinterpret_cast<const float4*>(Zp + (size_t)kx * NY + ja));
                xa = make_float2(v.x, v.y);
                xb = make_float2(v.z, v.w);
            } else if (ja < NY) {
                xa = __ldg(Zp + (size_t)kx * NY + ja);
            }
            float2* row = smf + (size_t)p * NX;
            // Z[k] = Xa[k] + i Xb[k], Z[-k] = conj Xa[k] + i conj Xb[k]; stored swapped (re <-> im)
            if (kx == 0) {
                row[__ldg(slotX)] = make_float2(xb.x, xa.x);
                row[__ldg(slotX + half)] = make_float2(xb.y, xa.y);
            } else {
                row[__ldg(slotX + kx)] = make_float2(xa.y + xb.x, xa.x - xb.y);
                row[__ldg(slotX + NX - kx)] = make_float2(xb.x - xa.y, xa.x + xb.y);
            }
        }
        __syncthreads();
        fft2_forward_dit(smf, NX, nb, planX, twX, tid, FFTD_THREADS);
        // r = DFT(swap(Z)): row a = r.y, row b = r.x (the 1/(NX NY) scale is inside H)
        float* dst = out + ((size_t)im * C + c) * plane;
        for (int idx = tid; idx < nb * W; idx += FFTD_THREADS) {
            const int p = fast_div(idx, W, inv_w);
            const int x = idx - p * W;
            const float2 z = smf[(size_t)p * NX + x + 3 * pad];
            const int ya = j0 + 2 * p - 3 * pad;
            if (ya >= 0 && ya < H) dst[(size_t)ya * W + x] = fminf(fmaxf(z.y, 0.0f), 1.0f);
            if (ya + 1 >= 0 && ya + 1 < H) dst[(size_t)(ya + 1) * W + x] = fminf(fmaxf(z.x, 0.0f), 1.0f);
        }
        __syncthreads();
    }
}

// ---- host side ------------------------------------------------------------------------------

// smallest even length >= n that the fft2 core does in <= 3 stages (<= 4 if none within 8 %)
int fft_engine_length(int n) {
    Fft2Plan p;
    int fallback = 0;
    for (int m = n + (n & 1); m < 2 * n + 64; m += 2) {
        if (make_fft2_plan(m, &p) != 0) continue;
        if (p.ns <= 3) return (fallback && (double)m > 1.08 * n) ? fallback : m;
        if (p.ns <= 4 && !fallback) fallback = m;
        if (fallback && (double)m > 1.08 * n) return fallback;
    }
    return fallback;
}

static int rows_nb(int NX) {
    int nb = (int)((64 * 1024) / ((size_t)NX * sizeof(float2)));
    if (nb < 1) nb = 1;
    if (nb > 8) nb = 8;
    return nb;
}
static int cols_cb(int NY) {
    int cb = (int)((100 * 1024) / ((size_t)NY * 12));
    if (cb < 1) cb = 1;
    if (cb > 8) cb = 8;
    return cb;
}

bool fft_engine_supported(int H, int W, int pad) {
    const int NX = fft_engine_length(W + 6 * pad), NY = fft_engine_length(H + 6 * pad);
    if (NX <= 0 || NY <= 0) return false;
    const size_t lim = PB_SMEM_MAX - 4096;
    return (size_t)NX * sizeof(float2) <= lim && (size_t)NY * 12 + 4 * NY + 512 <= lim;
}

size_t fft_engine_workspace(int B, int C, int H, int W, int pad, FftEngineLayout* L) {
    FftEngineLayout l;
    l.NX = fft_engine_length(W + 6 * pad);
    l.NY = fft_engine_length(H + 6 * pad);
    size_t o = 0;
    auto take = [&](size_t bytes) {
        size_t at = o;
        o = align_up(o + bytes, 256);
        return at;
    };
    l.off_twX = take((size_t)l.NX * sizeof(float2));
    l.off_twY = take((size_t)l.NY * sizeof(float2));
    l.off_slotX = take((size_t)l.NX * sizeof(int));
    l.off_slotY = take((size_t)l.NY * sizeof(int));
    l.off_freqY = take((size_t)l.NY * sizeof(int));
    l.off_Z = take((size_t)B * C * (l.NX / 2) * l.NY * sizeof(float2));
    l.total = o;
    if (L) *L = l;
    return o;
}

int fft_engine_prepare(char* base, const FftEngineLayout& L, FftEngineTables* T, cudaStream_t stream) {
    T->NX = L.NX;
    T->NY = L.NY;
    if (make_fft2_plan(L.NX, &T->planX) || make_fft2_plan(L.NY, &T->planY)) {
        set_error("no FFT plan for the %d x %d torus", L.NY, L.NX);
        return PB_ERR_UNSUPPORTED;
    }
    T->twX = reinterpret_cast<float2*>(base + L.off_twX);
    T->twY = reinterpret_cast<float2*>(base + L.off_twY);
    T->slotX = reinterpret_cast<int*>(base + L.off_slotX);
    T->slotY = reinterpret_cast<int*>(base + L.off_slotY);
    T->freqY = reinterpret_cast<int*>(base + L.off_freqY);
    T->Z = reinterpret_cast<float2*>(base + L.off_Z);
    int rc;
    if ((rc = launch_twiddles(T->twX, L.NX, stream))) return rc;
    if ((rc = launch_twiddles(T->twY, L.NY, stream))) return rc;
    ProfScope prof(PROF_SETUP, stream);
    k_fft2_perm<<<(L.NX + 255) / 256, 256, 0, stream>>>(T->slotX, nullptr, T->planX);
    k_fft2_perm<<<(L.NY + 255) / 256, 256, 0, stream>>>(T->slotY, T->freqY, T->planY);
    PB_LAUNCH_CHECK("k_fft2_perm");
    return PB_OK;
}

int launch_deconv_fft(const float* img, float* out, const ImgKernel* kern, const int* list, const int* count,
                      int B, int C, int H, int W, const FftEngineTables& T, float a3, float a2, float a1,
                      float b0, cudaStream_t stream) {
    const int NX = T.NX, NY = T.NY;
    const int nb = rows_nb(NX), CB = cols_cb(NY);
    const size_t smem_rows = (size_t)nb * NX * sizeof(float2);
    const size_t smem_cols = (size_t)CB * NY * 12 + (size_t)NY * 4 + (size_t)(CB + 1) * 13 * 8 + 64;
    PB_CUDA_TRY(cudaFuncSetAttribute(k_fft_rows_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_rows));
    PB_CUDA_TRY(cudaFuncSetAttribute(k_fft_rows_inv, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_rows));
    PB_CUDA_TRY(cudaFuncSetAttribute(k_fft_cols, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_cols));
    const long long row_items = (long long)B * C * ((NY / 2 + nb - 1) / nb);
    const long long col_items = (long long)B * ((NX / 2 + CB - 1) / CB);
    const int cap = PB_NUM_SMS * 6;
    const int grid_rows = (int)(row_items < cap ? row_items : cap);
    const int grid_cols = (int)(col_items < cap ? col_items : cap);
    {
        ProfScope prof(PROF_FFT_ROWS_FWD, stream);
        k_fft_rows_fwd<<<grid_rows, FFTD_THREADS, smem_rows, stream>>>(img, T.Z, kern, list, count, C, H, W, NX, NY,
                                                                       nb, T.planX, T.twX, T.slotX);
        PB_LAUNCH_CHECK("k_fft_rows_fwd");
    }
    {
        ProfScope prof(PROF_FFT_COLS, stream);
        k_fft_cols<<<grid_cols, FFTD_THREADS, smem_cols, stream>>>(T.Z, kern, list, count, C, NX, NY, CB, T.planY,
                                                                   T.twX, T.twY, T.freqY, T.slotY, a3, a2, a1, b0);
        PB_LAUNCH_CHECK("k_fft_cols");
    }
    {
        ProfScope prof(PROF_FFT_ROWS_INV, stream);
        k_fft_rows_inv<<<grid_rows, FFTD_THREADS, smem_rows, stream>>>(T.Z, out, kern, list, count, C, H, W, NX, NY,
                                                                       nb, T.planX, T.twX, T.slotX);
        PB_LAUNCH_CHECK("k_fft_rows_inv");
    }
    return PB_OK;
}

}  // namespace pb
