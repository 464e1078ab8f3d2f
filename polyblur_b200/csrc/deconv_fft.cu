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
#include <cstdlib>

#include <type_traits>

#include "fft2_static.cuh"
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
// Waits until the bulk stores have READ their shared-memory source (the buffer may be reused); the
// global writes need no further ordering inside this kernel (nobody re-reads Z before the next
// launch), and the plain wait_group would also flush the L1 (CCTL.IVALL) that caches the twiddles.
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

#ifndef FFTD_THREADS
#define FFTD_THREADS 256
#endif
// float2 per row pair in shared memory: the sequence + 4, so that the same slot of the nb row pairs
// of a CTA falls into different banks (the spectrum passes of P1 / P3 touch one slot of every pair)
__host__ __device__ inline int fftd_row_stride(int NX) { return NX + 4; }
// threads of the column kernel (288 = 9 warps would divide the butterfly counts of a 4-column block of
// 1152 evenly, but its 72-register budget spills: measured equal to 256)
#ifndef FFTC_THREADS
#define FFTC_THREADS 256
#endif
#ifndef PB_FFTD_MINB
#define PB_FFTD_MINB 3      // resident CTAs per SM (see estimate2.cu)
#endif
#ifndef PB_FFTC_MINB
#define PB_FFTC_MINB PB_FFTD_MINB
#endif

// extended coordinate (any torus of length >= n + 6 pad) -> source index, or -1 for the zero fill
//   n = image length, ext = 3 x kernel half-size (reach of the composite filter), and the source
//   geometry (n_in, off, pad) of SrcGeom: replicate pad on the fly, or an explicitly padded plane.
__device__ __forceinline__ int ext_src(int i, int n, int ext, int n_in, int off, int pad) {
    if (i >= n + 2 * ext) return -1;
    const int X = i - ext;                       // image coordinate of this extended sample
    if (pad > 0 && X >= 0 && X < n) return X + off;
    return geom_src(X, n_in, off, pad);
}

// ---------------------------------------------------------------------------------------------
// P1: rows forward.  Work item = (slot in the FFT class list, channel, block of nb row pairs).
// ---------------------------------------------------------------------------------------------
template <class SP>
__global__ void __launch_bounds__(FFTD_THREADS, PB_FFTD_MINB)
k_fft_rows_fwd(const float* __restrict__ img, float2* __restrict__ Z, const ImgKernel* __restrict__ kern,
               const int* __restrict__ list, const int* __restrict__ count, int C, int H, int W, int NX, int NY,
               int nb, Fft2Plan planX, const float2* __restrict__ twX, const int* __restrict__ slotX, SrcGeom G) {
    extern __shared__ __align__(16) float2 smf[];
    __shared__ int rowsrc[32];
    const int tid = threadIdx.x;
    const int blocks_per_plane = (NY / 2 + nb - 1) / nb;
    const int per_img = C * blocks_per_plane;
    const int total = count[0] * per_img;
    const int half = NX >> 1;
    const float inv_nb = 1.0f / (float)nb;
    const int RS = fftd_row_stride(NX);

    for (int w = blockIdx.x; w < total; w += gridDim.x) {
        const int slot = w / per_img;
        int r = w - slot * per_img;
        const int c = r / blocks_per_plane;
        const int rb = r - c * blocks_per_plane;
        const int im = list[slot];
        const int kpad = kern[im].ksize >> 1;
        const int pad = G.pad >= 0 ? G.pad : kpad;
        const int ext = 3 * kpad;
        const float* src = img + ((size_t)im * C + c) * (size_t)G.Hin * G.Win;
        const int Ws = G.Win;
        const int j0 = rb * 2 * nb;
        if (tid < 2 * nb) rowsrc[tid] = (j0 + tid < NY) ? ext_src(j0 + tid, H, ext, G.Hin, G.off, pad) : -1;
        __syncthreads();

        const int x_off = ext;
        if (((W | x_off | Ws | G.off) & 3) == 0) {
            // interior columns: image column x lives at extended column x + 3 pad; 128-bit loads of
            // both rows of a pair, four pairs of loads in flight per thread
            const int w4 = W >> 2;
            const float inv_w4 = 1.0f / (float)w4;
            const int tot = nb * w4;
            for (int base = tid; base < tot; base += 4 * FFTD_THREADS) {
                float4 va[4], vb[4];
                int off[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int idx = base + u * FFTD_THREADS;
                    va[u] = vb[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                    off[u] = -1;
                    if (idx < tot) {
                        const int p = fast_div(idx, w4, inv_w4);
                        const int x = (idx - p * w4) << 2;
                        const int sa = rowsrc[2 * p], sb = rowsrc[2 * p + 1];
                        if (sa >= 0) va[u] = PB_LD_STREAM(reinterpret_cast<const float4*>(src + (size_t)sa * Ws + x + G.off));
                        if (sb >= 0) vb[u] = PB_LD_STREAM(reinterpret_cast<const float4*>(src + (size_t)sb * Ws + x + G.off));
                        off[u] = p * RS + x_off + x;
                    }
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    if (off[u] >= 0) {
                        float4* d = reinterpret_cast<float4*>(smf + off[u]);
                        d[0] = make_float4(va[u].x, vb[u].x, va[u].y, vb[u].y);
                        d[1] = make_float4(va[u].z, vb[u].z, va[u].w, vb[u].w);
                    }
                }
            }
            // the 3 pad columns on the left and everything right of the image: torus map / zero fill
            const int nbord = NX - W;
            const float inv_nbord = 1.0f / (float)nbord;
            for (int idx = tid; idx < nb * nbord; idx += FFTD_THREADS) {
                const int p = fast_div(idx, nbord, inv_nbord);
                const int e = idx - p * nbord;
                const int i = e < x_off ? e : W + e;
                const int sx = ext_src(i, W, ext, G.Win, G.off, pad);
                float2 v = make_float2(0.f, 0.f);
                if (sx >= 0) {
                    const int sa = rowsrc[2 * p], sb = rowsrc[2 * p + 1];
                    if (sa >= 0) v.x = __ldg(src + (size_t)sa * Ws + sx);
                    if (sb >= 0) v.y = __ldg(src + (size_t)sb * Ws + sx);
                }
                smf[(size_t)p * RS + i] = v;
            }
        } else {
            const float inv_nx = 1.0f / (float)NX;
            for (int idx = tid; idx < nb * NX; idx += FFTD_THREADS) {
                const int p = fast_div(idx, NX, inv_nx);
                const int i = idx - p * NX;
                const int sx = ext_src(i, W, ext, G.Win, G.off, pad);
                float2 v = make_float2(0.f, 0.f);
                if (sx >= 0) {
                    const int sa = rowsrc[2 * p], sb = rowsrc[2 * p + 1];
                    if (sa >= 0) v.x = __ldg(src + (size_t)sa * Ws + sx);
                    if (sb >= 0) v.y = __ldg(src + (size_t)sb * Ws + sx);
                }
                smf[(size_t)p * RS + i] = v;
            }
        }
        __syncthreads();
        if constexpr (std::is_same<SP, NoStaticPlan>::value)
            fft2_forward_dif(smf, RS, nb, planX, twX, tid, FFTD_THREADS);
        else
            s_forward_dif<SP>(smf, RS, nb, twX, tid, FFTD_THREADS);
        // separate the two real rows: Xa[k] = (Z[k] + conj Z[-k]) / 2, Xb[k] = (Z[k] - conj Z[-k]) / (2i)
        float2* Zp = Z + ((size_t)slot * C + c) * half * NY;
        // four (kx, pair) items per trip: slot lookups first, then the scattered shared-memory reads
        for (int base = tid; base < nb * half; base += 4 * FFTD_THREADS) {
            int s1[4], s2[4], pp[4], kk[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int idx = base + u * FFTD_THREADS;
                kk[u] = -1;
                if (idx < nb * half) {
                    const int kx = fast_div(idx, nb, inv_nb);
                    const int p = idx - kx * nb;
                    if (j0 + 2 * p < NY) {
                        kk[u] = kx;
                        pp[u] = p;
                        s1[u] = __ldg(slotX + kx);
                        s2[u] = __ldg(slotX + (kx == 0 ? half : NX - kx));
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (kk[u] < 0) continue;
                const int ja = j0 + 2 * pp[u];
                const float2* row = smf + (size_t)pp[u] * RS;
                const float2 z1 = row[s1[u]], z2 = row[s2[u]];
                float2 xa, xb;
                if (kk[u] == 0) {                    // z1 = Z[0], z2 = Z[N/2]: both spectra real there
                    xa = make_float2(z1.x, z2.x);
                    xb = make_float2(z1.y, z2.y);
                } else {
                    xa = make_float2(0.5f * (z1.x + z2.x), 0.5f * (z1.y - z2.y));
                    xb = make_float2(0.5f * (z1.y + z2.y), 0.5f * (z2.x - z1.x));
                }
                float2* d = Zp + (size_t)kk[u] * NY + ja;
                if (ja + 1 < NY) {
                    PB_ST_STREAM(reinterpret_cast<float4*>(d), make_float4(xa.x, xa.y, xb.x, xb.y));
                } else {
                    d[0] = xa;
                }
            }
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------
// P2: columns.  Work item = (slot, block of CB columns); loops over the C planes of the image.
//   shared memory: data[CB][NY] float2 | Hs[CB][NY] float | Hn[NY], Hb[NY] float (block 0:
//                  Nyquist response, then the two mixing coefficients of column 0)
//                  | R[CB+1][13] float2 | mbarrier
//   Leaves r = DFT(swap(Y)) in Z: the inverse transform is swap(r), P3 swaps while loading.
// ---------------------------------------------------------------------------------------------
template <class SP>
__global__ void __launch_bounds__(FFTC_THREADS, PB_FFTC_MINB)
k_fft_cols(float2* __restrict__ Z, const ImgKernel* __restrict__ kern, const int* __restrict__ list,
           const int* __restrict__ count, int C, int NX, int NY, int CB, Fft2Plan planY,
           const float2* __restrict__ twX, const float2* __restrict__ stwY, const int* __restrict__ slotY,
           float a3, float a2, float a1, float b0) {
    extern __shared__ __align__(16) unsigned char smraw[];
    float2* data = reinterpret_cast<float2*>(smraw);
    float* Hs = reinterpret_cast<float*>(data + (size_t)CB * NY);
    float* Hn = Hs + (size_t)CB * NY;
    float* Hb = Hn + NY;
    float2* Rk = reinterpret_cast<float2*>(Hb + NY);                  // [(CB + 1)][13]
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
        // two threads per (col, dy): dx <= 0 and dx > 0, all loads of a thread issued together
        for (int base = 0; base < (CB + 1) * 13 * 2; base += FFTC_THREADS) {
            const int i2 = base + tid;
            const int idx = i2 >> 1, part = i2 & 1;
            const int col = idx / 13, dy = idx - col * 13;
            const int kx = (col < CB) ? kx0 + col : half;
            float2 acc = make_float2(0.f, 0.f);
            if (idx < (CB + 1) * 13 && (col < ncol || (col == CB && cb == 0))) {
                const int dx0 = part ? 1 : -PB_PAD;
                // t = kx * dx mod NX, stepped by kx (kx * NX < 2^31 for every supported length)
                unsigned t = ((unsigned)kx * (unsigned)(dx0 + NX)) % (unsigned)NX;
                const float* kr = &K->k[(dy + PB_PAD) * PB_KS + dx0 + PB_PAD];
                float kv[PB_PAD + 1];
                float2 e[PB_PAD + 1];
#pragma unroll
                for (int u = 0; u <= PB_PAD; ++u) {
                    const bool on = part ? (u < PB_PAD) : true;          // 13 taps below, 12 above
                    kv[u] = on ? __ldg(kr + u) : 0.f;
                    e[u] = on ? __ldg(twX + t) : make_float2(0.f, 0.f);
                    t += kx;
                    if (t >= (unsigned)NX) t -= NX;
                }
                if (dy == 0 && !part) kv[PB_PAD] -= 1.0f;     // spectrum of D = K - I (small where K^ ~ 1)
#pragma unroll
                for (int u = 0; u <= PB_PAD; ++u) {
                    acc.x = fmaf(kv[u], e[u].x, acc.x);
                    acc.y = fmaf(kv[u], e[u].y, acc.y);
                }
            }
            acc.x += __shfl_xor_sync(0xffffffffu, acc.x, 1);
            acc.y += __shfl_xor_sync(0xffffffffu, acc.y, 1);
            if (!part && idx < (CB + 1) * 13) Rk[idx] = acc;
        }
        // K^(ky, kx) = sum_dy R[dy] exp(-2 pi i ky dy / NY) is the length-NY DFT of the sparse
        // sequence r[dy mod NY] = R[dy]; it is real, so two columns share one complex transform
        // (r_A + i r_B -> K^_A + i K^_B), run by the same DIF core: the result arrives in slot order.
        const int ncolh = ncol + (cb == 0 ? 1 : 0);            // + the Nyquist column
        const int nseq = (ncolh + 1) >> 1;
        for (int idx = tid; idx < nseq * NY; idx += FFTC_THREADS) data[idx] = make_float2(0.f, 0.f);
        __syncthreads();
        for (int idx = tid; idx < nseq * PB_KS; idx += FFTC_THREADS) {
            const int q = idx / PB_KS, d = idx - q * PB_KS - PB_PAD;
            const int ca = 2 * q, cb2 = 2 * q + 1;
            const int ad = d < 0 ? -d : d;
            float2 ra = Rk[(ca < ncol ? ca : CB) * 13 + ad];
            float2 rb = make_float2(0.f, 0.f);
            if (cb2 < ncolh) rb = Rk[(cb2 < ncol ? cb2 : CB) * 13 + ad];
            if (d < 0) {
                ra.y = -ra.y;
                rb.y = -rb.y;
            }
            data[(size_t)q * NY + (d < 0 ? d + NY : d)] = make_float2(ra.x - rb.y, ra.y + rb.x);
        }
        __syncthreads();
        if constexpr (std::is_same<SP, NoStaticPlan>::value)
            fft2_forward_dif(data, NY, nseq, planY, stwY, tid, FFTC_THREADS);
        else
            s_forward_dif<SP>(data, NY, nseq, stwY, tid, FFTC_THREADS);
        // Hs[col][slot] = scale * P(K^)
        for (int idx = tid; idx < ncolh * NY; idx += FFTC_THREADS) {
            const int col = fast_div(idx, NY, inv_ny);
            const int s = idx - col * NY;
            const float2 z = data[(size_t)(col >> 1) * NY + s];
            const float kh = (col & 1) ? z.y : z.x;
            // a3..b0 hold c3, c2, c1, 1 of the D = K - I form: H = 1 + D^ (c1 + D^ (c2 + c3 D^))
            const float h = fmaf(fmaf(fmaf(a3, kh, a2), kh, a1), kh, b0) * scale;
            if (col == ncol) Hn[s] = h; else Hs[(size_t)col * NY + s] = h;
        }
        __syncthreads();
        if (cb == 0) {
            // column 0 carries DC (real part) and Nyquist (imaginary part) of two real spectra:
            // Y[k] = A Z[k] + B conj Z[-k], A = (H0 + Hn) / 2, B = (H0 - Hn) / 2; it is filtered by a
            // separate pass below, so its fused multiplier becomes 1
            for (int s = tid; s < NY; s += FFTC_THREADS) {
                const float h0 = Hs[s], hn = Hn[s];
                Hn[s] = 0.5f * (h0 + hn);
                Hb[s] = 0.5f * (h0 - hn);
                Hs[s] = 1.0f;
            }
            __syncthreads();
        }

        float2* Z0 = Z + ((size_t)slot * C * half + kx0) * NY;
        if (tid == 0) {
            fence_async_smem();
            mbar_expect_tx(bar, (uint32_t)((size_t)ncol * NY * sizeof(float2)));
            for (int col = 0; col < ncol; ++col)
                bulk_g2s(data + (size_t)col * NY, Z0 + (size_t)col * NY, (uint32_t)(NY * sizeof(float2)), bar);
        }
        for (int c = 0; c < C; ++c) {
            float2* Zc = Z0 + (size_t)c * half * NY;
            mbar_wait(bar, phase);
            phase ^= 1;
            if (cb == 0) {
                fft2_forward_dif(data, NY, ncol, planY, stwY, tid, FFTC_THREADS);
                // the pair {k, -k} of column 0 goes to one thread (in place, no hazard)
                for (int ky = tid; ky <= NY / 2; ky += FFTC_THREADS) {
                    const int s1 = __ldg(slotY + ky);
                    const int s2 = __ldg(slotY + (NY - ky) % NY);
                    const float2 z1 = data[s1], z2 = data[s2];
                    const float A1 = Hn[s1], B1 = Hb[s1], A2 = Hn[s2], B2 = Hb[s2];
                    data[s1] = make_float2(A1 * z1.x + B1 * z2.x, A1 * z1.y - B1 * z2.y);
                    if (s2 != s1) data[s2] = make_float2(A2 * z2.x + B2 * z1.x, A2 * z2.y - B2 * z1.y);
                }
                __syncthreads();
                // inverse-direction transform with the multiplication by H (and the re/im swap) folded
                // into its first stage
                fft2_forward_dit(data, NY, ncol, planY, stwY, tid, FFTC_THREADS, Hs, 2);
            } else {
                // forward, multiply by H (with the re/im swap), inverse: innermost stages fused in registers
                if constexpr (std::is_same<SP, NoStaticPlan>::value)
                    fft2_forward_mul_inverse(data, NY, ncol, planY, stwY, tid, FFTC_THREADS, Hs, 2);
                else
                    s_forward_mul_inverse<SP, 2>(data, NY, ncol, stwY, tid, FFTC_THREADS, Hs);
            }
            fence_async_smem();
            __syncthreads();
            if (tid == 0) {
                for (int col = 0; col < ncol; ++col)
                    bulk_s2g(Zc + (size_t)col * NY, data + (size_t)col * NY, (uint32_t)(NY * sizeof(float2)));
                bulk_commit();
                bulk_wait_all();
                // the buffer has been read out: fetch the next plane right away, the other threads
                // go straight to the mbarrier
                if (c + 1 < C) {
                    const float2* Zn = Zc + (size_t)half * NY;
                    mbar_expect_tx(bar, (uint32_t)((size_t)ncol * NY * sizeof(float2)));
                    for (int col = 0; col < ncol; ++col)
                        bulk_g2s(data + (size_t)col * NY, Zn + (size_t)col * NY, (uint32_t)(NY * sizeof(float2)), bar);
                }
            }
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------
// P3: rows inverse.  Work item = (slot, channel, block of nb row pairs that hold output rows).
// ---------------------------------------------------------------------------------------------
template <class SP>
__global__ void __launch_bounds__(FFTD_THREADS, PB_FFTD_MINB)
k_fft_rows_inv(const float2* __restrict__ Z, float* __restrict__ out, const ImgKernel* __restrict__ kern,
               const int* __restrict__ list, const int* __restrict__ count, int C, int H, int W, int NX, int NY,
               int nb, Fft2Plan planX, const float2* __restrict__ twX, const int* __restrict__ slotX,
               int clamp_out) {
    extern __shared__ __align__(16) float2 smf[];
    const int tid = threadIdx.x;
    const int blocks_per_plane = (NY / 2 + nb - 1) / nb;
    const int per_img = C * blocks_per_plane;
    const int total = count[0] * per_img;
    const size_t plane = (size_t)H * W;
    const int half = NX >> 1;
    const float inv_nb = 1.0f / (float)nb;
    const float lo = clamp_out ? 0.0f : -INFINITY, hi = clamp_out ? 1.0f : INFINITY;
    const int RS = fftd_row_stride(NX);

    for (int w = blockIdx.x; w < total; w += gridDim.x) {
        const int slot = w / per_img;
        int r = w - slot * per_img;
        const int c = r / blocks_per_plane;
        const int rb = r - c * blocks_per_plane;
        const int im = list[slot];
        const int ext = 3 * (kern[im].ksize >> 1);
        const int j0 = rb * 2 * nb;
        // rows of the extended image that are output rows: [ext, H + ext)
        if (j0 + 2 * nb <= ext || j0 >= H + ext) continue;
        const float2* Zp = Z + ((size_t)slot * C + c) * half * NY;
        // four (kx, pair) items per trip: their 128-bit spectrum loads and slot lookups are all in flight
        // before the first is used
        for (int base = tid; base < nb * half; base += 4 * FFTD_THREADS) {
            float4 v[4];
            int s1[4], s2[4], pp[4], kk[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int idx = base + u * FFTD_THREADS;
                v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                kk[u] = -1;
                if (idx < nb * half) {
                    const int kx = fast_div(idx, nb, inv_nb);
                    const int p = idx - kx * nb;
                    const int ja = j0 + 2 * p;
                    kk[u] = kx;
                    pp[u] = p;
                    // P2 leaves the spectra re/im swapped
                    if (ja + 1 < NY) {
                        v[u] = PB_LD_STREAM(reinterpret_cast<const float4*>(Zp + (size_t)kx * NY + ja));
                    } else if (ja < NY) {
                        const float2 t = __ldg(Zp + (size_t)kx * NY + ja);
                        v[u] = make_float4(t.x, t.y, 0.f, 0.f);
                    }
                    s1[u] = __ldg(slotX + kx);
                    s2[u] = __ldg(slotX + (kx == 0 ? half : NX - kx));
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (kk[u] < 0) continue;
                const float2 xa = make_float2(v[u].y, v[u].x), xb = make_float2(v[u].w, v[u].z);
                float2* row = smf + (size_t)pp[u] * RS;
                // Z[k] = Xa[k] + i Xb[k], Z[-k] = conj Xa[k] + i conj Xb[k]; stored swapped (re <-> im)
                if (kk[u] == 0) {
                    row[s1[u]] = make_float2(xb.x, xa.x);
                    row[s2[u]] = make_float2(xb.y, xa.y);
                } else {
                    row[s1[u]] = make_float2(xa.y + xb.x, xa.x - xb.y);
                    row[s2[u]] = make_float2(xb.x - xa.y, xa.x + xb.y);
                }
            }
        }
        __syncthreads();
        if constexpr (std::is_same<SP, NoStaticPlan>::value)
            fft2_forward_dit(smf, RS, nb, planX, twX, tid, FFTD_THREADS);
        else
            s_forward_dit<SP>(smf, RS, nb, twX, tid, FFTD_THREADS);
        // r = DFT(swap(Z)): row a = r.y, row b = r.x (the 1/(NX NY) scale is inside H)
        float* dst = out + ((size_t)im * C + c) * plane;
        const int x_off = ext;
        if (((W | x_off) & 3) == 0) {
            const int w4 = W >> 2;
            const float inv_w4 = 1.0f / (float)w4;
            for (int idx = tid; idx < nb * w4; idx += FFTD_THREADS) {
                const int p = fast_div(idx, w4, inv_w4);
                const int x = (idx - p * w4) << 2;
                const float4* sp = reinterpret_cast<const float4*>(smf + (size_t)p * RS + x_off + x);
                const float4 u0 = sp[0], u1 = sp[1];
                const int ya = j0 + 2 * p - ext;
                if (ya >= 0 && ya < H)
                    PB_ST_STREAM(reinterpret_cast<float4*>(dst + (size_t)ya * W + x),
                                 make_float4(fminf(fmaxf(u0.y, lo), hi), fminf(fmaxf(u0.w, lo), hi),
                                             fminf(fmaxf(u1.y, lo), hi), fminf(fmaxf(u1.w, lo), hi)));
                if (ya + 1 >= 0 && ya + 1 < H)
                    PB_ST_STREAM(reinterpret_cast<float4*>(dst + (size_t)(ya + 1) * W + x),
                                 make_float4(fminf(fmaxf(u0.x, lo), hi), fminf(fmaxf(u0.z, lo), hi),
                                             fminf(fmaxf(u1.x, lo), hi), fminf(fmaxf(u1.z, lo), hi)));
            }
        } else {
            const float inv_w = 1.0f / (float)W;
            for (int idx = tid; idx < nb * W; idx += FFTD_THREADS) {
                const int p = fast_div(idx, W, inv_w);
                const int x = idx - p * W;
                const float2 z = smf[(size_t)p * RS + x + x_off];
                const int ya = j0 + 2 * p - ext;
                if (ya >= 0 && ya < H) dst[(size_t)ya * W + x] = fminf(fmaxf(z.y, lo), hi);
                if (ya + 1 >= 0 && ya + 1 < H) dst[(size_t)(ya + 1) * W + x] = fminf(fmaxf(z.x, lo), hi);
            }
        }
        __syncthreads();
    }
}

// ---- host side ------------------------------------------------------------------------------

// Torus length for n samples: the even m in [n, 1.1 n] with the cheapest plan, cost = m x
// fft2_plan_cost -- three conflict-free stages beat four, and pure powers of two (whose last
// stage is 8- or 16-way bank conflicted) lose to their neighbours.
int fft_engine_length(int n) {
    Fft2Plan p;
    int best = 0;
    double best_cost = 1e300;
    const int hi = n + n / 10 + 16;
    for (int m = n + (n & 1); m <= hi; m += 2) {
        if (make_fft2_plan(m, &p) != 0) continue;
        const double cost = (double)m * fft2_plan_cost(p);
        if (cost < best_cost) {
            best_cost = cost;
            best = m;
        }
    }
    return best;
}

// tuning knobs (environment overrides are for experiments only)
static int env_int(const char* name, int dflt) {
    const char* v = getenv(name);
    return v ? atoi(v) : dflt;
}
static int rows_nb(int NX) {
    if (env_int("PB_FFT_NB", 0) > 0) return env_int("PB_FFT_NB", 0);
    int nb = (int)((64 * 1024) / ((size_t)NX * sizeof(float2)));
    if (nb < 1) nb = 1;
    if (nb > 8) nb = 8;
    return nb;
}
static int cols_cb(int NY) {
    if (env_int("PB_FFT_CB", 0) > 0) return env_int("PB_FFT_CB", 0);
    int cb = (int)((64 * 1024) / ((size_t)NY * 12));
    if (cb < 1) cb = 1;
    if (cb > 8) cb = 8;
    return cb;
}

bool fft_engine_supported(int H, int W, int pad) {
    const int NX = fft_engine_length(W + 6 * pad), NY = fft_engine_length(H + 6 * pad);
    if (NX <= 0 || NY <= 0) return false;
    const size_t lim = PB_SMEM_MAX - 4096;
    return (size_t)NX * sizeof(float2) <= lim && (size_t)NY * 12 + 8 * NY + 512 <= lim;
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
    l.off_stwX = take((size_t)l.NX * sizeof(float2));
    l.off_stwY = take((size_t)l.NY * sizeof(float2));
    l.off_slotX = take((size_t)l.NX * sizeof(int));
    l.off_slotY = take((size_t)l.NY * sizeof(int));
    l.off_freqY = take((size_t)l.NY * sizeof(int));
    l.off_Z = take((size_t)B * C * (l.NX / 2) * l.NY * sizeof(float2));
    l.total = o;
    if (L) *L = l;
    return o;
}

int fft_engine_prepare(char* base, const FftEngineLayout& L, FftEngineTables* T, TableJobs* jobs) {
    T->NX = L.NX;
    T->NY = L.NY;
    if (make_fft2_plan(L.NX, &T->planX) || make_fft2_plan(L.NY, &T->planY)) {
        set_error("no FFT plan for the %d x %d torus", L.NY, L.NX);
        return PB_ERR_UNSUPPORTED;
    }
    T->twX = reinterpret_cast<float2*>(base + L.off_twX);
    T->twY = reinterpret_cast<float2*>(base + L.off_twY);
    T->stwX = reinterpret_cast<float2*>(base + L.off_stwX);
    T->stwY = reinterpret_cast<float2*>(base + L.off_stwY);
    T->slotX = reinterpret_cast<int*>(base + L.off_slotX);
    T->slotY = reinterpret_cast<int*>(base + L.off_slotY);
    T->freqY = reinterpret_cast<int*>(base + L.off_freqY);
    T->Z = reinterpret_cast<float2*>(base + L.off_Z);
    jobs->add(TJ_TWIDDLES, L.NX, T->twX, nullptr, nullptr);
    jobs->add(TJ_TWIDDLES, L.NY, T->twY, nullptr, nullptr);
    jobs->add(TJ_STAGE_TW, T->planX.tw_total, T->stwX, nullptr, &T->planX);
    jobs->add(TJ_STAGE_TW, T->planY.tw_total, T->stwY, nullptr, &T->planY);
    jobs->add(TJ_PERM, L.NX, T->slotX, nullptr, &T->planX);
    jobs->add(TJ_PERM, L.NY, T->slotY, T->freqY, &T->planY);
    return PB_OK;
}

int launch_deconv_fft(const float* img, float* out, const ImgKernel* kern, const int* list, const int* count,
                      int B, int C, int H, int W, const FftEngineTables& T, float a3, float a2, float a1,
                      float b0, const SrcGeom& G, cudaStream_t stream) {
    const int NX = T.NX, NY = T.NY;
    const int nb = rows_nb(NX), CB = cols_cb(NY);
    const size_t smem_rows = (size_t)nb * fftd_row_stride(NX) * sizeof(float2);
    const size_t smem_cols = (size_t)CB * NY * 12 + (size_t)NY * 8 + (size_t)(CB + 1) * 13 * 8 + 64;
    const long long row_items = (long long)B * C * ((NY / 2 + nb - 1) / nb);
    const long long col_items = (long long)B * ((NX / 2 + CB - 1) / CB);
    const int cap = PB_NUM_SMS * 6;
    const int grid_rows = (int)(row_items < cap ? row_items : cap);
    const int grid_cols = (int)(col_items < cap ? col_items : cap);
    // compile-time plans for the standard tori (1080p: 2016 x 1152, 4K: 4000 x 2304), else the run-time core
#define PB_FFT_ROWS(SP)                                                                                          \
    do {                                                                                                         \
        PB_CUDA_TRY(cudaFuncSetAttribute(k_fft_rows_fwd<SP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_rows)); \
        PB_CUDA_TRY(cudaFuncSetAttribute(k_fft_rows_inv<SP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_rows)); \
        if (fwd) {                                                                                               \
            ProfScope prof(PROF_FFT_ROWS_FWD, stream);                                                           \
            k_fft_rows_fwd<SP><<<grid_rows, FFTD_THREADS, smem_rows, stream>>>(img, T.Z, kern, list, count, C, H, W, NX, \
                                                                               NY, nb, T.planX, T.stwX, T.slotX, G); \
        } else {                                                                                                 \
            ProfScope prof(PROF_FFT_ROWS_INV, stream);                                                           \
            k_fft_rows_inv<SP><<<grid_rows, FFTD_THREADS, smem_rows, stream>>>(T.Z, out, kern, list, count, C, H, W, NX, \
                                                                               NY, nb, T.planX, T.stwX, T.slotX, \
                                                                               G.clamp_out);                     \
        }                                                                                                        \
    } while (0)
#define PB_FFT_COLS(SP)                                                                                          \
    do {                                                                                                         \
        PB_CUDA_TRY(cudaFuncSetAttribute(k_fft_cols<SP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_cols)); \
        ProfScope prof(PROF_FFT_COLS, stream);                                                                   \
        k_fft_cols<SP><<<grid_cols, FFTC_THREADS, smem_cols, stream>>>(T.Z, kern, list, count, C, NX, NY, CB, T.planY, \
                                                                       T.twX, T.stwY, T.slotY, a3, a2, a1, b0);  \
    } while (0)
    for (int pass = 0; pass < 3; ++pass) {
        const bool fwd = pass == 0;
        if (pass == 1) {
            if (PlanY1152::matches(T.planY)) PB_FFT_COLS(PlanY1152);
            else if (PlanY2304::matches(T.planY)) PB_FFT_COLS(PlanY2304);
            else PB_FFT_COLS(NoStaticPlan);
        } else {
            if (PlanX2016::matches(T.planX)) PB_FFT_ROWS(PlanX2016);
            else if (PlanX4000::matches(T.planX)) PB_FFT_ROWS(PlanX4000);
            else PB_FFT_ROWS(NoStaticPlan);
        }
    }
#undef PB_FFT_ROWS
#undef PB_FFT_COLS
    PB_LAUNCH_CHECK("fft deconvolution passes");
    return PB_OK;
}

}  // namespace pb
