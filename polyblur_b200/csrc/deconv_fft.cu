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
#include <mutex>
#include <unordered_map>

#include <type_traits>

#include "fft2_static.cuh"
#include "kernels.cuh"

namespace pb {

// Timing experiment only (results are wrong): -DPB_ZMOD=n folds the half spectra of all images onto n slots, i.e. what the
// passes would cost if `Z` never left the L2 (profiles/r02_l2_probe.md).
#ifdef PB_ZMOD
#define PB_ZSLOT(s) ((s) % PB_ZMOD)
#else
#define PB_ZSLOT(s) (s)
#endif
static int env_int(const char* name, int dflt);

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
__device__ __forceinline__ void bulk_prefetch_l2(const void* src, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}
// 128-bit read-only load that asks L2 to fetch the whole 256-byte neighbourhood: the row passes read the spectrum in
// 64-byte pieces (4 row pairs of one column) and the same CTA comes back for the adjacent pieces with its next work items
__device__ __forceinline__ float4 ldg_f4_l2_256(const float4* p) {
    float4 v;
    asm volatile("ld.global.nc.L2::256B.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
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
// second-generation row kernels (64-bit shared-memory accesses only): the mirror-unit stage reads slot A R_last + q of
// every pair of the CTA for consecutive blocks A; with an even last radix the pairs must sit an odd number of float2 apart
__host__ __device__ constexpr int fftd_row_stride2(int NX, int RL) { return NX + ((RL & 1) ? 4 : 5); }
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
        float2* Zp = Z + ((size_t)PB_ZSLOT(slot) * C + c) * half * NY;
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
template <class SP, int THREADS = FFTC_THREADS>
__global__ void __launch_bounds__(THREADS, (THREADS > 256 ? 2 : PB_FFTC_MINB))
k_fft_cols(float2* __restrict__ Z, const ImgKernel* __restrict__ kern, const int* __restrict__ list,
           const int* __restrict__ count, int C, int NX, int NY, int CB, Fft2Plan planY,
           const float2* __restrict__ twX, const float2* __restrict__ stwY, const int* __restrict__ slotY,
           float a3, float a2, float a1, float b0, int mode) {
    // mode 0: every block of columns; 1: only the block that holds kx = 0 (the other columns are left to another
    // launch), one CTA per plane; 2: every block but that one -- Hn / Hb are then not needed and not allocated
    // (for a long column they are the difference between one and two resident CTAs)
    const int first_block_only = mode == 1;
    extern __shared__ __align__(16) unsigned char smraw[];
    float2* data = reinterpret_cast<float2*>(smraw);
    float* Hs = reinterpret_cast<float*>(data + (size_t)CB * NY);
    float* Hn = Hs + (size_t)CB * NY;
    float* Hb = Hn + NY;
    float2* Rk = reinterpret_cast<float2*>(mode == 2 ? Hn : Hb + NY);  // [(CB + 1)][13]
    uint64_t* bar = reinterpret_cast<uint64_t*>(Rk + (CB + 1) * 13 + 1);
    bar = reinterpret_cast<uint64_t*>(((uintptr_t)bar + 7) & ~(uintptr_t)7);
    const int tid = threadIdx.x;
    const int half = NX >> 1;
    // first_block_only: just the block that holds kx = 0 (the second-generation kernel does the other columns)
    const int blk0 = mode == 2 ? 1 : 0;
    const int nblk = first_block_only ? 1 : (half + CB - 1) / CB - blk0;
    const int total = count[0] * nblk;
    const float scale = 1.0f / ((float)NX * (float)NY);
    const float inv_ny = 1.0f / (float)NY;
    uint32_t phase = 0;
    if (tid == 0) mbar_init(bar, 1);
    __syncthreads();

    for (int w = blockIdx.x; w < total; w += gridDim.x) {
        const int slot = w / nblk;
        const int cb = w - slot * nblk + blk0;
        const int im = list[slot];
        const ImgKernel* K = kern + im;
        const int kx0 = cb * CB;
        const int ncol = min(CB, half - kx0);

        // R[col][dy] = sum_dx K[dy][dx] exp(-2 pi i kx dx / NX), dy = 0..12 (R[-dy] = conj R[dy]);
        // entry CB is the Nyquist column kx = NX / 2 (needed by the block that holds kx = 0)
        // two threads per (col, dy): dx <= 0 and dx > 0, all loads of a thread issued together
        for (int base = 0; base < (CB + 1) * 13 * 2; base += THREADS) {
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
        for (int idx = tid; idx < nseq * NY; idx += THREADS) data[idx] = make_float2(0.f, 0.f);
        __syncthreads();
        for (int idx = tid; idx < nseq * PB_KS; idx += THREADS) {
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
            fft2_forward_dif(data, NY, nseq, planY, stwY, tid, THREADS);
        else
            s_forward_dif<SP>(data, NY, nseq, stwY, tid, THREADS);
        // Hs[col][slot] = scale * P(K^)
        for (int idx = tid; idx < ncolh * NY; idx += THREADS) {
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
            for (int s = tid; s < NY; s += THREADS) {
                const float h0 = Hs[s], hn = Hn[s];
                Hn[s] = 0.5f * (h0 + hn);
                Hb[s] = 0.5f * (h0 - hn);
                Hs[s] = 1.0f;
            }
            __syncthreads();
        }

        float2* Z0 = Z + ((size_t)PB_ZSLOT(slot) * C * half + kx0) * NY;
        // first_block_only: one CTA per plane (blockIdx.y), so that the single column of every plane runs in parallel
        const int c_begin = first_block_only ? (int)blockIdx.y : 0, c_end = first_block_only ? (int)blockIdx.y + 1 : C;
        if (tid == 0) {
            fence_async_smem();
            mbar_expect_tx(bar, (uint32_t)((size_t)ncol * NY * sizeof(float2)));
            for (int col = 0; col < ncol; ++col)
                bulk_g2s(data + (size_t)col * NY, Z0 + ((size_t)c_begin * half + col) * NY, (uint32_t)(NY * sizeof(float2)), bar);
        }
        for (int c = c_begin; c < c_end; ++c) {
            float2* Zc = Z0 + (size_t)c * half * NY;
            mbar_wait(bar, phase);
            phase ^= 1;
            if (cb == 0) {
                fft2_forward_dif(data, NY, ncol, planY, stwY, tid, THREADS);
                // the pair {k, -k} of column 0 goes to one thread (in place, no hazard)
                for (int ky = tid; ky <= NY / 2; ky += THREADS) {
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
                fft2_forward_dit(data, NY, ncol, planY, stwY, tid, THREADS, Hs, 2);
            } else {
                // forward, multiply by H (with the re/im swap), inverse: innermost stages fused in registers
                if constexpr (std::is_same<SP, NoStaticPlan>::value)
                    fft2_forward_mul_inverse(data, NY, ncol, planY, stwY, tid, THREADS, Hs, 2);
                else
                    s_forward_mul_inverse<SP, 2>(data, NY, ncol, stwY, tid, THREADS, Hs);
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
                if (c + 1 < c_end) {
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
        const float2* Zp = Z + ((size_t)PB_ZSLOT(slot) * C + c) * half * NY;
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

// =============================================================================================
// Second-generation row passes (compile-time plans only): the first / last FFT stage works straight
// from / to global memory and the Hermitian separation (P1) / rebuild (P3) happens in the registers of
// the stage that touches the spectrum, so a row pair crosses shared memory once per inner stage only:
//   P1: global --(stage 0 in registers)--> smem --(inner stages)--> smem --(last stage + separation)--> Z
//   P3: Z --(rebuild + first inverse stage)--> smem --(inner stages)--> smem --(stage 0 + crop + clamp)--> out
// (first generation: load pass, ns stages, separation pass = ns + 2 round trips, and 50 % of the warp
// instructions of P1 / 47 % of P3 in the load / separation passes -- profiles/r02_sass_segments.md.)
// =============================================================================================

// shared-memory layout of the second-generation row kernels: padded stage-1 blocks for the three-stage plans
// (fft2_static.cuh: LayoutPad1), and the distance RS between the row pairs of a CTA -- the mirror-unit stage reads
// one block of R_last slots per (unit, pair): with an odd R_last consecutive units already fall into different
// bank pairs and the pairs sit 4 (mod 16) apart, with an even R_last the pairs must sit an odd distance apart
template <class SP, bool PADL>
struct Rows2Layout {
    using LY = typename std::conditional<(PADL && SP::ns == 3), PadFor<SP>, LayoutFlat>::type;
    static constexpr int LEN = SP::n + SP::R(0) * LY::PAD;
    static constexpr int WANT = (SP::R(SP::ns - 1) & 1) ? 4 : 5;
    static constexpr int RS = LEN + ((WANT - LEN) % 16 + 16) % 16;
};

// feeds stage 0 of P1: sample e = j + m M of row pair f = (extended rows 2f, 2f + 1) of one plane
struct RowPairSrc {
    const float* src;      // plane
    const int* rowsrc;     // [2 nb] source row of each extended row, -1 = zero row
    int Ws, off, W, ext, Win, pad;
    bool inner_interior;   // samples with 1 <= m <= R - 2 are all interior image columns
    template <int R, int M>
    __device__ __forceinline__ void load(int f, int j, float2 (&v)[R]) const {
        const int sa = rowsrc[2 * f], sb = rowsrc[2 * f + 1];
        if (inner_interior && sa >= 0 && sb >= 0) {
            // fast path (every pair but the ones that touch the zero fill): one base pointer per row, the inner
            // samples at immediate offsets, only the two outer ones go through the torus map
            const float* pa = src + (size_t)sa * Ws + (j - ext + off);
            const float* pb = src + (size_t)sb * Ws + (j - ext + off);
#pragma unroll
            for (int m = 1; m < R - 1; ++m) v[m] = make_float2(__ldg(pa + m * M), __ldg(pb + m * M));
            const int s0 = ext_src(j, W, ext, Win, off, pad);
            const int s1 = ext_src(j + (R - 1) * M, W, ext, Win, off, pad);
            const float* ra = src + (size_t)sa * Ws;
            const float* rb = src + (size_t)sb * Ws;
            v[0] = s0 >= 0 ? make_float2(__ldg(ra + s0), __ldg(rb + s0)) : make_float2(0.f, 0.f);
            v[R - 1] = s1 >= 0 ? make_float2(__ldg(ra + s1), __ldg(rb + s1)) : make_float2(0.f, 0.f);
            return;
        }
        const float* ra = src + (size_t)(sa < 0 ? 0 : sa) * Ws;
        const float* rb = src + (size_t)(sb < 0 ? 0 : sb) * Ws;
#pragma unroll
        for (int m = 0; m < R; ++m) {
            const int sx = ext_src(j + m * M, W, ext, Win, off, pad);
            float a = 0.f, b = 0.f;
            if (sx >= 0) {
                if (sa >= 0) a = __ldg(ra + sx);
                if (sb >= 0) b = __ldg(rb + sx);
            }
            v[m] = make_float2(a, b);
        }
    }
};

// drains stage 0 of P3: result e = j + m M of row pair f; r = DFT(swap(Z)): row a = r.y, row b = r.x
struct RowPairDst {
    float* dst;            // plane
    int W, H, ext, ya0;    // ya0 = image row of extended row 0 of this block (j0 - ext)
    float lo, hi;
    bool inner_interior;
    template <int R, int M>
    __device__ __forceinline__ void store(int f, int j, const float2 (&v)[R]) const {
        const int ya = ya0 + 2 * f;
        const bool oka = ya >= 0 && ya < H, okb = ya + 1 >= 0 && ya + 1 < H;
        if (inner_interior && oka && okb) {
            float* pa = dst + (size_t)ya * W + (j - ext);
            float* pb = pa + W;
#pragma unroll
            for (int m = 1; m < R - 1; ++m) {
                pa[m * M] = fminf(fmaxf(v[m].y, lo), hi);
                pb[m * M] = fminf(fmaxf(v[m].x, lo), hi);
            }
            if (j >= ext) {
                pa[0] = fminf(fmaxf(v[0].y, lo), hi);
                pb[0] = fminf(fmaxf(v[0].x, lo), hi);
            }
            if (j + (R - 1) * M - ext < W) {
                pa[(R - 1) * M] = fminf(fmaxf(v[R - 1].y, lo), hi);
                pb[(R - 1) * M] = fminf(fmaxf(v[R - 1].x, lo), hi);
            }
            return;
        }
        float* pa = dst + (size_t)(oka ? ya : 0) * W;
        float* pb = dst + (size_t)(okb ? ya + 1 : 0) * W;
#pragma unroll
        for (int m = 0; m < R; ++m) {
            const int X = j + m * M - ext;
            if (X >= 0 && X < W) {
                if (oka) pa[X] = fminf(fmaxf(v[m].y, lo), hi);
                if (okb) pb[X] = fminf(fmaxf(v[m].x, lo), hi);
            }
        }
    }
};

// Which of {k, NX - k} is the half-spectrum column of slot q of a mirror unit's block A (k = fA + KS q, fA < KS):
// known at compile time for every q but the one whose range straddles NX / 2.
template <int NX, int KS>
__device__ __forceinline__ bool unit_lower_half(int q, int kA) {
    if (2 * KS * (q + 1) <= NX) return true;
    if (2 * KS * q > NX) return false;
    return 2 * kA < NX;
}
template <int NX, int KS>
__device__ __forceinline__ bool unit_maybe_nyquist(int q) { return 2 * KS * q <= NX && NX < 2 * KS * (q + 1); }

#ifndef PB_ROWS_CHUNK
#define PB_ROWS_CHUNK 1      // measured: 4 consecutive blocks per CTA is 8 % slower (1.48 / 1.63 ms against 1.38 / 1.50 ms per step)
#endif

template <class SP, int NY, int THREADS, bool PADL>
__global__ void __launch_bounds__(THREADS, (THREADS > 256 || SP::R(0) > 20 ? 2 : PB_FFTD_MINB))
k_fft_rows_fwd2(const float* __restrict__ img, float2* __restrict__ Z, const ImgKernel* __restrict__ kern,
                const int* __restrict__ list, const int* __restrict__ count, int C, int H, int W, int nb,
                const float2* __restrict__ twX, const int4* __restrict__ units, SrcGeom G) {
    constexpr int NX = SP::n, NS = SP::ns, R0 = SP::R(0), M0 = NX / R0, RL = SP::R(NS - 1), KS = NX / RL;
    extern __shared__ __align__(16) float2 smf[];
    __shared__ int rowsrc[32];
    const int tid = threadIdx.x;
    const int blocks_per_plane = (NY / 2 + nb - 1) / nb;
    const int per_img = C * blocks_per_plane;
    const int total = count[0] * per_img;
    constexpr int half = NX >> 1;
    using LY = typename Rows2Layout<SP, PADL>::LY;
    constexpr int RS = Rows2Layout<SP, PADL>::RS;
    const int nunits = units[0].x;

    // a CTA takes PB_ROWS_CHUNK consecutive row blocks of a plane (they share 256-byte pieces of the spectrum columns)
    for (int w0 = blockIdx.x * PB_ROWS_CHUNK; w0 < total; w0 += gridDim.x * PB_ROWS_CHUNK)
    for (int w = w0; w < min(w0 + PB_ROWS_CHUNK, total); ++w) {
        const int slot = w / per_img;
        int r = w - slot * per_img;
        const int c = r / blocks_per_plane;
        const int rb = r - c * blocks_per_plane;
        const int im = list[slot];
        const int kpad = kern[im].ksize >> 1;
        const int pad = G.pad >= 0 ? G.pad : kpad;
        const int ext = 3 * kpad;
        const int j0 = rb * 2 * nb;
        if (tid < 2 * nb) rowsrc[tid] = (j0 + tid < NY) ? ext_src(j0 + tid, H, ext, G.Hin, G.off, pad) : -1;
        __syncthreads();
        RowPairSrc S;
        S.src = img + ((size_t)im * C + c) * (size_t)G.Hin * G.Win;
        S.rowsrc = rowsrc;
        S.Ws = G.Win;
        S.off = G.off;
        S.W = W;
        S.ext = ext;
        S.Win = G.Win;
        S.pad = pad;
        S.inner_interior = (M0 >= ext) && ((R0 - 1) * M0 - ext <= W);
#ifdef PB_ROWS_PREFETCH        // measured slower (+4 % on both row passes): off
        {
            // the image rows of this CTA's next work item on their way into L2 (128-byte lines)
            const int wn = w + gridDim.x * PB_ROWS_CHUNK;
            if (wn < total) {
                const int slotn = wn / per_img;
                const int rn = wn - slotn * per_img;
                const int cn = rn / blocks_per_plane;
                const int jn = (rn - cn * blocks_per_plane) * 2 * nb;
                const float* srcn = img + ((size_t)list[slotn] * C + cn) * (size_t)G.Hin * G.Win;
                const int lines = (G.Win + 31) / 32;
                for (int i = tid; i < 2 * nb * lines; i += THREADS) {
                    const int rr = i / lines, ln = i - rr * lines;
                    const int sr = (jn + rr < NY) ? ext_src(jn + rr, H, ext, G.Hin, G.off, pad) : -1;
                    if (sr >= 0) asm volatile("prefetch.global.L2 [%0];" ::"l"(srcn + (size_t)sr * G.Win + ln * 32));
                }
            }
        }
#endif
        s_dif_first<R0, M0, NX, RowPairSrc, LY>(smf, RS, nb, twX + SP::tw_off(0), tid, THREADS, S);
        __syncthreads();
        SDifRun<SP, 1, NS - 2, false, LY>::run(smf, RS, nb, twX, tid, THREADS);
        // last stage (M = 1) on the two blocks of a mirror unit, then
        //   Xa[k] = (Z[k] + conj Z[-k]) / 2,  Xb[k] = (Z[k] - conj Z[-k]) / (2i)   ->  Z[plane][kx][row pair]
        float2* Zp = Z + ((size_t)PB_ZSLOT(slot) * C + c) * half * NY;
        for (int idx = tid; idx < nunits * nb; idx += THREADS) {
            const int u = idx / nb, p = idx - u * nb;
            const int ja = j0 + 2 * p;
            if (ja >= NY) continue;
            const int4 U = __ldg(units + 1 + u);
            const float2* row = smf + (size_t)p * RS;
            float2 va[RL], vb[RL];
#pragma unroll
            for (int q = 0; q < RL; ++q) va[q] = row[LY::off(U.x * RL) + q];
            Dft<RL>::run(va);
            if (U.x == 0 || U.x == U.y) {
                // the two self-mirror blocks (block 0; the block that holds NX / 2): rare, generic code
#pragma unroll
                for (int q = 0; q < RL; ++q) {
                    const int kA = U.z + KS * q;
                    const float2 z1 = va[q];
                    const float2 z2 = (U.x == 0) ? va[(RL - q) % RL] : va[RL - 1 - q];
                    if (kA == 0 || 2 * kA == NX) {
                        // DC and Nyquist are real for both rows and share column 0: (Xa, Xb) = ((dc, ny), (dc, ny))
                        float* d = reinterpret_cast<float*>(Zp + ja) + (kA == 0 ? 0 : 1);
                        d[0] = z1.x;
                        if (ja + 1 < NY) d[2] = z1.y;
                    } else if (2 * kA < NX) {
                        const float2 xa = make_float2(0.5f * (z1.x + z2.x), 0.5f * (z1.y - z2.y));
                        const float2 xb = make_float2(0.5f * (z1.y + z2.y), 0.5f * (z2.x - z1.x));
                        float2* d = Zp + (size_t)kA * NY + ja;
                        if (ja + 1 < NY) *reinterpret_cast<float4*>(d) = make_float4(xa.x, xa.y, xb.x, xb.y);
                        else d[0] = xa;
                    }
                }
                continue;
            }
#pragma unroll
            for (int q = 0; q < RL; ++q) vb[q] = row[LY::off(U.y * RL) + q];
            Dft<RL>::run(vb);
            // slot q of block A holds k = fA + KS q, slot RL - 1 - q of block B its mirror NX - k; the lower of the two
            // is the half-spectrum column.  Columns of one unit are KS apart: two base pointers, immediate offsets.
            float2* dlo = Zp + (size_t)U.z * NY + ja;                              // column fA (+ KS q)
            float2* dhi = Zp + (size_t)(NX - U.z - KS * (RL - 1)) * NY + ja;       // column NX - fA - KS (RL - 1) (+ KS (RL - 1 - q))
            const bool pairrow = (NY & 1) == 0 || ja + 1 < NY;
#pragma unroll
            for (int q = 0; q < RL; ++q) {
                const int kA = U.z + KS * q;
                const bool lower = unit_lower_half<NX, KS>(q, kA);
                const float2 zk = lower ? va[q] : vb[RL - 1 - q];
                const float2 zm = lower ? vb[RL - 1 - q] : va[q];
                const float2 xa = make_float2(0.5f * (zk.x + zm.x), 0.5f * (zk.y - zm.y));
                const float2 xb = make_float2(0.5f * (zk.y + zm.y), 0.5f * (zm.x - zk.x));
                float2* d = lower ? dlo + (size_t)(KS * q) * NY : dhi + (size_t)(KS * (RL - 1 - q)) * NY;
                if (pairrow) *reinterpret_cast<float4*>(d) = make_float4(xa.x, xa.y, xb.x, xb.y);
                else d[0] = xa;
            }
        }
        __syncthreads();
    }
}

template <class SP, int NY, int THREADS, bool PADL>
__global__ void __launch_bounds__(THREADS, (THREADS > 256 || SP::R(0) > 20 ? 2 : PB_FFTD_MINB))
k_fft_rows_inv2(const float2* __restrict__ Z, float* __restrict__ out, const ImgKernel* __restrict__ kern,
                const int* __restrict__ list, const int* __restrict__ count, int C, int H, int W, int nb,
                const float2* __restrict__ twX, const int4* __restrict__ units, int clamp_out) {
    constexpr int NX = SP::n, NS = SP::ns, R0 = SP::R(0), M0 = NX / R0, RL = SP::R(NS - 1), KS = NX / RL;
    extern __shared__ __align__(16) float2 smf[];
    const int tid = threadIdx.x;
    const int blocks_per_plane = (NY / 2 + nb - 1) / nb;
    const int per_img = C * blocks_per_plane;
    const int total = count[0] * per_img;
    const size_t plane = (size_t)H * W;
    constexpr int half = NX >> 1;
    using LY = typename Rows2Layout<SP, PADL>::LY;
    constexpr int RS = Rows2Layout<SP, PADL>::RS;
    const int nunits = units[0].x;

    // a CTA takes PB_ROWS_CHUNK consecutive row blocks of a plane (they share 256-byte pieces of the spectrum columns)
    for (int w0 = blockIdx.x * PB_ROWS_CHUNK; w0 < total; w0 += gridDim.x * PB_ROWS_CHUNK)
    for (int w = w0; w < min(w0 + PB_ROWS_CHUNK, total); ++w) {
        const int slot = w / per_img;
        int r = w - slot * per_img;
        const int c = r / blocks_per_plane;
        const int rb = r - c * blocks_per_plane;
        const int im = list[slot];
        const int ext = 3 * (kern[im].ksize >> 1);
        const int j0 = rb * 2 * nb;
        // rows of the extended image that are output rows: [ext, H + ext)
        if (j0 + 2 * nb <= ext || j0 >= H + ext) continue;
        const float2* Zp = Z + ((size_t)PB_ZSLOT(slot) * C + c) * half * NY;
#ifdef PB_ROWS_PREFETCH        // measured slower (+4 % on both row passes): off
        {
            // this CTA's next work item reads the same columns a few rows further: start its 64-byte pieces (one
            // per half-spectrum column) on their way into L2 now, the loads below then cost an L2 hit, not DRAM
            const int wn = w + gridDim.x * PB_ROWS_CHUNK;
            if (wn < total) {
                const int slotn = wn / per_img;
                const int rn = wn - slotn * per_img;
                const int cn = rn / blocks_per_plane;
                const float2* Zn = Z + ((size_t)PB_ZSLOT(slotn) * C + cn) * half * NY + (size_t)(rn - cn * blocks_per_plane) * 2 * nb;
                for (int kx = tid; kx < half; kx += THREADS)
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(Zn + (size_t)kx * NY));
            }
        }
#endif
        // rebuild Z[k] = Xa[k] + i Xb[k], Z[-k] = conj Xa[k] + i conj Xb[k] of a mirror unit in registers (P2 leaves
        // the spectra re/im swapped, and the inverse-by-forward trick wants them swapped), first inverse stage (M = 1)
        for (int idx = tid; idx < nunits * nb; idx += THREADS) {
            const int u = idx / nb, p = idx - u * nb;
            const int ja = j0 + 2 * p;
            if (ja >= NY) continue;        // pair beyond the torus (never an output row; its slots stay unused)
            const int4 U = __ldg(units + 1 + u);
            float2* row = smf + (size_t)p * RS;
            const bool pairrow = (NY & 1) == 0 || ja + 1 < NY;
            float2 va[RL], vb[RL];
            if (U.x == 0 || U.x == U.y) {
                // the two self-mirror blocks: rare, generic code
#pragma unroll
                for (int q = 0; q < RL; ++q) {
                    const int kA = U.z + KS * q;
                    const int kx = (2 * kA <= NX) ? kA : NX - kA;
                    const float2* s = Zp + (size_t)(kx == half ? 0 : kx) * NY + ja;
                    float4 ld = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (ja + 1 < NY) ld = __ldg(reinterpret_cast<const float4*>(s));
                    else if (ja < NY) {
                        const float2 t = __ldg(s);
                        ld = make_float4(t.x, t.y, 0.f, 0.f);
                    }
                    const float2 xa = make_float2(ld.y, ld.x), xb = make_float2(ld.w, ld.z);
                    float2 zk, zm;              // swapped Z[kx], Z[-kx]
                    if (kA == 0) {
                        zk = zm = make_float2(xb.x, xa.x);
                    } else if (2 * kA == NX) {
                        zk = zm = make_float2(xb.y, xa.y);
                    } else {
                        zk = make_float2(xa.y + xb.x, xa.x - xb.y);
                        zm = make_float2(xb.x - xa.y, xa.x + xb.y);
                    }
                    va[q] = (2 * kA <= NX) ? zk : zm;
                }
                Dft<RL>::run(va);
#pragma unroll
                for (int q = 0; q < RL; ++q) row[LY::off(U.x * RL) + q] = va[q];
                continue;
            }
            const float2* slo = Zp + (size_t)U.z * NY + ja;
            const float2* shi = Zp + (size_t)(NX - U.z - KS * (RL - 1)) * NY + ja;
            float4 ld[RL];
#pragma unroll
            for (int q = 0; q < RL; ++q) {
                const int kA = U.z + KS * q;
                const bool lower = unit_lower_half<NX, KS>(q, kA);
                const float2* s = lower ? slo + (size_t)(KS * q) * NY : shi + (size_t)(KS * (RL - 1 - q)) * NY;
#ifdef PB_ROWS_L2HINT
                if (pairrow) ld[q] = ldg_f4_l2_256(reinterpret_cast<const float4*>(s));
#else
                if (pairrow) ld[q] = __ldg(reinterpret_cast<const float4*>(s));
#endif
                else {
                    const float2 t = ja < NY ? __ldg(s) : make_float2(0.f, 0.f);
                    ld[q] = make_float4(t.x, t.y, 0.f, 0.f);
                }
            }
#pragma unroll
            for (int q = 0; q < RL; ++q) {
                const int kA = U.z + KS * q;
                const bool lower = unit_lower_half<NX, KS>(q, kA);
                const float2 xa = make_float2(ld[q].y, ld[q].x), xb = make_float2(ld[q].w, ld[q].z);
                const float2 zk = make_float2(xa.y + xb.x, xa.x - xb.y);       // swapped Z[kx]
                const float2 zm = make_float2(xb.x - xa.y, xa.x + xb.y);       // swapped Z[-kx]
                va[q] = lower ? zk : zm;
                vb[RL - 1 - q] = lower ? zm : zk;
            }
            Dft<RL>::run(va);
#pragma unroll
            for (int q = 0; q < RL; ++q) row[LY::off(U.x * RL) + q] = va[q];
            Dft<RL>::run(vb);
#pragma unroll
            for (int q = 0; q < RL; ++q) row[LY::off(U.y * RL) + q] = vb[q];
        }
        __syncthreads();
        SDitRun<SP, NS - 2, NS - 2, false, LY>::run(smf, RS, nb, twX, tid, THREADS);
        RowPairDst D;
        D.dst = out + ((size_t)im * C + c) * plane;
        D.W = W;
        D.H = H;
        D.ext = ext;
        D.ya0 = j0 - ext;
        D.lo = clamp_out ? 0.0f : -INFINITY;
        D.hi = clamp_out ? 1.0f : INFINITY;
        D.inner_interior = (M0 >= ext) && ((R0 - 1) * M0 - ext <= W);
        s_dit_last<R0, M0, NX, RowPairDst, LY>(smf, RS, nb, twX + SP::tw_off(0), tid, THREADS, D);
        __syncthreads();
    }
}

// =============================================================================================
// Second-generation column pass: two stages NY = RA x RB with 32-36-point butterflies in registers.
//   stage A  (DIF, radix RA, stride RB)   reads the column as the bulk copy delivered it, writes RA blocks of RB
//                                         slots, each block padded to RB + 1 (conflict-free for the next stage)
//   middle   (radix RB on one block)      DFT -> x H (real, with the re/im swap of the inverse) -> DFT, in place
//   stage A' (DIT, radix RA)              reads the padded blocks, writes the column for the bulk store
// Three shared-memory round trips per plane instead of five, two of six twiddle passes.  The layouts on the two
// sides of stage A differ, so every thread holds its whole butterfly across a barrier (one butterfly per thread:
// CB x RB <= threads).  Column kx = 0 (DC + Nyquist of the row transform, which mixes k and -k) is left to the
// first-generation kernel, launched for that one column.
// =============================================================================================
#ifndef FFTC2_THREADS
#define FFTC2_THREADS 256
#endif
// floats per block of H: = 4 (mod 8), so that the LDS.128 of a quarter warp (consecutive blocks) fall into different banks
#define FFTC2_HSTRIDE(RB) ((RB) + ((4 - ((RB) & 7)) & 7))
// float2 per padded column: RA blocks of RB + 1, rounded up to = 4 (mod 16): 16-byte aligned for the bulk copies, and
// the butterflies of a warp that straddles two columns in stage A stay on different banks
#define FFTC2_CSTRIDE(RA, RB) ((RA) * ((RB) + 1) + ((4 - (((RA) * ((RB) + 1)) & 15)) & 15))

// DIF stage A of k_fft_cols2 on natural-order data -> padded blocks: all loads, barrier, then the stores
// (the two layouts overlap in shared memory).  Thread = butterfly bj of column bc; idle when bc >= nseq.
// WARPCOL (RB == 32: a warp is a column): only this warp reads and writes the column, so the barrier between the loads and
// the stores is a warp barrier (an idle warp reads a duplicate column that its owner may be rewriting: never stored).
template <int RA, int RB, bool WARPCOL = false>
__device__ __forceinline__ void cols2_stage_a(float2* data, const float2* stwA, int bc, int bj, int nseq) {
    constexpr int BS = RB + 1, CS = FFTC2_CSTRIDE(RA, RB);
    // every thread loads and transforms (idle ones a duplicate of the last live column: values that are defined on
    // one side of the barrier only make ptxas spill the whole butterfly around it); only the stores are predicated
    float2 v[RA];
    const bool on = bc < nseq;
    if constexpr (WARPCOL) {
        // `on` is uniform over the warp: an idle warp skips the stage (and touches no column that its owner is rewriting)
        if (on) {
            float2* p = data + (size_t)bc * CS + bj;
#pragma unroll
            for (int m = 0; m < RA; ++m) v[m] = p[m * RB];
            __syncwarp();
            Dft<RA>::run(v);
            p[0] = v[0];
#pragma unroll
            for (int q = 1; q < RA; ++q) p[q * BS] = c_mul(v[q], stwA[(q - 1) * RB + bj]);
        }
        __syncthreads();
        return;
    }
    float2* p = data + (size_t)(on ? bc : nseq - 1) * CS + bj;
#pragma unroll
    for (int m = 0; m < RA; ++m) v[m] = p[m * RB];
    __syncthreads();
    Dft<RA>::run(v);
    if (on) p[0] = v[0];
#pragma unroll
    for (int q = 1; q < RA; ++q) {
        const float2 t = c_mul(v[q], stwA[(q - 1) * RB + bj]);
        if (on) p[q * BS] = t;
    }
    __syncthreads();
}

// PC (per-column hand-off, RB == 32): every column has its own mbarrier and its warp stores it, waits for the read-out and
// fetches the same column of the next plane on its own -- no CTA barrier between stage A' of one plane and stage A of the next.
template <int RA, int RB, bool PC = false>
__global__ void __launch_bounds__(FFTC2_THREADS, 2)
k_fft_cols2(float2* __restrict__ Z, const ImgKernel* __restrict__ kern, const int* __restrict__ list,
            const int* __restrict__ count, int C, int NX, int CB, const float2* __restrict__ twX,
            const float2* __restrict__ stwA, float a3, float a2, float a1, float b0, int rev) {
    constexpr int NY = RA * RB, BS = RB + 1, CS = FFTC2_CSTRIDE(RA, RB);
    constexpr int HS = FFTC2_HSTRIDE(RB), HC = RA * HS;
    extern __shared__ __align__(16) unsigned char smraw[];
    float2* data = reinterpret_cast<float2*>(smraw);                         // [CB][CS]
    float* Hs = reinterpret_cast<float*>(data + (size_t)CB * CS);            // [CB][RA][HS]
    float2* Rk = reinterpret_cast<float2*>(Hs + (size_t)CB * HC);            // [CB][13]
    float2* tws = Rk + CB * 13;                                              // [(RA - 1) RB] stage-A twiddles
    uint64_t* bar = reinterpret_cast<uint64_t*>(tws + (RA - 1) * RB + 1);
    bar = reinterpret_cast<uint64_t*>(((uintptr_t)bar + 7) & ~(uintptr_t)7);
    const int tid = threadIdx.x;
    const int half = NX >> 1;
    const int nblk = (half - 1 + CB - 1) / CB;          // blocks over columns 1 .. half - 1
    const int total = count[0] * nblk;
    const float scale = 1.0f / ((float)NX * (float)NY);
    static_assert(!PC || RB == 32, "per-column hand-off needs one warp per column");
    uint32_t phase = 0;                                  // PC: of this warp's own barrier bar[bc]
    if (tid == 0) {
        mbar_init(bar, 1);
        if (PC)
            for (int i = 1; i < FFTC2_THREADS / RB; ++i) mbar_init(bar + i, 1);
    }
    // the stage twiddles stay in shared memory for the life of the (persistent) CTA: with ~20 KB of L1 left beside
    // the 2 x 110 KB of shared memory, the table reads of stage A / A' kept missing it (long-scoreboard stalls)
    for (int i = tid; i < (RA - 1) * RB; i += FFTC2_THREADS) tws[i] = __ldg(stwA + i);
    __syncthreads();
    // this thread's butterfly of stage A / A': column bc, position bj (idle when bc >= ncol)
    const int bc = tid / RB, bj = tid - bc * RB;

    // rev: the work items are taken last first -- P1 wrote the spectrum forwards, so its last ~100 MB are still in the L2
    // when this pass starts with them, and this pass ends at the first image, where P3 (forwards) begins
    for (int w0 = blockIdx.x; w0 < total; w0 += gridDim.x) {
        const int w = rev ? total - 1 - w0 : w0;
        const int slot = w / nblk;
        const int cb = w - slot * nblk;
        const int im = list[slot];
        const ImgKernel* K = kern + im;
        const int kx0 = 1 + cb * CB;
        const int ncol = min(CB, half - kx0);

        // R[col][dy] = sum_dx K[dy][dx] exp(-2 pi i kx dx / NX), dy = 0..12 (R[-dy] = conj R[dy]: point-symmetric taps);
        // two threads per (col, dy): dx <= 0 and dx > 0
        for (int base = 0; base < CB * 13 * 2; base += FFTC2_THREADS) {
            const int i2 = base + tid;
            const int idx = i2 >> 1, part = i2 & 1;
            const int col = idx / 13, dy = idx - col * 13;
            float2 acc = make_float2(0.f, 0.f);
            if (idx < CB * 13 && col < ncol) {
                const int kx = kx0 + col;
                const int dx0 = part ? 1 : -PB_PAD;
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
            if (!part && idx < CB * 13) Rk[idx] = acc;
        }
        // K^(ky, kx) = length-NY DFT of the sparse sequence r[dy mod NY] = R[dy]; real, so two columns share one
        // complex transform, run through the same two stages: the result arrives in the slot order of the data
        const int nseq = (ncol + 1) >> 1;
        for (int idx = tid; idx < nseq * CS; idx += FFTC2_THREADS) {
            const int q = idx / CS;
            data[(size_t)q * CS + (idx - q * CS)] = make_float2(0.f, 0.f);
        }
        __syncthreads();
        for (int idx = tid; idx < nseq * PB_KS; idx += FFTC2_THREADS) {
            const int q = idx / PB_KS, d = idx - q * PB_KS - PB_PAD;
            const int ca = 2 * q, cb2 = 2 * q + 1;
            const int ad = d < 0 ? -d : d;
            float2 ra = Rk[ca * 13 + ad];
            float2 rb = make_float2(0.f, 0.f);
            if (cb2 < ncol) rb = Rk[cb2 * 13 + ad];
            if (d < 0) {
                ra.y = -ra.y;
                rb.y = -rb.y;
            }
            data[(size_t)q * CS + (d < 0 ? d + NY : d)] = make_float2(ra.x - rb.y, ra.y + rb.x);
        }
        __syncthreads();
        cols2_stage_a<RA, RB, PC>(data, tws, bc, bj, nseq);
        // last DIF stage of the kernel spectra on each padded block, then Hs[col][block][q] = scale * P(K^)
        for (int idx = tid; idx < nseq * RA; idx += FFTC2_THREADS) {
            const int f = idx / RA, blk = idx - f * RA;
            float2* p = data + (size_t)f * CS + blk * BS;
            float2 v[RB];
#pragma unroll
            for (int m = 0; m < RB; ++m) v[m] = p[m];
            Dft<RB>::run(v);
            float* h0 = Hs + (size_t)(2 * f) * HC + blk * HS;
            const bool two = 2 * f + 1 < ncol;
#pragma unroll
            for (int m = 0; m < RB; ++m) {
                // a3..b0 hold c3, c2, c1, 1 of the D = K - I form: H = 1 + D^ (c1 + D^ (c2 + c3 D^))
                h0[m] = fmaf(fmaf(fmaf(a3, v[m].x, a2), v[m].x, a1), v[m].x, b0) * scale;
                if (two) h0[HC + m] = fmaf(fmaf(fmaf(a3, v[m].y, a2), v[m].y, a1), v[m].y, b0) * scale;
            }
        }
        __syncthreads();

        float2* Z0 = Z + ((size_t)PB_ZSLOT(slot) * C * half + kx0) * NY;
        if (tid == 0) {
            fence_async_smem();
            if (!PC) mbar_expect_tx(bar, (uint32_t)((size_t)ncol * NY * sizeof(float2)));
            for (int col = 0; col < ncol; ++col) {
                if (PC) mbar_expect_tx(bar + col, (uint32_t)(NY * sizeof(float2)));
                bulk_g2s(data + (size_t)col * CS, Z0 + (size_t)col * NY, (uint32_t)(NY * sizeof(float2)), PC ? bar + col : bar);
            }
        }
        for (int c = 0; c < C; ++c) {
            float2* Zc = Z0 + (size_t)c * half * NY;
            if (!PC) {
                mbar_wait(bar, phase);
                phase ^= 1;
            } else if (bc < ncol) {
                mbar_wait(bar + bc, phase);
                phase ^= 1;
            }
            if (tid == 32) {
                // the next plane's columns (or the first plane of this CTA's next work item) on their way into L2
                // while this one is transformed: the bulk load, issued once the buffer is free again, finds them there
                const float2* Zn = nullptr;
                int ncn = ncol;
                if (c + 1 < C) {
                    Zn = Zc + (size_t)half * NY;
                } else if (w0 + (int)gridDim.x < total) {
                    const int w2 = rev ? total - 1 - (w0 + (int)gridDim.x) : w0 + (int)gridDim.x;
                    const int slot2 = w2 / nblk, kx2 = 1 + (w2 - slot2 * nblk) * CB;
                    Zn = Z + ((size_t)PB_ZSLOT(slot2) * C * half + kx2) * NY;
                    ncn = min(CB, half - kx2);
                }
                if (Zn)
                    for (int col = 0; col < ncn; ++col) bulk_prefetch_l2(Zn + (size_t)col * NY, (uint32_t)(NY * sizeof(float2)));
            }
            cols2_stage_a<RA, RB, PC>(data, tws, bc, bj, ncol);
            // middle: DFT_RB -> y = H z with the re/im swap -> DFT_RB, on one padded block per thread
            for (int idx = tid; idx < ncol * RA; idx += FFTC2_THREADS) {
                const int f = idx / RA, blk = idx - f * RA;
                float2* p = data + (size_t)f * CS + blk * BS;
                float2 v[RB];
#pragma unroll
                for (int m = 0; m < RB; ++m) v[m] = p[m];
                Dft<RB>::run(v);
                const float4* h4 = reinterpret_cast<const float4*>(Hs + (size_t)f * HC + blk * HS);
#pragma unroll
                for (int m4 = 0; m4 < RB / 4; ++m4) {
                    const float4 h = h4[m4];
                    v[4 * m4 + 0] = make_float2(h.x * v[4 * m4 + 0].y, h.x * v[4 * m4 + 0].x);
                    v[4 * m4 + 1] = make_float2(h.y * v[4 * m4 + 1].y, h.y * v[4 * m4 + 1].x);
                    v[4 * m4 + 2] = make_float2(h.z * v[4 * m4 + 2].y, h.z * v[4 * m4 + 2].x);
                    v[4 * m4 + 3] = make_float2(h.w * v[4 * m4 + 3].y, h.w * v[4 * m4 + 3].x);
                }
                Dft<RB>::run(v);
#pragma unroll
                for (int m = 0; m < RB; ++m) p[m] = v[m];
            }
            __syncthreads();
            // DIT stage A': padded blocks -> natural order.  All loads, barrier, then the stores.
            if constexpr (PC) {
                // a warp is a column: idle warps skip the stage; this warp's column goes out, and the same column of the
                // next plane comes in as soon as it has been read out
                if (bc < ncol) {
                    float2 v[RA];
                    float2* p = data + (size_t)bc * CS + bj;
                    v[0] = p[0];
#pragma unroll
                    for (int q = 1; q < RA; ++q) v[q] = c_mul(p[q * BS], tws[(q - 1) * RB + bj]);
                    __syncwarp();
                    Dft<RA>::run(v);
#pragma unroll
                    for (int m = 0; m < RA; ++m) p[m * RB] = v[m];
                    fence_async_smem();
                    __syncwarp();
                    if (bj == 0) {
                        bulk_s2g(Zc + (size_t)bc * NY, data + (size_t)bc * CS, (uint32_t)(NY * sizeof(float2)));
                        bulk_commit();
                        bulk_wait_all();
                        if (c + 1 < C) {
                            mbar_expect_tx(bar + bc, (uint32_t)(NY * sizeof(float2)));
                            bulk_g2s(data + (size_t)bc * CS, Zc + ((size_t)half + bc) * NY, (uint32_t)(NY * sizeof(float2)), bar + bc);
                        }
                    }
                }
                continue;
            }
            {
                float2 v[RA];
                const bool on = bc < ncol;
                float2* p = data + (size_t)(on ? bc : ncol - 1) * CS + bj;
                v[0] = p[0];
#pragma unroll
                for (int q = 1; q < RA; ++q) v[q] = c_mul(p[q * BS], tws[(q - 1) * RB + bj]);
                __syncthreads();
                Dft<RA>::run(v);
#pragma unroll
                for (int m = 0; m < RA; ++m)
                    if (on) p[m * RB] = v[m];
            }
            fence_async_smem();
            __syncthreads();
            if (tid == 0) {
                for (int col = 0; col < ncol; ++col)
                    bulk_s2g(Zc + (size_t)col * NY, data + (size_t)col * CS, (uint32_t)(NY * sizeof(float2)));
                bulk_commit();
                bulk_wait_all();
                if (c + 1 < C) {
                    const float2* Zn = Zc + (size_t)half * NY;
                    mbar_expect_tx(bar, (uint32_t)((size_t)ncol * NY * sizeof(float2)));
                    for (int col = 0; col < ncol; ++col)
                        bulk_g2s(data + (size_t)col * CS, Zn + (size_t)col * NY, (uint32_t)(NY * sizeof(float2)), bar);
                }
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
    // memoised: the search plans ~n / 20 candidate lengths, and every entry point lays out its workspace twice per call
    // (pb_workspace_bytes, then the call itself) -- 0.7 ms of host time per 1080p call, 1.7 ms per 4K call without this
    static std::mutex mu;
    static std::unordered_map<int, int> cache;
    {
        std::lock_guard<std::mutex> lock(mu);
        auto it = cache.find(n);
        if (it != cache.end()) return it->second;
    }
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
    std::lock_guard<std::mutex> lock(mu);
    cache[n] = best;
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
    l.off_unitsX = take((size_t)(l.NX / 2 + 2) * sizeof(int4));
    l.off_stwY2 = take((size_t)l.NY * sizeof(float2));
    l.off_stwX2 = take((size_t)l.NX * sizeof(float2));
    l.off_unitsX2 = take((size_t)(l.NX / 2 + 2) * sizeof(int4));
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
    T->unitsX = reinterpret_cast<int4*>(base + L.off_unitsX);
    T->Z = reinterpret_cast<float2*>(base + L.off_Z);
    jobs->add(TJ_TWIDDLES, L.NX, T->twX, nullptr, nullptr);
    jobs->add(TJ_TWIDDLES, L.NY, T->twY, nullptr, nullptr);
    jobs->add(TJ_STAGE_TW, T->planX.tw_total, T->stwX, nullptr, &T->planX);
    jobs->add(TJ_STAGE_TW, T->planY.tw_total, T->stwY, nullptr, &T->planY);
    jobs->add(TJ_PERM, L.NX, T->slotX, nullptr, &T->planX);
    jobs->add(TJ_PERM, L.NY, T->slotY, T->freqY, &T->planY);
    jobs->add(TJ_UNITS, 256, T->unitsX, nullptr, &T->planX);
    T->stwY2 = reinterpret_cast<float2*>(base + L.off_stwY2);
    T->ra2 = T->rb2 = 0;
    if (L.NY == 36 * 32) {              // two-stage column plans with a compiled kernel (k_fft_cols2<RA, RB>)
        const bool big_first = env_int("PB_FFT_RA36", 1) != 0;      // measured: 1.92 ms against 2.02 ms per step (C2)
        T->ra2 = big_first ? 36 : 32;
        T->rb2 = big_first ? 32 : 36;
    }
    T->stwX2 = reinterpret_cast<float2*>(base + L.off_stwX2);
    T->unitsX2 = reinterpret_cast<int4*>(base + L.off_unitsX2);
    T->rowplan2 = 0;
    if (L.NX == 4000 && L.NY == 2304 && env_int("PB_FFT_ROWS_4STAGE", 0) == 0) {
        Fft2Plan px;
        px.n = 4000;
        px.ns = 3;
        px.radix[0] = 20;
        px.radix[1] = 20;
        px.radix[2] = 10;
        fft2_plan_offsets(&px);
        T->rowplan2 = 1;
        jobs->add(TJ_STAGE_TW, px.tw_total, T->stwX2, nullptr, &px);
        jobs->add(TJ_UNITS, 256, T->unitsX2, nullptr, &px);
    }
    if (L.NX == 12096 && L.NY == 9216 && env_int("PB_FFT_ROWS_C4_3STAGE", 1) != 0) {
        Fft2Plan px;
        px.n = 12096;
        px.ns = 3;
        px.radix[0] = 36;
        px.radix[1] = 24;
        px.radix[2] = 14;
        fft2_plan_offsets(&px);
        T->rowplan2 = 2;
        jobs->add(TJ_STAGE_TW, px.tw_total, T->stwX2, nullptr, &px);
        jobs->add(TJ_UNITS, 256, T->unitsX2, nullptr, &px);
    }
    if (T->ra2) {
        Fft2Plan p2;
        p2.n = L.NY;
        p2.ns = 2;
        p2.radix[0] = T->ra2;
        p2.radix[1] = T->rb2;
        fft2_plan_offsets(&p2);
        jobs->add(TJ_STAGE_TW, p2.tw_total, T->stwY2, nullptr, &p2);
    }
    return PB_OK;
}

// A side stream (per host thread and device) for the one-column DC / Nyquist launch of the column pass: it touches
// column 0 of every plane, the main launch columns 1 .., so the two run side by side -- forked and joined with
// events, which is also how a stream capture records them as parallel branches of the graph.
struct SideStream {
    int dev = -1;
    cudaStream_t s = nullptr;
    cudaEvent_t fork = nullptr, join = nullptr;
};
static int side_stream(SideStream** out) {
    static thread_local SideStream ss;
    int dev = 0;
    PB_CUDA_TRY(cudaGetDevice(&dev));
    if (ss.dev != dev) {                       // (a thread that moves to another device keeps one set per visit)
        PB_CUDA_TRY(cudaStreamCreateWithFlags(&ss.s, cudaStreamNonBlocking));
        PB_CUDA_TRY(cudaEventCreateWithFlags(&ss.fork, cudaEventDisableTiming));
        PB_CUDA_TRY(cudaEventCreateWithFlags(&ss.join, cudaEventDisableTiming));
        ss.dev = dev;
    }
    *out = &ss;
    return PB_OK;
}

int launch_deconv_fft(const float* img, float* out, const ImgKernel* kern, const int* list, const int* count,
                      int B, int C, int H, int W, const FftEngineTables& T, float a3, float a2, float a1,
                      float b0, const SrcGeom& G, cudaStream_t stream) {
    int rc = PB_OK;
    const int NX = T.NX, NY = T.NY;
    const int nb = rows_nb(NX), CB = cols_cb(NY);
    const size_t smem_rows = (size_t)nb * fftd_row_stride(NX) * sizeof(float2);
    const size_t smem_cols = (size_t)CB * NY * 12 + (size_t)NY * 8 + (size_t)(CB + 1) * 13 * 8 + 64;
    const long long row_items = (long long)B * C * ((NY / 2 + nb - 1) / nb);
    const long long col_items = (long long)B * ((NX / 2 + CB - 1) / CB);
    // CTAs launched per SM for the grid-stride kernels (0 = one work item per CTA).  Measured at C2: 6 per SM (two
    // waves of persistent CTAs, 15 items each) 1.375 / 1.498 ms per step for P1 / P3, 24 per SM 1.383 / 1.412, one item
    // per CTA 1.33 / 1.36: the hardware's CTA scheduler balances the tail better than long-lived CTAs do.
    static const int cap_per_sm = env_int("PB_FFT_CTAS_PER_SM", 0);
    const int cap = cap_per_sm > 0 ? PB_NUM_SMS * cap_per_sm : (1 << 30);
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
    // second-generation row passes (fused first / last stages) for the compile-time plans
#define PB_FFT_ROWS2(SP, NYC, TW, UN, TH, PADL)                                                                                        \
    do {                                                                                                         \
        auto kf = k_fft_rows_fwd2<SP, NYC, TH, PADL>;                                                                      \
        auto ki = k_fft_rows_inv2<SP, NYC, TH, PADL>;                                                                      \
        const size_t smem_rows = (size_t)nb * Rows2Layout<SP, PADL>::RS * sizeof(float2);      \
        PB_CUDA_TRY(cudaFuncSetAttribute(kf, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_rows));      \
        PB_CUDA_TRY(cudaFuncSetAttribute(ki, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_rows));      \
        if (fwd) {                                                                                               \
            ProfScope prof(PROF_FFT_ROWS_FWD, stream);                                                           \
            kf<<<grid_rows, TH, smem_rows, stream>>>(img, T.Z, kern, list, count, C, H, W, nb, TW, UN, G); \
        } else {                                                                                                 \
            ProfScope prof(PROF_FFT_ROWS_INV, stream);                                                           \
            ki<<<grid_rows, TH, smem_rows, stream>>>(T.Z, out, kern, list, count, C, H, W, nb, TW, UN,  \
                                                               G.clamp_out);                                     \
        }                                                                                                        \
    } while (0)
#define PB_FFT_COLS(SP)                                                                                          \
    do {                                                                                                         \
        PB_CUDA_TRY(cudaFuncSetAttribute(k_fft_cols<SP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_cols)); \
        ProfScope prof(PROF_FFT_COLS, stream);                                                                   \
        k_fft_cols<SP><<<grid_cols, FFTC_THREADS, smem_cols, stream>>>(T.Z, kern, list, count, C, NX, NY, CB, T.planY, \
                                                                       T.twX, T.stwY, T.slotY, a3, a2, a1, b0, 0); \
    } while (0)
    // second-generation column pass: columns 1 .. NX/2 - 1 by k_fft_cols2, column 0 by the first-generation kernel
#define PB_FFT_COLS2(RA, RB)                                                                                     \
    do {                                                                                                         \
        const int CB2 = FFTC2_THREADS / (RB) < 8 ? FFTC2_THREADS / (RB) : 8;                                     \
        const int cb2 = env_int("PB_FFT_CB2", 7) < CB2 ? env_int("PB_FFT_CB2", 7) : CB2;                         \
        const size_t cs2 = FFTC2_CSTRIDE(RA, RB);                                                                \
        const size_t smem2 = cb2 * cs2 * sizeof(float2) + (size_t)cb2 * (RA) * FFTC2_HSTRIDE(RB) * sizeof(float) + \
                             (size_t)cb2 * 13 * sizeof(float2) + (size_t)((RA) - 1) * (RB) * sizeof(float2) + 128;  \
        const size_t smem0 = (size_t)NY * 12 + (size_t)NY * 8 + 2 * 13 * 8 + 64;                                 \
        const long long items2 = (long long)B * ((NX / 2 - 1 + cb2 - 1) / cb2);                                  \
        /* 0 = one work item per CTA: with the per-column hand-off 1.566 ms per C2 step against 1.578 at 4 or 8 per SM */ \
        static const int c2_per_sm = env_int("PB_FFT_COLS2_CTAS_PER_SM", 0);                                     \
        const long long cap2 = c2_per_sm > 0 ? (long long)c2_per_sm * PB_NUM_SMS : (1LL << 30);                  \
        const int grid2 = (int)(items2 < cap2 ? items2 : cap2);                                                  \
        static const bool percol = env_int("PB_FFT_COLS2_PERCOL", 1) != 0;     /* per-column hand-off between planes */ \
        auto kc2 = percol ? k_fft_cols2<RA, RB, (RB) == 32> : k_fft_cols2<RA, RB, false>;                        \
        PB_CUDA_TRY(cudaFuncSetAttribute(kc2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));         \
        PB_CUDA_TRY(cudaFuncSetAttribute(k_fft_cols<NoStaticPlan>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem0)); \
        ProfScope prof(PROF_FFT_COLS, stream);                                                                   \
        static const bool side_on = env_int("PB_FFT_SIDE", 1) != 0;                                              \
        SideStream* ss = nullptr;                                                                                \
        if (side_on && (rc = side_stream(&ss))) return rc;                                                       \
        cudaStream_t s0 = stream;                                                                                \
        if (ss) {                                                                                                \
            PB_CUDA_TRY(cudaEventRecord(ss->fork, stream));                                                      \
            PB_CUDA_TRY(cudaStreamWaitEvent(ss->s, ss->fork, 0));                                                \
            s0 = ss->s;                                                                                          \
        }                                                                                                        \
        k_fft_cols<NoStaticPlan><<<dim3(B < cap ? B : cap, C), FFTC_THREADS, smem0, s0>>>(T.Z, kern, list, count, C, NX, NY, 1, \
                                                                     T.planY, T.twX, T.stwY, T.slotY, a3, a2, a1, b0, 1); \
        if (ss) PB_CUDA_TRY(cudaEventRecord(ss->join, ss->s));                                                   \
        static const int rev2 = env_int("PB_REVERSE", 1);                                                        \
        kc2<<<grid2, FFTC2_THREADS, smem2, stream>>>(T.Z, kern, list, count, C, NX, cb2, T.twX, T.stwY2,         \
                                                     a3, a2, a1, b0, rev2);                                      \
        if (ss) PB_CUDA_TRY(cudaStreamWaitEvent(stream, ss->join, 0));                                           \
    } while (0)
    // long columns (one column per CTA): the block of column 0 in its own launch, so that the other CTAs do without its
    // two extra NY-float arrays and two of them fit an SM
#define PB_FFT_COLS_LONG(SP, TH)                                                                                 \
    do {                                                                                                         \
        const size_t smem1 = (size_t)CB * NY * 12 + (size_t)(CB + 1) * 13 * 8 + 64;                              \
        const long long items1 = (long long)B * ((NX / 2 + CB - 1) / CB - 1);                                    \
        PB_CUDA_TRY(cudaFuncSetAttribute(k_fft_cols<SP, TH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1)); \
        PB_CUDA_TRY(cudaFuncSetAttribute(k_fft_cols<SP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_cols)); \
        ProfScope prof(PROF_FFT_COLS, stream);                                                                   \
        static const bool side_on = env_int("PB_FFT_SIDE", 1) != 0;                                              \
        SideStream* ss = nullptr;                                                                                \
        if (side_on && (rc = side_stream(&ss))) return rc;                                                       \
        cudaStream_t s0 = stream;                                                                                \
        if (ss) {                                                                                                \
            PB_CUDA_TRY(cudaEventRecord(ss->fork, stream));                                                      \
            PB_CUDA_TRY(cudaStreamWaitEvent(ss->s, ss->fork, 0));                                                \
            s0 = ss->s;                                                                                          \
        }                                                                                                        \
        k_fft_cols<SP><<<dim3(B < cap ? B : cap, C), FFTC_THREADS, smem_cols, s0>>>(                             \
            T.Z, kern, list, count, C, NX, NY, CB, T.planY, T.twX, T.stwY, T.slotY, a3, a2, a1, b0, 1);          \
        if (ss) PB_CUDA_TRY(cudaEventRecord(ss->join, ss->s));                                                   \
        k_fft_cols<SP, TH><<<(int)(items1 < cap ? items1 : cap), TH, smem1, stream>>>(                           \
            T.Z, kern, list, count, C, NX, NY, CB, T.planY, T.twX, T.stwY, T.slotY, a3, a2, a1, b0, 2);          \
        if (ss) PB_CUDA_TRY(cudaStreamWaitEvent(stream, ss->join, 0));                                           \
    } while (0)
    static const bool rows_v1 = env_int("PB_FFT_ROWS_V1", 0) != 0;     // A/B against the first-generation passes
    // padded stage-1 blocks (Rows2Layout): 1 = the 4K plan only, 2 = the 1080p plan too.  Measured: 4K P1 1.58 -> 1.53,
    // P3 1.60 -> 1.57 ms per step (8 images); 1080p P1 1.33 -> 1.52, P3 1.36 -> 1.66 -- there the three CTAs grow from
    // 194 to 211 KB of shared memory and the 16 KB of L1 that remain no longer hold the twiddle tables.
    static const int rows_pad = env_int("PB_FFT_ROWS_PAD", 1);
    static const bool cols_v1 = env_int("PB_FFT_COLS_V1", 0) != 0;
    for (int pass = 0; pass < 3; ++pass) {
        const bool fwd = pass == 0;
        if (pass == 1) {
            if (T.ra2 == 36 && T.rb2 == 32 && !cols_v1) PB_FFT_COLS2(36, 32);
            else if (T.ra2 == 32 && T.rb2 == 36 && !cols_v1) PB_FFT_COLS2(32, 36);
            else if (PlanY1152::matches(T.planY)) PB_FFT_COLS(PlanY1152);
            else if (PlanY2304::matches(T.planY)) {
                // (the block of column 0 in its own launch, so that three CTAs take 166 instead of 222 KB of shared memory and
                // the L1 that remains holds the 18 KB of stage twiddles, measured slower here: 2.23 against 2.16 ms per step)
                static const int split = env_int("PB_FFT_COLS_SPLIT", 0);
                if (split && !cols_v1) PB_FFT_COLS_LONG(PlanY2304, 256); else PB_FFT_COLS(PlanY2304);
            }
            else if (PlanY9216::matches(T.planY) && CB == 1 && !cols_v1) {
                static const int th_long = env_int("PB_FFT_COLS_LONG_T", 384);
                if (th_long == 384) PB_FFT_COLS_LONG(PlanY9216, 384); else PB_FFT_COLS_LONG(PlanY9216, 256);
            }
            else PB_FFT_COLS(NoStaticPlan);
        } else {
            // (the second-generation kernels also fix NY at compile time: the 1080p and 4K tori)
            if (PlanX2016::matches(T.planX)) {
                if (rows_v1 || NY != 1152) PB_FFT_ROWS(PlanX2016);
                else if (rows_pad >= 2) PB_FFT_ROWS2(PlanX2016, 1152, T.stwX, T.unitsX, FFTD_THREADS, true);
                else PB_FFT_ROWS2(PlanX2016, 1152, T.stwX, T.unitsX, FFTD_THREADS, false);
            } else if (PlanX4000::matches(T.planX)) {
                if (rows_v1 || NY != 2304) PB_FFT_ROWS(PlanX4000);
                else if (T.rowplan2 == 1 && rows_pad) PB_FFT_ROWS2(PlanX4000b, 2304, T.stwX2, T.unitsX2, FFTD_THREADS, true);
                else if (T.rowplan2 == 1) PB_FFT_ROWS2(PlanX4000b, 2304, T.stwX2, T.unitsX2, FFTD_THREADS, false);
                else PB_FFT_ROWS2(PlanX4000, 2304, T.stwX, T.unitsX, FFTD_THREADS, false);
            } else if (PlanX12096::matches(T.planX) && NY == 9216 && !rows_v1) {
                // one row pair (95 KB) per CTA, two CTAs per SM (384 threads per CTA, which fill the register file at
                // 2 x 384 x 80, measured slower: P1 7.1 against 6.5 ms per C4 step, P3 equal)
                if (T.rowplan2 == 2) PB_FFT_ROWS2(PlanX12096b, 9216, T.stwX2, T.unitsX2, 256, false);
                else PB_FFT_ROWS2(PlanX12096, 9216, T.stwX, T.unitsX, 256, false);
            } else PB_FFT_ROWS(NoStaticPlan);
        }
    }
#undef PB_FFT_ROWS
#undef PB_FFT_ROWS2
#undef PB_FFT_COLS
#undef PB_FFT_COLS2
#undef PB_FFT_COLS_LONG
    PB_LAUNCH_CHECK("fft deconvolution passes");
    return PB_OK;
}

}  // namespace pb
