// Quantile range normalisation of the gray image (q > 0):
//   blur_estimation.normalize (polyblur/blur_estimation.py:96-109) = torch.quantile(q) and
//   torch.quantile(1 - q) per image, linear interpolation between order statistics, followed by
//   clamp_((x - lo) / (hi - lo), 0, 1) (:92-93).
//
// The four order statistics per image (floor / ceil rank of the two quantiles) are found exactly
// by a three-pass radix select over the order-preserving integer encoding of the floats
// (11 + 11 + 10 bits): each pass histograms the gray plane in shared memory (4 B/px read), a tiny
// per-image kernel locates the bin that holds each target rank and narrows the prefix.  No sort,
// no host round trip.
#include "kernels.cuh"

namespace pb {

#define Q_TARGETS 4
#define Q_BINS 2048

struct QState {                  // per image
    unsigned prefix[Q_TARGETS];  // bits fixed so far (right aligned)
    unsigned rank[Q_TARGETS];    // rank of the target among the keys that share the prefix
};

// gray = channel mean (blur_estimation.py:36-37), same arithmetic as k_rows2 / k_cols
__global__ void __launch_bounds__(256)
k_gray(const float* __restrict__ img, float* __restrict__ gray, int C, size_t plane) {
    const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= plane) return;
    const float* p = img + (size_t)blockIdx.y * C * plane + i;
    float g = __ldg(p);
    for (int c = 1; c < C; ++c) g = __fadd_rn(g, __ldg(p + (size_t)c * plane));
    if (C > 1) g = __fdiv_rn(g, (float)C);
    gray[(size_t)blockIdx.y * plane + i] = g;
}

__global__ void k_q_init(QState* __restrict__ st, int B, unsigned r0, unsigned r1, unsigned r2, unsigned r3) {
    const int im = blockIdx.x * blockDim.x + threadIdx.x;
    if (im < B) {
        const unsigned r[Q_TARGETS] = {r0, r1, r2, r3};
        for (int t = 0; t < Q_TARGETS; ++t) {
            st[im].prefix[t] = 0u;
            st[im].rank[t] = r[t];
        }
    }
}

// PASS 0: bins = key >> 21 (one histogram, shared by the four targets)
// PASS 1: keys with (key >> 21) == prefix[t]:  bin = (key >> 10) & 2047   (one histogram per target)
// PASS 2: keys with (key >> 10) == prefix[t]:  bin = key & 1023
template <int PASS>
__global__ void __launch_bounds__(256)
k_q_hist(const float* __restrict__ gray, const QState* __restrict__ st, unsigned* __restrict__ hist, size_t plane,
         size_t per_cta) {
    __shared__ unsigned sh[(PASS == 0 ? 1 : Q_TARGETS) * Q_BINS];
    constexpr int NH = (PASS == 0 ? 1 : Q_TARGETS);
    const int im = blockIdx.y;
    for (int i = threadIdx.x; i < NH * Q_BINS; i += 256) sh[i] = 0u;
    unsigned pre[Q_TARGETS];
    for (int t = 0; t < Q_TARGETS; ++t) pre[t] = st[im].prefix[t];
    __syncthreads();
    const float* g = gray + (size_t)im * plane;
    const size_t i0 = (size_t)blockIdx.x * per_cta;
    const size_t i1 = i0 + per_cta < plane ? i0 + per_cta : plane;
    for (size_t i = i0 + threadIdx.x; i < i1; i += 256) {
        const unsigned key = f2ord(__ldg(g + i));
        if (PASS == 0) {
            atomicAdd(&sh[key >> 21], 1u);
        } else {
#pragma unroll
            for (int t = 0; t < Q_TARGETS; ++t) {
                if (PASS == 1) {
                    if ((key >> 21) == pre[t]) atomicAdd(&sh[t * Q_BINS + ((key >> 10) & 2047u)], 1u);
                } else {
                    if ((key >> 10) == pre[t]) atomicAdd(&sh[t * Q_BINS + (key & 1023u)], 1u);
                }
            }
        }
    }
    __syncthreads();
    unsigned* h = hist + (size_t)im * Q_TARGETS * Q_BINS;
    for (int i = threadIdx.x; i < NH * Q_BINS; i += 256)
        if (sh[i]) atomicAdd(&h[i], sh[i]);
}

// one warp per image: lane t < 4 walks its histogram until the bin that holds its rank
template <int PASS>
__global__ void k_q_select(QState* __restrict__ st, const unsigned* __restrict__ hist, int B, float frac_lo,
                           float frac_hi, float* __restrict__ qrange) {
    const int im = blockIdx.x;
    const int t = threadIdx.x;
    if (t < Q_TARGETS) {
        const unsigned* h = hist + (size_t)im * Q_TARGETS * Q_BINS + (PASS == 0 ? 0 : t * Q_BINS);
        const int nb = (PASS == 2) ? 1024 : Q_BINS;
        unsigned r = st[im].rank[t], cum = 0;
        int b = 0;
        for (; b < nb - 1; ++b) {
            const unsigned c = h[b];
            if (r < cum + c) break;
            cum += c;
        }
        st[im].rank[t] = r - cum;
        st[im].prefix[t] = (PASS == 0) ? (unsigned)b : ((st[im].prefix[t] << (PASS == 1 ? 11 : 10)) | (unsigned)b);
    }
    if (PASS == 2) {
        __syncwarp();
        if (t == 0) {
            // targets 0,1 = floor / ceil rank of q; 2,3 = of 1 - q; lerp in fp32 like torch.quantile
            const float a0 = ord2f(st[im].prefix[0]), a1 = ord2f(st[im].prefix[1]);
            const float b0 = ord2f(st[im].prefix[2]), b1 = ord2f(st[im].prefix[3]);
            qrange[2 * im + 0] = __fadd_rn(a0, __fmul_rn(frac_lo, __fsub_rn(a1, a0)));
            qrange[2 * im + 1] = __fadd_rn(b0, __fmul_rn(frac_hi, __fsub_rn(b1, b0)));
        }
    }
}

size_t quantile_workspace_bytes(int B) {
    return align_up((size_t)B * sizeof(QState), 256) + align_up((size_t)B * Q_TARGETS * Q_BINS * sizeof(unsigned), 256) +
           align_up((size_t)B * 2 * sizeof(float), 256);
}

// img (B,C,H,W) -> gray (B,H,W) un-normalised, qrange[2 im] = (lo, hi) of image im (inside ws).
int launch_quantile_range(const float* img, float* gray, int B, int C, int H, int W, double q, void* ws,
                          float** qrange_out, cudaStream_t stream) {
    if (B > 65535) {
        set_error("batch too large for the quantile grids");
        return PB_ERR_ARG;
    }
    const size_t plane = (size_t)H * W;
    char* base = static_cast<char*>(ws);
    QState* st = reinterpret_cast<QState*>(base);
    unsigned* hist = reinterpret_cast<unsigned*>(base + align_up((size_t)B * sizeof(QState), 256));
    float* qrange = reinterpret_cast<float*>(base + align_up((size_t)B * sizeof(QState), 256) +
                                             align_up((size_t)B * Q_TARGETS * Q_BINS * sizeof(unsigned), 256));
    // ranks exactly as torch.quantile forms them: pos = fl32(q) * fl32(n - 1), floor, ceil, fraction in fp32
    auto ranks = [&](double qq, unsigned* lo, unsigned* hi, float* frac) {
        const float pos = (float)qq * (float)(plane - 1);
        const float fl = floorf(pos);
        *lo = (unsigned)fl;
        *hi = (*lo + 1 < plane) ? *lo + 1 : (unsigned)(plane - 1);
        *frac = pos - fl;
    };
    unsigned r0, r1, r2, r3;
    float f_lo, f_hi;
    ranks(q, &r0, &r1, &f_lo);
    ranks(1.0 - q, &r2, &r3, &f_hi);
    ProfScope prof(PROF_OTHER, stream);
    k_gray<<<dim3((unsigned)((plane + 255) / 256), B), 256, 0, stream>>>(img, gray, C, plane);
    k_q_init<<<(B + 127) / 128, 128, 0, stream>>>(st, B, r0, r1, r2, r3);
    const int ctas = (int)((plane + 65535) / 65536);      // 64 Ki pixels per CTA
    const size_t per_cta = (plane + ctas - 1) / ctas;
    const size_t hist_bytes = (size_t)B * Q_TARGETS * Q_BINS * sizeof(unsigned);
    PB_CUDA_TRY(cudaMemsetAsync(hist, 0, hist_bytes, stream));
    k_q_hist<0><<<dim3(ctas, B), 256, 0, stream>>>(gray, st, hist, plane, per_cta);
    k_q_select<0><<<B, 32, 0, stream>>>(st, hist, B, f_lo, f_hi, qrange);
    PB_CUDA_TRY(cudaMemsetAsync(hist, 0, hist_bytes, stream));
    k_q_hist<1><<<dim3(ctas, B), 256, 0, stream>>>(gray, st, hist, plane, per_cta);
    k_q_select<1><<<B, 32, 0, stream>>>(st, hist, B, f_lo, f_hi, qrange);
    PB_CUDA_TRY(cudaMemsetAsync(hist, 0, hist_bytes, stream));
    k_q_hist<2><<<dim3(ctas, B), 256, 0, stream>>>(gray, st, hist, plane, per_cta);
    k_q_select<2><<<B, 32, 0, stream>>>(st, hist, B, f_lo, f_hi, qrange);
    PB_LAUNCH_CHECK("quantile select");
    *qrange_out = qrange;
    return PB_OK;
}

}  // namespace pb
