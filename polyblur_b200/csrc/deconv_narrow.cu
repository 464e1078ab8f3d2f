// Narrow-kernel deconvolution engine: the three Horner stencils of
//     o = a3 p;  o = K (*) o + a2 p;  o = K (*) o + a1 p;  o = K (*) o + b p
// evaluated in the algebraically identical, better conditioned basis D = K - I
//     o = c3 p;  o = D (*) o + c2 p;  o = D (*) o + c1 p;  o = D (*) o + p
// (deblurring.inverse_filtering_rank3 / compute_polynomial_fft, polyblur/deblurring.py:211-239,
// 141-169, on the replicate-padded torus of SURVEY.md A.6) kept entirely in registers for blur
// kernels whose significant taps fit (2 RX + 1) x (2 RY + 1), RX, RY <= 2 -- what the estimator
// returns for sharp or mildly blurred images (sigma, rho <~ 0.45).
//
// One warp owns a strip of 128 columns and marches down the rows.  Each lane owns 4 adjacent
// columns (one 128-bit load / store per row); the +-RX neighbours come from warp shuffles, the
// +-RY neighbour rows from rolling register windows of the three stages (the row loop is unrolled
// over the window period so that the window slots are compile-time register names).  Nothing
// goes through shared memory and there is no block-level barrier.  The outermost lanes of a warp
// miss their halo, so a warp's valid output is the inner 128 - 8 HL columns and neighbouring
// warps overlap by 2 HL lanes (HL = ceil(3 RX / 4)).
//
// HBM traffic: 4 B read + 4 B written per pixel-channel; the overlap columns (6-12 %) and the
// 6 RY halo rows per tile are re-read through L2.
//
// Work items (image of this class, channel, tile) are taken grid-stride from the per-class image
// list that k_params fills on the device, so no host synchronisation is needed to choose the
// engine per image.
#include <cstdlib>

#include "kernels.cuh"

namespace pb {

#ifndef PB_NARROW_PF
#define PB_NARROW_PF 2       // window periods of prefetch for the 3x3 kernel
#endif

template <int RX, int RY>
struct NarrowCfg {
    static constexpr int HL = (3 * RX + 3) / 4;       // halo lanes per side
    static constexpr int VW = 128 - 8 * HL;           // valid output columns per warp
    static constexpr int NW = 2 * RY + 1;             // rows in a rolling window
    static constexpr int D = NW * (RY == 1 ? PB_NARROW_PF : 1);  // prefetch distance (rows) = unroll period
    static constexpr int PW = 4 + 2 * RX;             // row segment a lane sees
};

#define NARROW_THREADS 128
#ifndef PB_NARROW_MINB
#define PB_NARROW_MINB 4    // 3x3 kernel: 4 CTAs x 4 warps at <= 128 registers
#endif

template <int RX, int RY>
__global__ void __launch_bounds__(NARROW_THREADS, (RX == 1 ? PB_NARROW_MINB : 1))
k_deconv_narrow(const float* __restrict__ img, float* __restrict__ out, const ImgKernel* __restrict__ kern,
                const int* __restrict__ list, const int* __restrict__ count, int C, int H, int W, int TH,
                float a3, float a2, float a1, float b0, SrcGeom G) {
    using Cfg = NarrowCfg<RX, RY>;
    constexpr int HL = Cfg::HL, VW = Cfg::VW, NW = Cfg::NW, D = Cfg::D, PW = Cfg::PW;
    constexpr unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int gw = (blockIdx.x * NARROW_THREADS + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * NARROW_THREADS) >> 5;
    const int tilesX = (W + VW - 1) / VW, tilesY = (H + TH - 1) / TH;
    const int per_plane = tilesX * tilesY;
    const int per_img = C * per_plane;
    const int total = count[0] * per_img;
    const size_t plane = (size_t)H * W;

    for (int w = gw; w < total; w += nwarps) {
        const int slot = w / per_img;
        int r = w - slot * per_img;
        const int c = r / per_plane;
        r -= c * per_plane;
        const int ty = r / tilesX;
        const int tx = r - ty * tilesX;
        const int im = list[slot];
        const ImgKernel* K = kern + im;
        const int pad = G.pad >= 0 ? G.pad : (K->ksize >> 1);

        float wk[NW][2 * RX + 1];
#pragma unroll
        for (int dy = 0; dy < NW; ++dy)
#pragma unroll
            for (int dx = 0; dx < 2 * RX + 1; ++dx)
                wk[dy][dx] = __ldg(&K->k[(dy - RY + PB_PAD) * PB_KS + (dx - RX + PB_PAD)]);
        wk[RY][RX] -= 1.0f;       // D = K - I: the engines evaluate I + c1 D + c2 D^2 + c3 D^3 (api.cu)

        const float* src = img + ((size_t)im * C + c) * (size_t)G.Hin * G.Win;
        float* dst = out + ((size_t)im * C + c) * plane;
        const int y0 = ty * TH;
        const int rows = min(TH, H - y0);
        const int nsteps = rows + 6 * RY;
        const int cx0 = tx * VW - 4 * HL;
        const int cx = cx0 + 4 * lane;
        const bool fastx = (cx0 + G.off >= 0) && (cx0 + G.off + 128 <= G.Win) && (((G.Win | G.off) & 3) == 0);
        int sx[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) sx[i] = geom_src(cx + i, G.Win, G.off, pad);

        // Row j of the tile's input window (rows past the end repeat the last one: their results
        // are never stored, and an unconditional load keeps the row loop free of branches).  Tiles
        // whose window lies inside the source plane skip the torus map.
        const int ytop = y0 - 3 * RY + G.off;
        const bool rows_inside = ytop >= 0 && ytop + nsteps <= G.Hin;
        auto load_row = [&](int j) -> float4 {
            const int jj = min(j, nsteps - 1);
            const int sy = rows_inside ? ytop + jj : geom_src(y0 - 3 * RY + jj, G.Hin, G.off, pad);
            const float* rp = src + (size_t)sy * G.Win;
            float4 v;
            if (fastx) {
                v = __ldg(reinterpret_cast<const float4*>(rp + cx + G.off));
            } else {
                v.x = __ldg(rp + sx[0]);
                v.y = __ldg(rp + sx[1]);
                v.z = __ldg(rp + sx[2]);
                v.w = __ldg(rp + sx[3]);
            }
            return v;
        };

        float4 pre[D];
#pragma unroll
        for (int u = 0; u < D; ++u) pre[u] = load_row(u);
        // interior tiles (no torus map in either direction): the prefetch pointer just walks down
        const bool walk = rows_inside && fastx;
        const float* pl = src + (size_t)(walk ? ytop + D : 0) * G.Win + cx + G.off;
        float* gs = dst + (size_t)y0 * W + cx;               // output row pointer, advanced once rows start

        float P[NW][PW], O1[NW][PW], O2[NW][PW], Q[NW][4];
#pragma unroll
        for (int s = 0; s < NW; ++s) {
#pragma unroll
            for (int i = 0; i < PW; ++i) P[s][i] = O1[s][i] = O2[s][i] = 0.f;
#pragma unroll
            for (int i = 0; i < 4; ++i) Q[s][i] = 0.f;
        }

        // fill the +-RX halo of a freshly produced row segment from the neighbouring lanes
#define NARROW_HALO(ROW)                                                              \
    _Pragma("unroll") for (int i = 0; i < RX; ++i) {                                  \
        ROW[RX - 1 - i] = __shfl_up_sync(FULL, ROW[RX + 3 - i], 1);                   \
        ROW[RX + 4 + i] = __shfl_down_sync(FULL, ROW[RX + i], 1);                     \
    }
        // stencil of window WIN centred on the row whose slot is CS, for the lane's 4 columns
#define NARROW_STENCIL(WIN, CS, ACC)                                                  \
    _Pragma("unroll") for (int cc = 0; cc < 4; ++cc) {                                \
        float acc_ = 0.f;                                                             \
        _Pragma("unroll") for (int dy = 0; dy < NW; ++dy) {                           \
            const int ps_ = ((CS) + dy - RY + 2 * NW) % NW;                           \
            _Pragma("unroll") for (int dx = 0; dx < 2 * RX + 1; ++dx)                 \
                acc_ = fmaf(wk[dy][dx], WIN[ps_][cc + dx], acc_);                     \
        }                                                                             \
        ACC[cc] = acc_;                                                               \
    }

        // Every step runs all three stages (the first 6 RY steps work on the zero-filled windows;
        // nothing is stored for them), so the loop body is straight-line code.
        const bool lane_ok = lane >= HL && lane < 32 - HL && cx < W;
        const bool vec_ok = ((W & 3) == 0) && cx + 3 < W;
        for (int jb = 0; jb < nsteps; jb += D) {
#pragma unroll
            for (int u = 0; u < D; ++u) {
                const int j = jb + u;
                const float4 v = pre[u];
                if (walk) {
                    if (j + D < nsteps) pre[u] = __ldg(reinterpret_cast<const float4*>(pl));
                    pl += G.Win;
                } else {
                    pre[u] = load_row(j + D);
                }
                const int s0 = u % NW;                           // slot of input row j
                P[s0][RX + 0] = v.x;
                P[s0][RX + 1] = v.y;
                P[s0][RX + 2] = v.z;
                P[s0][RX + 3] = v.w;
                NARROW_HALO(P[s0]);
                {
                    const int s1 = (u - RY + 2 * NW) % NW;       // slot of row j - RY
                    float acc[4];
                    NARROW_STENCIL(P, s1, acc);
#pragma unroll
                    for (int cc = 0; cc < 4; ++cc) O1[s1][RX + cc] = fmaf(a3, acc[cc], a2 * P[s1][RX + cc]);
                    NARROW_HALO(O1[s1]);
                }
                {
                    const int s2 = (u - 2 * RY + 4 * NW) % NW;   // slot of row j - 2 RY
                    float acc[4];
                    NARROW_STENCIL(O1, s2, acc);
#pragma unroll
                    for (int cc = 0; cc < 4; ++cc) {
                        const float pv = P[s2][RX + cc];
                        O2[s2][RX + cc] = fmaf(a1, pv, acc[cc]);
                        Q[s2][cc] = pv;                          // b p is added by the last stage
                    }
                    NARROW_HALO(O2[s2]);
                }
                {
                    const int s3 = (u - 3 * RY + 6 * NW) % NW;   // slot of row j - 3 RY
                    float acc[4];
                    NARROW_STENCIL(O2, s3, acc);
                    const int yo = j - 6 * RY;                   // output row within the tile
                    if (lane_ok && yo >= 0 && yo < rows) {
                        float o[4];
#pragma unroll
                        for (int cc = 0; cc < 4; ++cc) {
                            o[cc] = fmaf(b0, Q[s3][cc], acc[cc]);
                            if (G.clamp_out) o[cc] = fminf(fmaxf(o[cc], 0.0f), 1.0f);
                        }
                        float* g = gs;
                        if (vec_ok) {
                            *reinterpret_cast<float4*>(g) = make_float4(o[0], o[1], o[2], o[3]);
                        } else {
#pragma unroll
                            for (int cc = 0; cc < 4; ++cc)
                                if (cx + cc < W) g[cc] = o[cc];
                        }
                    }
                    if (yo >= 0) gs += W;
                }
            }
        }
#undef NARROW_HALO
#undef NARROW_STENCIL
    }
}

template <int RX, int RY>
static int launch_one(const float* img, float* out, const ImgKernel* kern, const int* list, const int* count,
                      int B, int C, int H, int W, float a3, float a2, float a1, float b0, const SrcGeom& G,
                      cudaStream_t stream) {
    using Cfg = NarrowCfg<RX, RY>;
    // tile height: tall tiles amortise the 6 RY warm-up rows; keep >= ~8 waves of warps
    int TH = 120;
    if (H < TH) TH = H;
    const int tilesX = (W + Cfg::VW - 1) / Cfg::VW, tilesY = (H + TH - 1) / TH;
    const long long items = (long long)B * C * tilesX * tilesY;
    const long long want = (items + 3) / 4;
    static const int per_sm = [] { const char* v = getenv("PB_NARROW_CTAS_PER_SM"); return v ? atoi(v) : 16; }();   // measured at C2 white: 8 per SM 1.58 ms per step, 16: 1.50, unlimited: 1.51
    const long long capn = per_sm > 0 ? (long long)PB_NUM_SMS * per_sm : (1LL << 30);
    int grid = (int)(want < capn ? want : capn);
    if (grid < 1) grid = 1;
    k_deconv_narrow<RX, RY><<<grid, NARROW_THREADS, 0, stream>>>(img, out, kern, list, count, C, H, W, TH, a3, a2,
                                                                 a1, b0, G);
    PB_LAUNCH_CHECK("k_deconv_narrow");
    return PB_OK;
}

int launch_deconv_narrow(int cls, const float* img, float* out, const ImgKernel* kern, const int* list,
                         const int* count, int B, int C, int H, int W, float a3, float a2, float a1, float b0,
                         const SrcGeom& G, cudaStream_t stream) {
    ProfScope prof(PROF_DECONV_NARROW, stream);
    if (cls == PB_CLS_N11) return launch_one<1, 1>(img, out, kern, list, count, B, C, H, W, a3, a2, a1, b0, G, stream);
    return launch_one<2, 2>(img, out, kern, list, count, B, C, H, W, a3, a2, a1, b0, G, stream);
}

}  // namespace pb
