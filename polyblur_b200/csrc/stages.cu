// Optional stages of the Polyblur loop, all on the device:
//   bilateral prefilter        filters.bilateral_filter            polyblur/filters.py:107-148
//   domain-transform RF        domain_transform.recursive_filter   polyblur/domain_transform.py:6-85
//   prefilter residual         impred = deconv(smooth) + (impred - smooth), clip   deblurring.py:80-88
//   halo masking               deblurring.halo_masking             polyblur/deblurring.py:173-208
//   edgetaper                  edgetaper.edgetaper                 polyblur/edgetaper.py:10-33
#include "kernels.cuh"

namespace pb {

// ---------------------------------------------------------------------------------------------
// 5x5 bilateral filter, per-channel range weight, replicate border (filters.py:107-148):
//   F = exp(-(S - I)^2 / (2 sc^2)) * exp(-(dx^2 + dy^2) / (2 ss^2));  out = sum F S / (sum F + 1e-5)
// One thread per pixel-channel; the 5x5 window comes through L1.  Row sums are formed first and
// then accumulated, like the reference's loop over kernel rows.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_bilateral(const float* __restrict__ img, float* __restrict__ out, int H, int W, float var2_spatial,
            float var2_color) {
    __shared__ float gw[25];
    if (threadIdx.x < 25) {
        const int dy = threadIdx.x / 5 - 2, dx = threadIdx.x % 5 - 2;
        gw[threadIdx.x] = expf(-(float)(dx * dx + dy * dy) / var2_spatial);      // (filters.py:110-112)
    }
    __syncthreads();
    // one thread = 4 adjacent pixels of one row: the 5 x 8 window is loaded once
    const int x0 = (blockIdx.x * 32 + (threadIdx.x & 31)) * 4;
    const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x0 >= W || y >= H) return;
    const size_t pl = (size_t)blockIdx.z * H * W;
    const float* p = img + pl;
    const float nic = -1.0f / var2_color;
    float win[5][8];
#pragma unroll
    for (int dy = 0; dy < 5; ++dy) {
        const float* row = p + (size_t)min(max(y + dy - 2, 0), H - 1) * W;
#pragma unroll
        for (int i = 0; i < 8; ++i) win[dy][i] = __ldg(row + min(max(x0 + i - 2, 0), W - 1));
    }
    float o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float I = win[2][k + 2];
        float J = 0.f, Wt = 0.f;
#pragma unroll
        for (int dy = 0; dy < 5; ++dy) {
            float jr = 0.f, wr = 0.f;
#pragma unroll
            for (int dx = 0; dx < 5; ++dx) {
                const float S = win[dy][k + dx];
                const float d = __fsub_rn(S, I);
                // exp(-(d d) / var2) * gw: the hardware exponential is accurate to ~1e-7 relative for
                // the arguments that matter (weights that are not negligible)
                const float F = __fmul_rn(__expf(__fmul_rn(__fmul_rn(d, d), nic)), gw[dy * 5 + dx]);
                jr = __fadd_rn(jr, __fmul_rn(F, S));
                wr = __fadd_rn(wr, F);
            }
            J = __fadd_rn(J, jr);
            Wt = __fadd_rn(Wt, wr);
        }
        o[k] = __fdiv_rn(J, __fadd_rn(Wt, 1e-5f));
    }
    float* g = out + pl + (size_t)y * W + x0;
    if (((W & 3) == 0) && x0 + 3 < W) {
        *reinterpret_cast<float4*>(g) = make_float4(o[0], o[1], o[2], o[3]);
    } else {
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (x0 + k < W) g[k] = o[k];
    }
}

// Vector-Jacobian product of the filter above: out_p = J_p / (W_p + eps), J_p = sum_q F_pq x_q, W_p = sum_q F_pq,
// F_pq = gw(q - p) exp(-(x_q - x_p)^2 / var2), q over the 5 x 5 window with replicate clamping.  With a = g_p / (W_p + eps)
// and Fbar_pq = a (x_q - out_p):   xbar_q += a F_pq - t_pq,   xbar_p += t_pq,   t_pq = Fbar_pq F_pq 2 (x_q - x_p) / var2.
// One thread per pixel-channel p recomputes its window, scatters into the (clamped) window pixels with atomics and
// adds its own term once; grad_in must be zeroed by the caller (launch_bilateral_vjp does).
__global__ void __launch_bounds__(256)
k_bilateral_vjp(const float* __restrict__ img, const float* __restrict__ gout, float* __restrict__ gin, int H, int W,
                float var2_spatial, float var2_color) {
    __shared__ float gw[25];
    if (threadIdx.x < 25) {
        const int dy = threadIdx.x / 5 - 2, dx = threadIdx.x % 5 - 2;
        gw[threadIdx.x] = expf(-(float)(dx * dx + dy * dy) / var2_spatial);
    }
    __syncthreads();
    const int x = blockIdx.x * 32 + (threadIdx.x & 31);
    const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= W || y >= H) return;
    const size_t pl = (size_t)blockIdx.z * H * W;
    const float* p = img + pl;
    const float nic = -1.0f / var2_color;
    const float I = __ldg(p + (size_t)y * W + x);
    float S[25], F[25];
    float J = 0.f, Wt = 0.f;
#pragma unroll
    for (int dy = 0; dy < 5; ++dy) {
        const float* row = p + (size_t)min(max(y + dy - 2, 0), H - 1) * W;
        float jr = 0.f, wr = 0.f;
#pragma unroll
        for (int dx = 0; dx < 5; ++dx) {
            const float s = __ldg(row + min(max(x + dx - 2, 0), W - 1));
            const float d = __fsub_rn(s, I);
            const float f = __fmul_rn(__expf(__fmul_rn(__fmul_rn(d, d), nic)), gw[dy * 5 + dx]);
            S[dy * 5 + dx] = s;
            F[dy * 5 + dx] = f;
            jr = __fadd_rn(jr, __fmul_rn(f, s));
            wr = __fadd_rn(wr, f);
        }
        J = __fadd_rn(J, jr);
        Wt = __fadd_rn(Wt, wr);
    }
    const float den = __fadd_rn(Wt, 1e-5f);
    const float o = __fdiv_rn(J, den);
    const float a = gout[pl + (size_t)y * W + x] / den;
    const float k2 = 2.0f / var2_color;
    float own = 0.f;
    float* g = gin + pl;
#pragma unroll
    for (int dy = 0; dy < 5; ++dy) {
        const size_t ro = (size_t)min(max(y + dy - 2, 0), H - 1) * W;
#pragma unroll
        for (int dx = 0; dx < 5; ++dx) {
            const int i = dy * 5 + dx;
            const float d = S[i] - I;
            const float t = a * (S[i] - o) * F[i] * k2 * d;
            own += t;
            atomicAdd(g + ro + min(max(x + dx - 2, 0), W - 1), a * F[i] - t);
        }
    }
    atomicAdd(g + (size_t)y * W + x, own);
}

int launch_bilateral_vjp(const float* img, const float* gout, float* gin, int planes, int H, int W, float sigma_spatial,
                         float sigma_color, cudaStream_t stream) {
    if (planes > 65535) {
        set_error("B*C = %d exceeds the grid z limit", planes);
        return PB_ERR_ARG;
    }
    dim3 grid((W + 31) / 32, (H + 7) / 8, planes);
    ProfScope prof(PROF_OTHER, stream);
    PB_CUDA_TRY(cudaMemsetAsync(gin, 0, (size_t)planes * H * W * sizeof(float), stream));
    k_bilateral_vjp<<<grid, 256, 0, stream>>>(img, gout, gin, H, W, 2.0f * sigma_spatial * sigma_spatial,
                                             2.0f * sigma_color * sigma_color);
    PB_LAUNCH_CHECK("k_bilateral_vjp");
    return PB_OK;
}

int launch_bilateral(const float* img, float* out, int planes, int H, int W, float sigma_spatial,
                     float sigma_color, cudaStream_t stream) {
    if (planes > 65535) {
        set_error("B*C = %d exceeds the grid z limit", planes);
        return PB_ERR_ARG;
    }
    dim3 grid((W + 127) / 128, (H + 7) / 8, planes);
    ProfScope prof(PROF_OTHER, stream);
    k_bilateral<<<grid, 256, 0, stream>>>(img, out, H, W, 2.0f * sigma_spatial * sigma_spatial,
                                         2.0f * sigma_color * sigma_color);
    PB_LAUNCH_CHECK("k_bilateral");
    return PB_OK;
}

// ---------------------------------------------------------------------------------------------
// Domain-transform recursive filter (domain_transform.py:6-85).
//   k_rf_weights : Vh = a^(1 + s dIdx), Vv = a^(1 + s dIdy), dId* = L1 over channels of the forward
//                  differences of the joint image (0 at the first column / row)        (:27-38,55,59)
//   k_rf_rows    : per row and channel, left->right then right->left first-order recurrence (:78-83),
//                  one lane per row, tiles transposed through shared memory.
//   k_rf_cols    : the same recurrence down the columns, one thread per column (coalesced).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_rf_weights(const float* __restrict__ joint, float* __restrict__ Vh, float* __restrict__ Vv, int C, int H, int W,
             float log2a, float ratio) {
    const int x = blockIdx.x * 32 + (threadIdx.x & 31);
    const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= W || y >= H) return;
    const size_t plane = (size_t)H * W;
    const float* J = joint + (size_t)blockIdx.z * C * plane;
    float dx = 0.f, dy = 0.f;
    for (int c = 0; c < C; ++c) {
        const float v = __ldg(J + c * plane + (size_t)y * W + x);
        if (x > 0) dx = __fadd_rn(dx, fabsf(__fsub_rn(v, __ldg(J + c * plane + (size_t)y * W + x - 1))));
        if (y > 0) dy = __fadd_rn(dy, fabsf(__fsub_rn(v, __ldg(J + c * plane + (size_t)(y - 1) * W + x))));
    }
    const size_t o = (size_t)blockIdx.z * plane + (size_t)y * W + x;
    // a ^ e = 2 ^ (e log2 a), log2 a rounded from the double on the host: exp2f is 2 ulp, the product adds
    // |e log2 a| 2^-24 <~ 4e-7 relative for the exponents that occur -- against ~35 instructions of powf
    Vh[o] = exp2f(__fmul_rn(__fadd_rn(1.0f, __fmul_rn(ratio, dx)), log2a));
    Vv[o] = exp2f(__fmul_rn(__fadd_rn(1.0f, __fmul_rn(ratio, dy)), log2a));
}

// the same for widths that are a multiple of 4: a thread owns 4 adjacent pixels (128-bit loads and stores; the left
// neighbours of three of them are its own values), same operations in the same order per pixel
__global__ void __launch_bounds__(256)
k_rf_weights4(const float* __restrict__ joint, float* __restrict__ Vh, float* __restrict__ Vv, int C, int H, int W,
              float log2a, float ratio) {
    const int x = (blockIdx.x * 32 + (threadIdx.x & 31)) * 4;
    const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= W || y >= H) return;
    const size_t plane = (size_t)H * W;
    const float* J = joint + (size_t)blockIdx.z * C * plane + (size_t)y * W + x;
    float dx[4] = {0.f, 0.f, 0.f, 0.f}, dy[4] = {0.f, 0.f, 0.f, 0.f};
    for (int c = 0; c < C; ++c) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(J + c * plane));
        const float a[4] = {v.x, v.y, v.z, v.w};
        if (x > 0) dx[0] = __fadd_rn(dx[0], fabsf(__fsub_rn(a[0], __ldg(J + c * plane - 1))));
#pragma unroll
        for (int i = 1; i < 4; ++i) dx[i] = __fadd_rn(dx[i], fabsf(__fsub_rn(a[i], a[i - 1])));
        if (y > 0) {
            const float4 u = __ldg(reinterpret_cast<const float4*>(J + c * plane - W));
            const float b[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) dy[i] = __fadd_rn(dy[i], fabsf(__fsub_rn(a[i], b[i])));
        }
    }
    float h[4], w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        h[i] = exp2f(__fmul_rn(__fadd_rn(1.0f, __fmul_rn(ratio, dx[i])), log2a));
        w[i] = exp2f(__fmul_rn(__fadd_rn(1.0f, __fmul_rn(ratio, dy[i])), log2a));
    }
    const size_t o = (size_t)blockIdx.z * plane + (size_t)y * W + x;
    *reinterpret_cast<float4*>(Vh + o) = make_float4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<float4*>(Vv + o) = make_float4(w[0], w[1], w[2], w[3]);
}

// One warp = 32 consecutive rows of one plane, lane = row.  The rows are walked in tiles of 32
// columns: the tile of F and of V is transposed through shared memory (coalesced 128-byte global
// accesses, conflict-free 33-float pitch), every lane runs the reference's update on its row's 32
// samples with the carry in a register, and the tile is written back.  Left -> right sweep, then
// right -> left.  The recurrence is sequential along a row, so B*C*H lanes are all the parallelism there is
// (a few warps per SM): the tiles are double buffered with cp.async, the next tile's 64 loads per lane are in
// flight while the current one is filtered (was: four exposed load round trips per tile, 1.1 ms -> see profiles/).
#define RF_ROW_WARPS 4

__device__ __forceinline__ void rf_cp4(float* dst_smem, const float* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(dst_smem)), "l"(src)
                 : "memory");
}
__device__ __forceinline__ void rf_cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void rf_cp_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

__global__ void __launch_bounds__(RF_ROW_WARPS * 32)
k_rf_rows(const float* in, float* out, const float* __restrict__ Vh, int C, int H, int W,
          int groups_per_plane, int groups_total) {   // in may alias out (iterations >= 2)
    extern __shared__ float rf_sm[];                  // [warps][2 buffers][F, V][32][33]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int grp = blockIdx.x * RF_ROW_WARPS + warp;
    if (grp >= groups_total) return;
    const int pl = grp / groups_per_plane;               // image * C + channel
    const int y0 = (grp - pl * groups_per_plane) * 32;
    const size_t plane = (size_t)H * W;
    const float* f_in = in + (size_t)pl * plane;
    float* f_out = out + (size_t)pl * plane;
    const float* v = Vh + (size_t)(pl / C) * plane;
    float* base = rf_sm + (size_t)warp * 4 * 32 * 33;
    const int nrows = min(32, H - y0);
    const int ntiles = (W + 31) / 32;
    // asynchronous transposing load of tile t of `src` and of V into buffer `buf` (clamped at the edges)
    auto prefetch = [&](const float* src, int t, int buf) {
        float* F = base + (size_t)buf * 2 * 32 * 33;
        float* V = F + 32 * 33;
        const int x = min(t * 32 + lane, W - 1);
#pragma unroll 8
        for (int r = 0; r < 32; ++r) {
            const size_t o = (size_t)(y0 + min(r, nrows - 1)) * W + x;
            rf_cp4(F + r * 33 + lane, src + o);
            rf_cp4(V + r * 33 + lane, v + o);
        }
        rf_cp_commit();
    };
    // ---- left -> right:  F[x] += V[x] (F[x-1] - F[x]),  x >= 1
    float carry = 0.f;
    prefetch(f_in, 0, 0);
    for (int t = 0; t < ntiles; ++t) {
        const int x0 = t * 32;
        float* F = base + (size_t)(t & 1) * 2 * 32 * 33;
        float* V = F + 32 * 33;
        rf_cp_wait_all();
        __syncwarp();
        if (t + 1 < ntiles) prefetch(f_in, t + 1, (t + 1) & 1);
        if (lane < nrows) {
            const int n = min(32, W - x0);
#pragma unroll 8
            for (int i = 0; i < n; ++i) {
                float cur = F[lane * 33 + i];
                if (x0 + i > 0) cur = __fadd_rn(cur, __fmul_rn(V[lane * 33 + i], __fsub_rn(carry, cur)));
                F[lane * 33 + i] = cur;
                carry = cur;
            }
        }
        __syncwarp();
        for (int r = 0; r < nrows; ++r) {
            const int x = x0 + lane;
            if (x < W) f_out[(size_t)(y0 + r) * W + x] = F[r * 33 + lane];
        }
        __syncwarp();
    }
    // ---- right -> left:  F[x] += V[x+1] (F[x+1] - F[x]),  x <= W-2   (reads the sweep above back)
    float vnext = 0.f;                                    // V[x+1] of the sample just processed
    __threadfence_block();
    prefetch(f_out, ntiles - 1, (ntiles - 1) & 1);
    for (int t = ntiles - 1; t >= 0; --t) {
        const int x0 = t * 32;
        float* F = base + (size_t)(t & 1) * 2 * 32 * 33;
        float* V = F + 32 * 33;
        rf_cp_wait_all();
        __syncwarp();
        if (t > 0) prefetch(f_out, t - 1, (t - 1) & 1);
        if (lane < nrows) {
            const int n = min(32, W - x0);
#pragma unroll 8
            for (int i = n - 1; i >= 0; --i) {
                float cur = F[lane * 33 + i];
                if (x0 + i < W - 1) cur = __fadd_rn(cur, __fmul_rn(vnext, __fsub_rn(carry, cur)));
                F[lane * 33 + i] = cur;
                carry = cur;
                vnext = V[lane * 33 + i];
            }
        }
        __syncwarp();
        for (int r = 0; r < nrows; ++r) {
            const int x = x0 + lane;
            if (x < W) f_out[(size_t)(y0 + r) * W + x] = F[r * 33 + lane];
        }
        __syncwarp();
    }
}

// the same recurrence down the columns, one thread per column (coalesced); the loads of the next block of rows are
// in flight (registers) while the current block is filtered
__global__ void __launch_bounds__(128)
k_rf_cols(float* __restrict__ img, const float* __restrict__ Vv, int C, int H, int W) {
    const int x = blockIdx.x * 128 + threadIdx.x;
    if (x >= W) return;
    const int pl = blockIdx.y;                           // image * C + channel
    const size_t plane = (size_t)H * W;
    float* f = img + (size_t)pl * plane + x;
    const float* v = Vv + (size_t)(pl / C) * plane + x;
    constexpr int U = 8;
    float prev = f[0];
    float nf[U], nv[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const int y = min(1 + u, H - 1);
        nf[u] = f[(size_t)y * W];
        nv[u] = __ldg(v + (size_t)y * W);
    }
    for (int y0 = 1; y0 < H; y0 += U) {
        float cf[U], cv[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            cf[u] = nf[u];
            cv[u] = nv[u];
        }
        if (y0 + U < H) {
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int y = min(y0 + U + u, H - 1);
                nf[u] = f[(size_t)y * W];
                nv[u] = __ldg(v + (size_t)y * W);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (y0 + u < H) {
                const float cur = __fadd_rn(cf[u], __fmul_rn(cv[u], __fsub_rn(prev, cf[u])));
                f[(size_t)(y0 + u) * W] = cur;
                prev = cur;
            }
        }
    }
    // bottom -> top: F[y] += V[y+1] (F[y+1] - F[y])
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const int y = max(H - 2 - u, 0);
        nf[u] = f[(size_t)y * W];
        nv[u] = __ldg(v + (size_t)(y + 1) * W);
    }
    for (int y0 = H - 2; y0 >= 0; y0 -= U) {
        float cf[U], cv[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            cf[u] = nf[u];
            cv[u] = nv[u];
        }
        if (y0 - U >= 0) {
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int y = max(y0 - U - u, 0);
                nf[u] = f[(size_t)y * W];
                nv[u] = __ldg(v + (size_t)(y + 1) * W);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (y0 - u >= 0) {
                const float cur = __fadd_rn(cf[u], __fmul_rn(cv[u], __fsub_rn(prev, cf[u])));
                f[(size_t)(y0 - u) * W] = cur;
                prev = cur;
            }
        }
    }
}

size_t rf_workspace_bytes(int B, int H, int W) { return 2 * align_up((size_t)B * H * W * sizeof(float), 256); }

// in -> out (may not alias), joint = guide image or NULL (= in); ws holds Vh, Vv.
int launch_recursive_filter(const float* in, const float* joint, float* out, int B, int C, int H, int W,
                            double sigma_s, double sigma_r, int num_iterations, void* ws, cudaStream_t stream) {
    if (B * C > 65535 || B > 65535) {
        set_error("batch too large for the recursive filter grids");
        return PB_ERR_ARG;
    }
    const size_t plane = (size_t)H * W;
    float* Vh = static_cast<float*>(ws);
    float* Vv = reinterpret_cast<float*>(static_cast<char*>(ws) + align_up((size_t)B * plane * sizeof(float), 256));
    ProfScope prof(PROF_OTHER, stream);
    const float* cur = in;
    for (int i = 0; i < num_iterations; ++i) {
        // sigma_H_i and the feedback coefficient in Python doubles (domain_transform.py:50-53)
        const double sigma_i = sigma_s * sqrt(3.0) * pow(2.0, (double)(num_iterations - (i + 1))) /
                               sqrt(pow(4.0, (double)num_iterations) - 1.0);
        // log2 of the feedback coefficient exp(-sqrt(2) / sigma_i) as torch.pow sees it (a float32 base)
        const float a = (float)log2((double)(float)exp(-sqrt(2.0) / sigma_i));
        if ((W & 3) == 0) {
            dim3 gw((W / 4 + 31) / 32, (H + 7) / 8, B);
            k_rf_weights4<<<gw, 256, 0, stream>>>(joint ? joint : in, Vh, Vv, C, H, W, a, (float)(sigma_s / sigma_r));
        } else {
            dim3 gw((W + 31) / 32, (H + 7) / 8, B);
            k_rf_weights<<<gw, 256, 0, stream>>>(joint ? joint : in, Vh, Vv, C, H, W, a, (float)(sigma_s / sigma_r));
        }
        const int groups_per_plane = (H + 31) / 32;
        const int groups_total = B * C * groups_per_plane;
        const size_t rows_smem = (size_t)RF_ROW_WARPS * 4 * 32 * 33 * sizeof(float);
        PB_CUDA_TRY(cudaFuncSetAttribute(k_rf_rows, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rows_smem));
        k_rf_rows<<<(groups_total + RF_ROW_WARPS - 1) / RF_ROW_WARPS, RF_ROW_WARPS * 32, rows_smem, stream>>>(
            cur, out, Vh, C, H, W, groups_per_plane, groups_total);
        dim3 gc((W + 127) / 128, B * C);
        k_rf_cols<<<gc, 128, 0, stream>>>(out, Vv, C, H, W);
        cur = out;
    }
    PB_LAUNCH_CHECK("recursive filter");
    return PB_OK;
}

// ---------------------------------------------------------------------------------------------
// Elementwise epilogues.
// ---------------------------------------------------------------------------------------------
// impred = clip(deconv(smooth) + (impred - smooth), 0, 1)          (deblurring.py:80-88)
__global__ void k_residual_add4(float4* __restrict__ dec, const float4* __restrict__ cur, const float4* __restrict__ smooth,
                                size_t n4) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n4) {
        const float4 c = __ldg(cur + i), s = __ldg(smooth + i), d = dec[i];
        dec[i] = make_float4(fminf(fmaxf(__fadd_rn(d.x, __fsub_rn(c.x, s.x)), 0.0f), 1.0f),
                             fminf(fmaxf(__fadd_rn(d.y, __fsub_rn(c.y, s.y)), 0.0f), 1.0f),
                             fminf(fmaxf(__fadd_rn(d.z, __fsub_rn(c.z, s.z)), 0.0f), 1.0f),
                             fminf(fmaxf(__fadd_rn(d.w, __fsub_rn(c.w, s.w)), 0.0f), 1.0f));
    }
}
__global__ void k_residual_add(float* __restrict__ dec, const float* __restrict__ cur, const float* __restrict__ smooth,
                               size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const float noise = __fsub_rn(cur[i], smooth[i]);
        dec[i] = fminf(fmaxf(__fadd_rn(dec[i], noise), 0.0f), 1.0f);
    }
}

int launch_residual_add(float* dec, const float* cur, const float* smooth, size_t n, cudaStream_t stream) {
    ProfScope prof(PROF_OTHER, stream);
    if ((n & 3) == 0 && ((((uintptr_t)dec | (uintptr_t)cur | (uintptr_t)smooth) & 15) == 0)) {
        k_residual_add4<<<(unsigned)((n / 4 + 255) / 256), 256, 0, stream>>>(
            reinterpret_cast<float4*>(dec), reinterpret_cast<const float4*>(cur), reinterpret_cast<const float4*>(smooth), n / 4);
    } else {
        k_residual_add<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(dec, cur, smooth, n);
    }
    PB_LAUNCH_CHECK("k_residual_add");
    return PB_OK;
}

// nM[plane] = sum over pixels of gx^2 + gy^2, deterministic two-stage sum (deblurring.py:181,204)
#define HALO_PARTS 64
__global__ void __launch_bounds__(256)
k_halo_norm_partial(const float* __restrict__ gx, const float* __restrict__ gy, float* __restrict__ partial,
                    size_t plane) {
    __shared__ float red[8];
    const float* a = gx + (size_t)blockIdx.y * plane;
    const float* b = gy + (size_t)blockIdx.y * plane;
    const size_t per = (plane + HALO_PARTS - 1) / HALO_PARTS;
    const size_t i0 = (size_t)blockIdx.x * per, i1 = i0 + per < plane ? i0 + per : plane;
    float s = 0.f;
    for (size_t i = i0 + threadIdx.x; i < i1; i += 256) s += a[i] * a[i] + b[i] * b[i];
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int w = 0; w < 8; ++w) t += red[w];
        partial[blockIdx.y * HALO_PARTS + blockIdx.x] = t;
    }
}
__global__ void k_halo_norm_final(const float* __restrict__ partial, float* __restrict__ nM, int planes) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < planes) {
        float t = 0.f;
        for (int i = 0; i < HALO_PARTS; ++i) t += partial[p * HALO_PARTS + i];
        nM[p] = t;
    }
}

int launch_halo_norm(const float* gx, const float* gy, float* partial, float* nM, int planes, size_t plane,
                     cudaStream_t stream) {
    if (planes > 65535) {
        set_error("B*C = %d exceeds the grid y limit", planes);
        return PB_ERR_ARG;
    }
    ProfScope prof(PROF_OTHER, stream);
    k_halo_norm_partial<<<dim3(HALO_PARTS, planes), 256, 0, stream>>>(gx, gy, partial, plane);
    k_halo_norm_final<<<(planes + 127) / 128, 128, 0, stream>>>(partial, nM, planes);
    PB_LAUNCH_CHECK("k_halo_norm");
    return PB_OK;
}

// out = clamp(imout + max(M / (nM + M), 0) (img - imout)),  M = -gx ox - gy gy  (sic, deblurring.py:174)
// `img` is the (possibly tapered) current image: plane pitch / row pitch / offset given explicitly.
__global__ void __launch_bounds__(256)
k_halo_apply(float* __restrict__ imout, const float* __restrict__ img, size_t img_plane, int img_pitch, int img_off,
             const float* __restrict__ gx, const float* __restrict__ gy, const float* __restrict__ ox,
             const float* __restrict__ nM, int H, int W) {
    const int x = blockIdx.x * 32 + (threadIdx.x & 31);
    const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= W || y >= H) return;
    const size_t i = (size_t)blockIdx.z * H * W + (size_t)y * W + x;
    const float cur = img[(size_t)blockIdx.z * img_plane + (size_t)(y + img_off) * img_pitch + x + img_off];
    const float M = __fadd_rn(__fmul_rn(-gx[i], ox[i]), __fmul_rn(-gy[i], gy[i]));
    const float z = fmaxf(__fdiv_rn(M, __fadd_rn(nM[blockIdx.z], M)), 0.0f);
    const float o = imout[i];
    imout[i] = fminf(fmaxf(__fadd_rn(o, __fmul_rn(z, __fsub_rn(cur, o))), 0.0f), 1.0f);
}

int launch_halo_apply(float* imout, const float* img, size_t img_plane, int img_pitch, int img_off, const float* gx,
                      const float* gy, const float* ox, const float* nM, int planes, int H, int W,
                      cudaStream_t stream) {
    dim3 grid((W + 31) / 32, (H + 7) / 8, planes);
    ProfScope prof(PROF_OTHER, stream);
    k_halo_apply<<<grid, 256, 0, stream>>>(imout, img, img_plane, img_pitch, img_off, gx, gy, ox, nM, H, W);
    PB_LAUNCH_CHECK("k_halo_apply");
    return PB_OK;
}

// ---------------------------------------------------------------------------------------------
// Edgetaper (edgetaper.py:10-33) on the replicate-padded image, method='fft' semantics:
//   alpha = v1 (x) v2,  v = 1 - z / max z,  z = circular autocorrelation (period n - 1) of the
//   kernel's row / column sums, last sample repeated;  3 x  p <- alpha p + (1 - alpha) (K (*) p)
//   with the circular correlation on the padded torus.  alpha = 1 away from the border band, so
//   the 625-tap sum is only evaluated where alpha < 1.
// ---------------------------------------------------------------------------------------------
__global__ void k_pad_replicate(const float* __restrict__ img, float* __restrict__ a, float* __restrict__ b, int H,
                                int W, int pad) {
    const int Wp = W + 2 * pad, Hp = H + 2 * pad;
    const int x = blockIdx.x * 32 + (threadIdx.x & 31);
    const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= Wp || y >= Hp) return;
    const float v = __ldg(img + (size_t)blockIdx.z * H * W + (size_t)min(max(y - pad, 0), H - 1) * W +
                          min(max(x - pad, 0), W - 1));
    const size_t o = (size_t)blockIdx.z * Hp * Wp + (size_t)y * Wp + x;
    a[o] = v;
    b[o] = v;
}

// per image: zl[0..48] = linear autocorrelation lags -24..24 of the row sums, zl[49..97] of the column
// sums; zmax[2 im], zmax[2 im + 1] = lag-0 values (the maxima).
__global__ void __launch_bounds__(128)
k_et_autocorr(const ImgKernel* __restrict__ kern, float* __restrict__ zl, float* __restrict__ zmax, int Hp, int Wp) {
    __shared__ float rh[PB_KS], rw[PB_KS];
    const ImgKernel* K = kern + blockIdx.x;
    const int t = threadIdx.x;
    if (t < PB_KS) {
        float s = 0.f, u = 0.f;
        for (int i = 0; i < PB_KS; ++i) {
            s += K->k[t * PB_KS + i];       // sum over columns: projection indexed by row
            u += K->k[i * PB_KS + t];       // sum over rows: projection indexed by column
        }
        // torch.fft.fft(x, n) TRUNCATES x when n < len(x) (edgetaper.py:11,17 with n = size - 1): only
        // matters for images one pixel high / wide, where the padded size minus one is below ker_size
        const int first = PB_PAD - K->ksize / 2;        // the ksize x ksize kernel sits centred in the 25 x 25 grid
        rh[t] = (t - first < Hp - 1) ? s : 0.f;
        rw[t] = (t - first < Wp - 1) ? u : 0.f;
    }
    __syncthreads();
    if (t < 2 * (2 * PB_KS - 1)) {
        const int which = t / (2 * PB_KS - 1);
        const int d = t % (2 * PB_KS - 1) - (PB_KS - 1);
        const float* r = which ? rw : rh;
        float s = 0.f;
        for (int i = 0; i < PB_KS; ++i) {
            const int j = i + d;
            if (j >= 0 && j < PB_KS) s += r[i] * r[j];
        }
        zl[(size_t)blockIdx.x * 98 + t] = s;
        if (d == 0) zmax[2 * blockIdx.x + which] = s;
    }
}

// v1[im][y], y < Hp and v2[im][x], x < Wp (stored back to back: Hp + Wp floats per image)
__global__ void __launch_bounds__(256)
k_et_weights(const float* __restrict__ zl, const float* __restrict__ zmax, float* __restrict__ v, int B, int Hp,
             int Wp, int batch_max) {
    const int im = blockIdx.y;
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= Hp + Wp) return;
    const int which = i >= Hp;
    const int n = which ? Wp : Hp;
    int m = which ? i - Hp : i;
    if (m == n - 1) m = 0;                                  // z = cat(z, z[0])
    const int period = n - 1;
    float z = 0.f;
    if (period >= 1) {
        for (int d = -(PB_KS - 1); d <= PB_KS - 1; ++d)
            if (pmod(d - m, period) == 0) z += zl[(size_t)im * 98 + which * 49 + d + PB_KS - 1];
    }
    float mx = zmax[2 * im + which];
    if (batch_max)
        for (int b = 0; b < B; ++b) mx = fmaxf(mx, zmax[2 * b + which]);
    v[(size_t)im * (Hp + Wp) + i] = __fsub_rn(1.0f, __fdiv_rn(z, mx));
}

__global__ void __launch_bounds__(256)
k_et_pass(const float* __restrict__ src, float* __restrict__ dst, const ImgKernel* __restrict__ kern,
          const float* __restrict__ v, int C, int Hp, int Wp) {
    const int x = blockIdx.x * 32 + (threadIdx.x & 31);
    const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= Wp || y >= Hp) return;
    const int im = blockIdx.z / C;
    const float* vv = v + (size_t)im * (Hp + Wp);
    const float alpha = __fmul_rn(vv[y], vv[Hp + x]);
    const size_t pl = (size_t)blockIdx.z * Hp * Wp;
    const float p = src[pl + (size_t)y * Wp + x];
    float o = p;
    if (alpha != 1.0f) {
        const ImgKernel* K = kern + im;
        float acc = 0.f;
        for (int dy = -PB_PAD; dy <= PB_PAD; ++dy) {
            const int lo = K->lo[dy + PB_PAD], hi = K->hi[dy + PB_PAD];
            if (lo > hi) continue;
            const float* row = src + pl + (size_t)pmod(y + dy, Hp) * Wp;
            float r = 0.f;
            for (int t = lo; t <= hi; ++t) r = fmaf(K->k[(dy + PB_PAD) * PB_KS + t], row[pmod(x + t - PB_PAD, Wp)], r);
            acc += r;
        }
        o = __fadd_rn(__fmul_rn(alpha, p), __fmul_rn(__fsub_rn(1.0f, alpha), acc));
    }
    dst[pl + (size_t)y * Wp + x] = o;
}

size_t edgetaper_scratch_bytes(int B, int Hp, int Wp) {
    return align_up((size_t)B * 98 * sizeof(float), 256) + align_up((size_t)B * 2 * sizeof(float), 256) +
           align_up((size_t)B * (Hp + Wp) * sizeof(float), 256);
}

// Builds the taper weights of every image (needs the kernel records) -> v in scratch.
int launch_edgetaper_weights(const ImgKernel* kern, void* scratch, int B, int Hp, int Wp, int batch_max,
                             float** v_out, cudaStream_t stream) {
    char* s = static_cast<char*>(scratch);
    float* zl = reinterpret_cast<float*>(s);
    float* zmax = reinterpret_cast<float*>(s + align_up((size_t)B * 98 * sizeof(float), 256));
    float* v = reinterpret_cast<float*>(s + align_up((size_t)B * 98 * sizeof(float), 256) +
                                        align_up((size_t)B * 2 * sizeof(float), 256));
    if (B > 65535) {
        set_error("batch too large for the edgetaper grids");
        return PB_ERR_ARG;
    }
    ProfScope prof(PROF_OTHER, stream);
    k_et_autocorr<<<B, 128, 0, stream>>>(kern, zl, zmax, Hp, Wp);
    k_et_weights<<<dim3((Hp + Wp + 255) / 256, B), 256, 0, stream>>>(zl, zmax, v, B, Hp, Wp, batch_max);
    PB_LAUNCH_CHECK("edgetaper weights");
    *v_out = v;
    return PB_OK;
}

int launch_pad_replicate(const float* img, float* a, float* b, int planes, int H, int W, int pad,
                         cudaStream_t stream) {
    if (planes > 65535) {
        set_error("B*C = %d exceeds the grid z limit", planes);
        return PB_ERR_ARG;
    }
    dim3 grid((W + 2 * pad + 31) / 32, (H + 2 * pad + 7) / 8, planes);
    ProfScope prof(PROF_OTHER, stream);
    k_pad_replicate<<<grid, 256, 0, stream>>>(img, a, b, H, W, pad);
    PB_LAUNCH_CHECK("k_pad_replicate");
    return PB_OK;
}

// utils.crop_with_kernel (utils.py:56-61) of a padded stack into a dense one
__global__ void __launch_bounds__(256)
k_crop(const float* __restrict__ padded, float* __restrict__ out, int H, int W, int pad) {
    const int x = blockIdx.x * 32 + (threadIdx.x & 31);
    const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= W || y >= H) return;
    const size_t pl = blockIdx.z;
    const int Wp = W + 2 * pad, Hp = H + 2 * pad;
    out[pl * (size_t)H * W + (size_t)y * W + x] = padded[pl * (size_t)Hp * Wp + (size_t)(y + pad) * Wp + x + pad];
}

int launch_crop(const float* padded, float* out, int planes, int H, int W, int pad, cudaStream_t stream) {
    if (planes > 65535) {
        set_error("B*C = %d exceeds the grid z limit", planes);
        return PB_ERR_ARG;
    }
    dim3 grid((W + 31) / 32, (H + 7) / 8, planes);
    ProfScope prof(PROF_OTHER, stream);
    k_crop<<<grid, 256, 0, stream>>>(padded, out, H, W, pad);
    PB_LAUNCH_CHECK("k_crop");
    return PB_OK;
}

// n_tapers passes ping-ponging between a and b (both hold the padded image on entry); returns the
// buffer that holds the result.
int launch_edgetaper_passes(float* a, float* b, const ImgKernel* kern, const float* v, int B, int C, int Hp, int Wp,
                            int n_tapers, float** result, cudaStream_t stream) {
    dim3 grid((Wp + 31) / 32, (Hp + 7) / 8, B * C);
    ProfScope prof(PROF_OTHER, stream);
    float *src = a, *dst = b;
    for (int i = 0; i < n_tapers; ++i) {
        k_et_pass<<<grid, 256, 0, stream>>>(src, dst, kern, v, C, Hp, Wp);
        float* t = src;
        src = dst;
        dst = t;
    }
    PB_LAUNCH_CHECK("k_et_pass");
    *result = src;
    return PB_OK;
}

// ---- vector-Jacobian product of the deconvolution with respect to the image -------------------
// Forward (deblurring.py:211-239, default flags): y = clamp(C T R x), R = replicate pad by P
// (utils.py:48-53), T = circular convolution with the polynomial of the blur on the (H+2P) x (W+2P)
// torus, C = crop (utils.py:56-61).  What autograd gives for it: x~ = R^T T~ C^T (y~ . pass), pass = 1
// where the unclamped value lies in [0, 1], T~ = the same filter with the kernel rotated by 180 degrees.
// k_vjp_embed is C^T with the clamp mask, k_vjp_fold is R^T; T~ runs on the ordinary engines, which
// read the embedded plane as a pure wrap-around torus (SrcGeom pad = 0).
__global__ void k_vjp_embed(const float* __restrict__ gout, const float* __restrict__ preclamp,
                            float* __restrict__ z, int H, int W, int pad) {
    const int Hp = H + 2 * pad, Wp = W + 2 * pad;
    const int xp = blockIdx.x * 32 + (threadIdx.x & 31);
    const int yp = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (xp >= Wp || yp >= Hp) return;
    const size_t pl = blockIdx.z;
    const int x = xp - pad, y = yp - pad;
    float v = 0.0f;
    if (x >= 0 && x < W && y >= 0 && y < H) {
        const size_t o = pl * (size_t)H * W + (size_t)y * W + x;
        v = gout[o];
        if (preclamp) {
            const float u = preclamp[o];
            if (!(u >= 0.0f && u <= 1.0f)) v = 0.0f;
        }
    }
    z[pl * (size_t)Hp * Wp + (size_t)yp * Wp + xp] = v;
}

__global__ void k_vjp_fold(const float* __restrict__ t, float* __restrict__ gin, int H, int W, int pad) {
    const int Hp = H + 2 * pad, Wp = W + 2 * pad;
    const int x = blockIdx.x * 32 + (threadIdx.x & 31);
    const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= W || y >= H) return;
    const size_t pl = blockIdx.z;
    // padded rows / columns that replicate this pixel
    const int y0 = (y == 0) ? 0 : y + pad, y1 = (y == H - 1) ? Hp - 1 : y + pad;
    const int x0 = (x == 0) ? 0 : x + pad, x1 = (x == W - 1) ? Wp - 1 : x + pad;
    const float* tp = t + pl * (size_t)Hp * Wp;
    float acc = 0.0f;
    for (int yy = y0; yy <= y1; ++yy)
        for (int xx = x0; xx <= x1; ++xx) acc += tp[(size_t)yy * Wp + xx];
    gin[pl * (size_t)H * W + (size_t)y * W + x] = acc;
}

// kernel rotated by 180 degrees (the adjoint of a correlation-free convolution)
__global__ void k_flip_kernels(const float* __restrict__ k, float* __restrict__ kf, int n_per, int total) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int b = i / n_per, e = i - b * n_per;
    kf[i] = k[(size_t)b * n_per + (n_per - 1 - e)];
}

int launch_vjp_embed(const float* gout, const float* preclamp, float* z, int planes, int H, int W, int pad,
                     cudaStream_t stream) {
    if (planes > 65535) {
        set_error("B*C = %d exceeds the grid z limit", planes);
        return PB_ERR_ARG;
    }
    dim3 grid((W + 2 * pad + 31) / 32, (H + 2 * pad + 7) / 8, planes);
    ProfScope prof(PROF_OTHER, stream);
    k_vjp_embed<<<grid, 256, 0, stream>>>(gout, preclamp, z, H, W, pad);
    PB_LAUNCH_CHECK("k_vjp_embed");
    return PB_OK;
}

int launch_vjp_fold(const float* t, float* gin, int planes, int H, int W, int pad, cudaStream_t stream) {
    dim3 grid((W + 31) / 32, (H + 7) / 8, planes);
    ProfScope prof(PROF_OTHER, stream);
    k_vjp_fold<<<grid, 256, 0, stream>>>(t, gin, H, W, pad);
    PB_LAUNCH_CHECK("k_vjp_fold");
    return PB_OK;
}

int launch_flip_kernels(const float* k, float* kf, int B, int ksize, cudaStream_t stream) {
    const int total = B * ksize * ksize;
    ProfScope prof(PROF_OTHER, stream);
    k_flip_kernels<<<(total + 255) / 256, 256, 0, stream>>>(k, kf, ksize * ksize, total);
    PB_LAUNCH_CHECK("k_flip_kernels");
    return PB_OK;
}

}  // namespace pb
