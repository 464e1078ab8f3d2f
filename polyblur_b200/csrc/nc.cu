// Domain-transform normalized convolution (the reference's native prototype,
// polyblur/domain_transform/NC.cpp:143-204 and :50-140):
//   ctH = cumsum_x (1 + s/r * dIdx),  ctV = cumsum_y (1 + s/r * dIdy)     (float32, sequential)
//   per iteration i: box radius R = sqrt(3) sigma_i in the transformed domain;
//     rows:    l[x] = #{ct <= ct[x] - R},  u[x] = #{ct <= ct[x] + R},
//              F'[x] = (SAT[u] - SAT[l]) / ((u - l) + 1e-4),  SAT = [0, cumsum F]
//     columns: the same on the transposed image with ctV.
// The running sums are kept sequential in fp32 (one lane walks the row in shared memory) so that the
// window indices, which compare sums that differ by less than an ulp of a parallel scan, come out as
// in the reference; the searches and the box averages run on all lanes.
#include "kernels.cuh"

namespace pb {

// ctH[b][y][x] (row-wise running sum) -- one warp per row
#define NC_WARPS 4
__global__ void __launch_bounds__(NC_WARPS * 32)
k_nc_ct_rows(const float* __restrict__ img, float* __restrict__ ct, int C, int H, int W, float ratio,
             int rows_total) {
    extern __shared__ float ncs[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row = blockIdx.x * NC_WARPS + warp;
    if (row >= rows_total) return;
    float* d = ncs + (size_t)warp * W;
    const int b = row / H, y = row - b * H;
    const size_t plane = (size_t)H * W;
    const float* I = img + (size_t)b * C * plane + (size_t)y * W;
    for (int x = lane; x < W; x += 32) {
        float s = 0.f;
        if (x > 0)
            for (int c = 0; c < C; ++c) s = __fadd_rn(s, fabsf(__fsub_rn(__ldg(I + c * plane + x), __ldg(I + c * plane + x - 1))));
        d[x] = __fadd_rn(1.0f, __fmul_rn(ratio, s));
    }
    __syncwarp();
    if (lane == 0) {
        float acc = 0.f;
        for (int x = 0; x < W; ++x) {
            acc = __fadd_rn(acc, d[x]);
            d[x] = acc;
        }
    }
    __syncwarp();
    float* o = ct + (size_t)b * plane + (size_t)y * W;
    for (int x = lane; x < W; x += 32) o[x] = d[x];
}

// ctVT[b][x][y] (column-wise running sum, stored transposed) -- one thread per column
__global__ void __launch_bounds__(128)
k_nc_ct_cols(const float* __restrict__ img, float* __restrict__ ctT, int C, int H, int W, float ratio) {
    const int x = blockIdx.x * 128 + threadIdx.x;
    if (x >= W) return;
    const int b = blockIdx.y;
    const size_t plane = (size_t)H * W;
    const float* I = img + (size_t)b * C * plane + x;
    float* o = ctT + (size_t)b * plane + (size_t)x * H;
    float acc = 0.f;
    for (int y = 0; y < H; ++y) {
        float s = 0.f;
        if (y > 0)
            for (int c = 0; c < C; ++c)
                s = __fadd_rn(s, fabsf(__fsub_rn(__ldg(I + c * plane + (size_t)y * W), __ldg(I + c * plane + (size_t)(y - 1) * W))));
        acc = __fadd_rn(acc, __fadd_rn(1.0f, __fmul_rn(ratio, s)));
        o[y] = acc;
    }
}

__global__ void k_transpose(const float* __restrict__ in, float* __restrict__ out, int H, int W) {
    __shared__ float tile[32][33];
    const size_t pl = (size_t)blockIdx.z * H * W;
    int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 32 + threadIdx.y;
    for (int j = 0; j < 32; j += 8)
        if (x < W && y + j < H) tile[threadIdx.y + j][threadIdx.x] = in[pl + (size_t)(y + j) * W + x];
    __syncthreads();
    x = blockIdx.y * 32 + threadIdx.x;
    y = blockIdx.x * 32 + threadIdx.y;
    for (int j = 0; j < 32; j += 8)
        if (x < H && y + j < W) out[pl + (size_t)(y + j) * H + x] = tile[threadIdx.x][threadIdx.y + j];
}

// first index with ct[idx] > v  (= torch.searchsorted(ct, v, right=True))
__device__ __forceinline__ int upper_bound(const float* ct, int n, float v) {
    int lo = 0, hi = n;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (ct[mid] <= v) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// box filter of every row of length n of (planes x rows x n) with the row's ct -- one warp per (image, row)
__global__ void __launch_bounds__(NC_WARPS * 32)
k_nc_box_rows(const float* __restrict__ in, float* __restrict__ out, const float* __restrict__ ct, int C, int rows,
              int n, float radius, int rows_total) {
    extern __shared__ float ncs[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row = blockIdx.x * NC_WARPS + warp;
    if (row >= rows_total) return;
    float* cts = ncs + (size_t)warp * (4 * n + 4);
    float* sat = cts + n;                       // n + 1 entries
    int* lo = reinterpret_cast<int*>(sat + n + 1);
    int* hi = lo + n;
    const int b = row / rows, y = row - b * rows;
    const size_t plane = (size_t)rows * n;
    const float* c0 = ct + (size_t)b * plane + (size_t)y * n;
    for (int x = lane; x < n; x += 32) cts[x] = __ldg(c0 + x);
    __syncwarp();
    for (int x = lane; x < n; x += 32) {
        lo[x] = upper_bound(cts, n, __fsub_rn(cts[x], radius));
        hi[x] = upper_bound(cts, n, __fadd_rn(cts[x], radius));
    }
    for (int c = 0; c < C; ++c) {
        const float* f = in + ((size_t)b * C + c) * plane + (size_t)y * n;
        float* o = out + ((size_t)b * C + c) * plane + (size_t)y * n;
        __syncwarp();
        for (int x = lane; x < n; x += 32) sat[x + 1] = __ldg(f + x);
        __syncwarp();
        if (lane == 0) {
            float acc = 0.f;
            sat[0] = 0.f;
            for (int x = 1; x <= n; ++x) {
                acc = __fadd_rn(acc, sat[x]);
                sat[x] = acc;
            }
        }
        __syncwarp();
        for (int x = lane; x < n; x += 32) {
            const int l = lo[x], u = hi[x];
            o[x] = __fdiv_rn(__fsub_rn(sat[u], sat[l]), __fadd_rn((float)(u - l), 1e-4f));
        }
    }
}

size_t nc_workspace_bytes(int B, int C, int H, int W) {
    const size_t plane = (size_t)H * W * sizeof(float);
    return 2 * align_up((size_t)B * plane, 256) + 2 * align_up((size_t)B * C * plane, 256);
}

int launch_normalized_convolution(const float* img, float* out, int B, int C, int H, int W, double sigma_s,
                                  double sigma_r, int num_iterations, void* ws, cudaStream_t stream) {
    if (B * C > 65535) {
        set_error("batch too large for the normalized convolution grids");
        return PB_ERR_ARG;
    }
    const int nmax = H > W ? H : W;
    const size_t smem_box = (size_t)NC_WARPS * (4 * nmax + 4) * sizeof(float);
    const size_t smem_ct = (size_t)NC_WARPS * W * sizeof(float);
    if (smem_box > PB_SMEM_MAX - 1024) {
        set_error("image side %d does not fit the normalized convolution's shared memory", nmax);
        return PB_ERR_UNSUPPORTED;
    }
    PB_CUDA_TRY(cudaFuncSetAttribute(k_nc_box_rows, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_box));
    PB_CUDA_TRY(cudaFuncSetAttribute(k_nc_ct_rows, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_ct));
    const size_t plane = (size_t)H * W;
    char* base = static_cast<char*>(ws);
    float* ctH = reinterpret_cast<float*>(base);
    float* ctVT = reinterpret_cast<float*>(base + align_up((size_t)B * plane * sizeof(float), 256));
    float* tA = reinterpret_cast<float*>(base + 2 * align_up((size_t)B * plane * sizeof(float), 256));
    float* tB = reinterpret_cast<float*>(base + 2 * align_up((size_t)B * plane * sizeof(float), 256) +
                                         align_up((size_t)B * C * plane * sizeof(float), 256));
    ProfScope prof(PROF_OTHER, stream);
    const float ratio = (float)(sigma_s / sigma_r);
    k_nc_ct_rows<<<(B * H + NC_WARPS - 1) / NC_WARPS, NC_WARPS * 32, smem_ct, stream>>>(img, ctH, C, H, W, ratio, B * H);
    k_nc_ct_cols<<<dim3((W + 127) / 128, B), 128, 0, stream>>>(img, ctVT, C, H, W, ratio);
    const float* cur = img;
    for (int i = 0; i < num_iterations; ++i) {
        const double sigma_i = sigma_s * sqrt(3.0) * pow(2.0, (double)(num_iterations - (i + 1))) /
                               sqrt(pow(4.0, (double)num_iterations) - 1.0);
        const float radius = (float)(sqrt(3.0) * sigma_i);
        // rows: cur (B,C,H,W) -> tA
        k_nc_box_rows<<<(B * H + NC_WARPS - 1) / NC_WARPS, NC_WARPS * 32, smem_box, stream>>>(cur, tA, ctH, C, H, W,
                                                                                              radius, B * H);
        // transpose -> tB (B,C,W,H); columns as rows of the transposed image -> tA; transpose back -> out
        k_transpose<<<dim3((W + 31) / 32, (H + 31) / 32, B * C), dim3(32, 8), 0, stream>>>(tA, tB, H, W);
        k_nc_box_rows<<<(B * W + NC_WARPS - 1) / NC_WARPS, NC_WARPS * 32, smem_box, stream>>>(tB, tA, ctVT, C, W, H,
                                                                                              radius, B * W);
        k_transpose<<<dim3((H + 31) / 32, (W + 31) / 32, B * C), dim3(32, 8), 0, stream>>>(tA, out, W, H);
        cur = out;
    }
    PB_LAUNCH_CHECK("normalized convolution");
    return PB_OK;
}

}  // namespace pb
