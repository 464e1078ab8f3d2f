// Micro-benchmark: scalar FADD/FFMA against the packed FADD2/FFMA2 forms of sm_100 on complex
// (float2) butterfly-like dependency chains.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float2 a) { return *reinterpret_cast<u64*>(&a); }
__device__ __forceinline__ float2 up(u64 a) { return *reinterpret_cast<float2*>(&a); }
__device__ __forceinline__ float2 add2(float2 a, float2 b) { u64 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(pk(a)), "l"(pk(b))); return up(d); }
__device__ __forceinline__ float2 sub2(float2 a, float2 b) { u64 d; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(pk(a)), "l"(pk(b))); return up(d); }
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) { u64 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(pk(a)), "l"(pk(b)), "l"(pk(c))); return up(d); }
__device__ __forceinline__ float2 sadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 ssub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 sfma(float2 a, float2 b, float2 c) { return make_float2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y)); }

template <int MODE>
__global__ void k(float2* out, int iters, float2 seed) {
    float2 v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = make_float2(seed.x + i + threadIdx.x, seed.y - i);
    const float2 w = make_float2(0.999f, 1.001f);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {             // radix-2 style butterflies on 8 independent values
            float2 a = v[i], b = v[i + 4];
            if (MODE == 0) { v[i] = sadd(a, b); v[i + 4] = ssub(a, b); }
            if (MODE == 1) { v[i] = add2(a, b); v[i + 4] = sub2(a, b); }
            if (MODE == 2) { v[i] = sfma(a, w, b); v[i + 4] = sfma(b, w, a); }
            if (MODE == 3) { v[i] = fma2(a, w, b); v[i + 4] = fma2(b, w, a); }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = make_float2(v[i].x * 0.5f, v[i].y * 0.5f);   // keep finite (scalar FMULs in all modes)
    }
    float2 s = v[0];
#pragma unroll
    for (int i = 1; i < 8; ++i) s = sadd(s, v[i]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
float run(float2* out, int iters) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    k<MODE><<<148 * 8, 256>>>(out, iters, make_float2(1.f, 2.f));
    cudaEventRecord(a);
    k<MODE><<<148 * 8, 256>>>(out, iters, make_float2(1.f, 2.f));
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); return ms;
}
int main() {
    float2* out; cudaMalloc(&out, 148 * 8 * 256 * sizeof(float2));
    const int iters = 20000;
    const double lanes = 148.0 * 8 * 256, ops = lanes * iters * 8.0 * 2.0;   // float ops of the butterfly part per launch
    const char* names[4] = {"scalar FADD", "packed FADD2", "scalar FFMA", "packed FFMA2"};
    float ms[4] = {run<0>(out, iters), run<1>(out, iters), run<2>(out, iters), run<3>(out, iters)};
    for (int m = 0; m < 4; ++m) printf("%-13s %8.3f ms   %.2f Tflop-lanes/s (butterfly part; + 16 scalar FMUL per iteration in every mode)\n", names[m], ms[m], ops / ms[m] / 1e9);
    return 0;
}
