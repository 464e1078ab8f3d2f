// 8-bit image I/O on the device: what main.py does around the hot path --
// img_as_float32 on the decoded image (main.py:80) and img_as_ubyte on the result (main.py:146) --
// fused with the HWC <-> NCHW layout change, so that only one byte per sample crosses PCIe.
#include "kernels.cuh"

namespace pb {

// (B,H,W,C) uint8 -> (B,C,H,W) float32 in [0,1]  (x / 255, like polyblur_b200.utils.to_float)
__global__ void __launch_bounds__(256)
k_u8hwc_to_f32nchw(const unsigned char* __restrict__ in, float* __restrict__ out, int C, size_t plane, size_t total) {
    const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;       // pixel index over B*H*W
    if (i >= total) return;
    const size_t b = i / plane, p = i - b * plane;
    const unsigned char* s = in + i * C;
    float* d = out + b * C * plane + p;
    for (int c = 0; c < C; ++c) d[(size_t)c * plane] = __fdiv_rn((float)s[c], 255.0f);
}

// (B,C,H,W) float32 -> (B,H,W,C) uint8: rint(clip(x, 0, 1) * 255)  (main.py:146 img_as_ubyte = utils.to_ubyte)
__global__ void __launch_bounds__(256)
k_f32nchw_to_u8hwc(const float* __restrict__ in, unsigned char* __restrict__ out, int C, size_t plane, size_t total) {
    const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= total) return;
    const size_t b = i / plane, p = i - b * plane;
    const float* s = in + b * C * plane + p;
    unsigned char* d = out + i * C;
    for (int c = 0; c < C; ++c) {
        const float v = fminf(fmaxf(__ldg(s + (size_t)c * plane), 0.0f), 1.0f);
        d[c] = (unsigned char)__float2int_rn(__fmul_rn(v, 255.0f));
    }
}

// RGB fast path: a thread converts 4 pixels = 12 bytes (three aligned 32-bit words) <-> one float4 per plane
__global__ void __launch_bounds__(256)
k_u8hwc_to_f32nchw_rgb4(const unsigned* __restrict__ in, float* __restrict__ out, size_t plane4, size_t total4) {
    const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;       // group of 4 pixels over B*H*W/4
    if (i >= total4) return;
    const size_t b = i / plane4, p = i - b * plane4;
    const unsigned w0 = __ldg(in + 3 * i), w1 = __ldg(in + 3 * i + 1), w2 = __ldg(in + 3 * i + 2);
    // bytes: r0 g0 b0 r1 | g1 b1 r2 g2 | b2 r3 g3 b3
    const unsigned char r[4] = {(unsigned char)(w0), (unsigned char)(w0 >> 24), (unsigned char)(w1 >> 16), (unsigned char)(w2 >> 8)};
    const unsigned char g[4] = {(unsigned char)(w0 >> 8), (unsigned char)(w1), (unsigned char)(w1 >> 24), (unsigned char)(w2 >> 16)};
    const unsigned char bl[4] = {(unsigned char)(w0 >> 16), (unsigned char)(w1 >> 8), (unsigned char)(w2), (unsigned char)(w2 >> 24)};
    float4* d = reinterpret_cast<float4*>(out + b * 3 * plane4 * 4) + p;
    d[0] = make_float4(__fdiv_rn((float)r[0], 255.0f), __fdiv_rn((float)r[1], 255.0f), __fdiv_rn((float)r[2], 255.0f), __fdiv_rn((float)r[3], 255.0f));
    d[plane4] = make_float4(__fdiv_rn((float)g[0], 255.0f), __fdiv_rn((float)g[1], 255.0f), __fdiv_rn((float)g[2], 255.0f), __fdiv_rn((float)g[3], 255.0f));
    d[2 * plane4] = make_float4(__fdiv_rn((float)bl[0], 255.0f), __fdiv_rn((float)bl[1], 255.0f), __fdiv_rn((float)bl[2], 255.0f), __fdiv_rn((float)bl[3], 255.0f));
}

__device__ __forceinline__ unsigned pb_to_byte(float x) {
    return (unsigned)__float2int_rn(__fmul_rn(fminf(fmaxf(x, 0.0f), 1.0f), 255.0f));
}

__global__ void __launch_bounds__(256)
k_f32nchw_to_u8hwc_rgb4(const float* __restrict__ in, unsigned* __restrict__ out, size_t plane4, size_t total4) {
    const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= total4) return;
    const size_t b = i / plane4, p = i - b * plane4;
    const float4* s = reinterpret_cast<const float4*>(in + b * 3 * plane4 * 4) + p;
    const float4 r = __ldg(s), g = __ldg(s + plane4), bl = __ldg(s + 2 * plane4);
    out[3 * i] = pb_to_byte(r.x) | (pb_to_byte(g.x) << 8) | (pb_to_byte(bl.x) << 16) | (pb_to_byte(r.y) << 24);
    out[3 * i + 1] = pb_to_byte(g.y) | (pb_to_byte(bl.y) << 8) | (pb_to_byte(r.z) << 16) | (pb_to_byte(g.z) << 24);
    out[3 * i + 2] = pb_to_byte(bl.z) | (pb_to_byte(r.w) << 8) | (pb_to_byte(g.w) << 16) | (pb_to_byte(bl.w) << 24);
}

static bool rgb4_ok(const void* a, const void* b, int H, int W, int C) {
    return C == 3 && (((size_t)H * W) & 3) == 0 && (((uintptr_t)a | (uintptr_t)b) & 15) == 0;
}

int launch_u8_to_f32(const unsigned char* in, float* out, int B, int H, int W, int C, cudaStream_t stream) {
    const size_t plane = (size_t)H * W, total = (size_t)B * plane;
    ProfScope prof(PROF_OTHER, stream);
    if (rgb4_ok(in, out, H, W, C)) {
        k_u8hwc_to_f32nchw_rgb4<<<(unsigned)((total / 4 + 255) / 256), 256, 0, stream>>>(
            reinterpret_cast<const unsigned*>(in), out, plane / 4, total / 4);
        PB_LAUNCH_CHECK("k_u8hwc_to_f32nchw_rgb4");
        return PB_OK;
    }
    k_u8hwc_to_f32nchw<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(in, out, C, plane, total);
    PB_LAUNCH_CHECK("k_u8hwc_to_f32nchw");
    return PB_OK;
}

int launch_f32_to_u8(const float* in, unsigned char* out, int B, int C, int H, int W, cudaStream_t stream) {
    const size_t plane = (size_t)H * W, total = (size_t)B * plane;
    ProfScope prof(PROF_OTHER, stream);
    if (rgb4_ok(in, out, H, W, C)) {
        k_f32nchw_to_u8hwc_rgb4<<<(unsigned)((total / 4 + 255) / 256), 256, 0, stream>>>(
            in, reinterpret_cast<unsigned*>(out), plane / 4, total / 4);
        PB_LAUNCH_CHECK("k_f32nchw_to_u8hwc_rgb4");
        return PB_OK;
    }
    k_f32nchw_to_u8hwc<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(in, out, C, plane, total);
    PB_LAUNCH_CHECK("k_f32nchw_to_u8hwc");
    return PB_OK;
}

}  // namespace pb
