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

int launch_u8_to_f32(const unsigned char* in, float* out, int B, int H, int W, int C, cudaStream_t stream) {
    const size_t plane = (size_t)H * W, total = (size_t)B * plane;
    ProfScope prof(PROF_OTHER, stream);
    k_u8hwc_to_f32nchw<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(in, out, C, plane, total);
    PB_LAUNCH_CHECK("k_u8hwc_to_f32nchw");
    return PB_OK;
}

int launch_f32_to_u8(const float* in, unsigned char* out, int B, int C, int H, int W, cudaStream_t stream) {
    const size_t plane = (size_t)H * W, total = (size_t)B * plane;
    ProfScope prof(PROF_OTHER, stream);
    k_f32nchw_to_u8hwc<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(in, out, C, plane, total);
    PB_LAUNCH_CHECK("k_f32nchw_to_u8hwc");
    return PB_OK;
}

}  // namespace pb
