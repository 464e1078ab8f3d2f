// Patch decomposition of PolyblurDeblurring.forward on the device (polyblur/deblurring.py:269-340): the image is
// centre-padded (replicate) to a whole number of steps, cut into overlapping patch_size x patch_size patches that
// step by int(patch_size (1 - overlap)), every patch is deblurred with its own blur estimate, and the results are
// blended with a separable Kaiser window:  restored = sum_p out_p w / (sum_p w + 1e-8), clamped, cropped back.
//
//   k_patch_extract : padded image -> patch batch, the replicate border by index clamping (no padded copy, no cat)
//   k_patch_blend   : gather form of the overlap-add: every output pixel sums the (at most 2 x 2 for overlaps
//                     <= 50 %) patches that cover it in the reference's patch order, normalises, clamps, crops
//
// Patch p = iy * nx + ix covers padded rows [iy step_h, iy step_h + ph); batch index of (patch p, image b) is
// p * B + b, which is the order torch.cat of the reference's loop produces.
#include "kernels.cuh"

namespace pb {

__global__ void __launch_bounds__(256)
k_patch_extract(const float* __restrict__ img, size_t plane_stride, size_t row_stride, float* __restrict__ patches,
                int B, int C, int h, int w, int ph, int pw, int step_h, int step_w, int nx, int pad_top, int pad_left) {
    const int x = blockIdx.x * 64 + (threadIdx.x & 63);
    const int y = blockIdx.y * 4 + (threadIdx.x >> 6);
    if (x >= pw || y >= ph) return;
    const int z = blockIdx.z;                    // (p * B + b) * C + c
    const int c = z % C, pb_ = z / C, b = pb_ % B, p = pb_ / B;
    const int iy = p / nx, ix = p - iy * nx;
    const int sy = min(max(iy * step_h + y - pad_top, 0), h - 1);
    const int sx = min(max(ix * step_w + x - pad_left, 0), w - 1);
    patches[((size_t)z * ph + y) * pw + x] = __ldg(img + ((size_t)b * C + c) * plane_stride + (size_t)sy * row_stride + sx);
}

__global__ void __launch_bounds__(256)
k_patch_blend(const float* __restrict__ patches, const float* __restrict__ win_y, const float* __restrict__ win_x,
              float* __restrict__ out, int B, int C, int h, int w, int ph, int pw, int step_h, int step_w, int ny,
              int nx, int pad_top, int pad_left) {
    const int X = blockIdx.x * 64 + (threadIdx.x & 63);
    const int Y = blockIdx.y * 4 + (threadIdx.x >> 6);
    if (X >= w || Y >= h) return;
    const int z = blockIdx.z;                    // b * C + c
    const int c = z % C, b = z / C;
    const int yy = Y + pad_top, xx = X + pad_left;
    // patches whose rows / columns cover this pixel: iy step <= yy < iy step + ph
    const int iy0 = max(0, (yy - ph + step_h) / step_h), iy1 = min(ny - 1, yy / step_h);
    const int ix0 = max(0, (xx - pw + step_w) / step_w), ix1 = min(nx - 1, xx / step_w);
    float acc = 0.0f, wsum = 0.0f;
    for (int iy = iy0; iy <= iy1; ++iy) {
        const int y = yy - iy * step_h;
        if (y < 0 || y >= ph) continue;
        const float wy = __ldg(win_y + y);
        for (int ix = ix0; ix <= ix1; ++ix) {
            const int x = xx - ix * step_w;
            if (x < 0 || x >= pw) continue;
            const float wgt = __fmul_rn(wy, __ldg(win_x + x));           // window[y][x] as the reference builds it
            const size_t pz = ((size_t)(iy * nx + ix) * B + b) * C + c;
            const float v = __ldg(patches + (pz * ph + y) * pw + x);
            acc = __fadd_rn(acc, __fmul_rn(v, wgt));                     // restored += out * window, in patch order
            wsum = __fadd_rn(wsum, wgt);
        }
    }
    const float r = __fdiv_rn(acc, __fadd_rn(wsum, 1e-8f));
    out[((size_t)z * h + Y) * w + X] = fminf(fmaxf(r, 0.0f), 1.0f);
}

int launch_patch_extract(const float* img, size_t plane_stride, size_t row_stride, float* patches, int B, int C, int h,
                         int w, int ph, int pw, int step_h, int step_w, int ny, int nx, int pad_top, int pad_left,
                         cudaStream_t stream) {
    const long long nz = (long long)ny * nx * B * C;
    if (nz > 65535) {
        set_error("patch planes (%lld) exceed the grid z limit", nz);
        return PB_ERR_ARG;
    }
    dim3 grid((pw + 63) / 64, (ph + 3) / 4, (unsigned)nz);
    ProfScope prof(PROF_OTHER, stream);
    k_patch_extract<<<grid, 256, 0, stream>>>(img, plane_stride, row_stride, patches, B, C, h, w, ph, pw, step_h, step_w, nx,
                                              pad_top, pad_left);
    PB_LAUNCH_CHECK("k_patch_extract");
    return PB_OK;
}

int launch_patch_blend(const float* patches, const float* win_y, const float* win_x, float* out, int B, int C, int h,
                       int w, int ph, int pw, int step_h, int step_w, int ny, int nx, int pad_top, int pad_left,
                       cudaStream_t stream) {
    if ((long long)B * C > 65535) {
        set_error("B*C = %d exceeds the grid z limit", B * C);
        return PB_ERR_ARG;
    }
    dim3 grid((w + 63) / 64, (h + 3) / 4, B * C);
    ProfScope prof(PROF_OTHER, stream);
    k_patch_blend<<<grid, 256, 0, stream>>>(patches, win_y, win_x, out, B, C, h, w, ph, pw, step_h, step_w, ny, nx, pad_top,
                                            pad_left);
    PB_LAUNCH_CHECK("k_patch_blend");
    return PB_OK;
}

}  // namespace pb
