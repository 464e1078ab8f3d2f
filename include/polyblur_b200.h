/* polyblur_b200 -- C ABI of the B200-native Polyblur engine (libpolyblur_sm100.so).
 *
 * This is the drop-in boundary for the hot path of teboli/polyblur: every entry point
 * below replaces a Python function of the reference (file:line cited per function,
 * paths relative to the reference checkout).  The reference has no FFI of its own -- it
 * is a pure-Python package over ATen -- so the "binding a maintainer would add" is a
 * ctypes stub; INTEGRATION.md shows it.
 *
 * Conventions
 *   - all image pointers are DEVICE pointers to float32, NCHW, contiguous;
 *   - the library allocates nothing: scratch lives in a caller-owned device workspace
 *     of pb_workspace_bytes() bytes (256-byte aligned);
 *   - work is enqueued on the cudaStream_t passed as `stream` (void* here so that the
 *     header needs no CUDA include); no entry point synchronises the host;
 *   - return value 0 = success, negative = error (pb_last_error() gives the text).
 *   - there is NO CPU fallback: without a CUDA device every compute entry point
 *     returns PB_ERR_CUDA.
 */
#ifndef POLYBLUR_B200_H
#define POLYBLUR_B200_H

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define PB_API __attribute__((visibility("default")))
#else
#define PB_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define PB_VERSION 100            /* 0.1.0 */
#define PB_KSIZE_MAX 25           /* kernel support of the reference (deblurring.py:24) */
#define PB_EST_STRIDE 12          /* floats per (iteration, image) in est_out, see below */

/* error codes */
#define PB_OK 0
#define PB_ERR_ARG (-1)           /* bad argument (shape, null pointer, unsupported option) */
#define PB_ERR_WORKSPACE (-2)     /* workspace too small / misaligned */
#define PB_ERR_CUDA (-3)          /* CUDA runtime error (no device, launch failure) */
#define PB_ERR_UNSUPPORTED (-4)   /* valid in the reference but not built yet */

/* option flags (keyword arguments of polyblur_deblurring, deblurring.py:23-25) */
#define PB_FLAG_REMOVE_HALO        0x01u
#define PB_FLAG_EDGETAPER          0x02u
#define PB_FLAG_PREFILTER          0x04u  /* bilateral, what prefiltering=True runs (deblurring.py:108) */
#define PB_FLAG_PREFILTER_RF       0x08u  /* domain-transform RF instead (the commented call, :107) */
#define PB_FLAG_DISCARD_SATURATION 0x10u
#define PB_FLAG_EDGETAPER_BATCHMAX 0x20u  /* bug-compatible batch-global max (edgetaper.py:15,21) */
#define PB_FLAG_NO_CLAMP           0x40u  /* pb_deconv_ex_f32 only: skip the final clamp (for the backward pass) */

/* deconvolution engine selection (pb_params.engine); AUTO decides per image on device */
#define PB_ENGINE_AUTO    0
#define PB_ENGINE_SPATIAL 1       /* sparse-tap fused Horner stencil in shared memory      */
#define PB_ENGINE_FFT     2       /* on-chip FFT of the padded torus (blur independent)    */

/* Scalars are doubles because the reference receives them as Python floats and rounds
 * derived quantities (c*c, the polynomial coefficients) to float32 only when they meet a
 * tensor (blur_estimation.py:176-177, deblurring.py:160-166). */
typedef struct pb_params {
    double   c;           /* affine blur model slope      (blur_estimation.py:171-185)  */
    double   b;           /* affine blur model intercept                                */
    double   alpha;       /* polynomial parameters        (deblurring.py:160-162)       */
    double   beta;
    double   sigma_s;     /* prefilter parameters (RF only; the bilateral ignores them) */
    double   sigma_r;
    double   q;           /* quantile of the normalisation (0 = min/max)                */
    int32_t  n_iter;      /* deblurring.py:23 n_iter                                    */
    int32_t  ker_size;    /* odd, <= PB_KSIZE_MAX                                       */
    uint32_t flags;       /* PB_FLAG_*                                                  */
    int32_t  engine;      /* PB_ENGINE_*                                                */
    float    tap_rel_threshold; /* spatial engine: drop taps < thr * max tap (0 = default 1e-8) */
    int32_t  chunk_images;      /* pb_polyblur_f32: images per engine pass (0 = the whole batch).  The batch runs
                                 * in groups of this many images, each through the whole loop, and the workspace
                                 * (pb_workspace_bytes) is sized for one group: bounds device memory for huge batches.
                                 * Results do not depend on it; not combinable with PB_FLAG_EDGETAPER_BATCHMAX. */
} pb_params;

/* ---- library / host-only entry points (usable without a GPU) ------------------------- */

PB_API int pb_version(void);
PB_API const char* pb_last_error(void);

/* Fills p with the defaults of polyblur_deblurring (deblurring.py:23-25). */
PB_API void pb_default_params(pb_params* p);

/* Polynomial coefficients a3,a2,a1,b of deblurring.py:160-162 -> out[4]. */
PB_API void pb_polynomial_coefficients(double alpha, double beta, float* out4);

/* (30,7) normalised Keys cubic weights of blur_estimation.py:138-148,157-158 -> out[210]. */
PB_API void pb_keys_weights(float* out210);

/* Radix sequence the on-chip FFT uses for a length-n transform; returns the number of
 * stages written to radices (<= 32), or a negative error. */
PB_API int pb_fft_plan(int n, int* radices);

/* Bytes of device workspace needed by any entry point below for this shape. */
PB_API size_t pb_workspace_bytes(int B, int C, int H, int W, const pb_params* p);

/* Per-kernel timing with CUDA events recorded around every launch the library makes
 * (bench.py's roofline leg).  begin() resets and enables; end() disables, waits for the
 * recorded events and returns the number of kernel classes, filling total milliseconds and
 * launch counts per class (arrays of at least max_classes entries). */
PB_API int pb_profile_begin(void);
PB_API int pb_profile_end(float* ms_per_class, int* launches_per_class, int max_classes);
PB_API const char* pb_profile_class_name(int cls);

/* ---- the hot path ---------------------------------------------------------------------- */

/* polyblur_deblurring (deblurring.py:23-96), method='fft' semantics, tensor path.
 *   in, out : (B,C,H,W) float32 device; out may not alias in.
 *   est_out : NULL or device float[n_iter][B][PB_EST_STRIDE] =
 *             {m_0..m_6, theta_deg, sigma, rho, m_normal, m_ortho} per image-iteration
 *             (the taps gaussian_blur_estimation computes, blur_estimation.py:18-79). */
PB_API int pb_polyblur_f32(const float* in, float* out, int B, int C, int H, int W,
                    const pb_params* p, void* workspace, size_t workspace_bytes,
                    float* est_out, void* stream);

/* ---- stage-level entry points (tests, stage-level drop-ins) --------------------------- */

/* filters.fourier_gradients (filters.py:159-186): spectral derivative of every channel. */
PB_API int pb_fourier_gradients_f32(const float* img, float* gx, float* gy, int B, int C, int H, int W,
                             void* workspace, size_t workspace_bytes, void* stream);

/* blur_estimation.gaussian_blur_estimation up to the parameters (blur_estimation.py:18-73):
 * est = device float[B][PB_EST_STRIDE] as above. */
PB_API int pb_estimate_f32(const float* img, int B, int C, int H, int W, double c, double b, double q,
                    uint32_t flags, float* est, void* workspace, size_t workspace_bytes,
                    void* stream);

/* blur_estimation.create_gaussian_filter (blur_estimation.py:211-232):
 * theta (radians), sigma, rho: device float[B] -> kernel device float[B][ksize][ksize].
 * Needs B * 4096 bytes of workspace. */
PB_API int pb_make_kernel_f32(const float* theta, const float* sigma, const float* rho, int B,
                       int ksize, float* kernel, void* workspace, size_t workspace_bytes,
                       void* stream);

/* deblurring.inverse_filtering_rank3 with default flags (deblurring.py:211-239):
 * replicate pad, polynomial on the torus, crop, clamp to [0,1];
 * kernel = device float[B][ksize][ksize] (one per image, broadcast over channels). */
PB_API int pb_deconv_f32(const float* img, float* out, int B, int C, int H, int W,
                  const float* kernel, int ksize, double alpha, double beta, int engine,
                  void* workspace, size_t workspace_bytes, void* stream);

/* deblurring.inverse_filtering_rank3 with its optional stages (deblurring.py:211-239):
 * flags = PB_FLAG_EDGETAPER (do_edgetaper, + PB_FLAG_EDGETAPER_BATCHMAX) | PB_FLAG_REMOVE_HALO |
 * PB_FLAG_NO_CLAMP;
 * grad_x / grad_y = grad_img of the reference (device, (B,C,H,W)) or both NULL = gradients of img.
 * Workspace: pb_workspace_bytes with the same flags / ker_size / engine in pb_params. */
PB_API int pb_deconv_ex_f32(const float* img, float* out, int B, int C, int H, int W, const float* kernel,
                     int ksize, double alpha, double beta, int engine, uint32_t flags, const float* grad_x,
                     const float* grad_y, void* workspace, size_t workspace_bytes, void* stream);

/* Vector-Jacobian product of deblurring.inverse_filtering_rank3 (default flags, deblurring.py:211-239)
 * with respect to the image, the kernel held constant -- what torch.autograd computes for the
 * reference's replicate pad -> circular polynomial filter -> crop -> clamp chain (README.md:70 claims
 * the module differentiable): grad_img = pad^T filter^T crop^T (grad_out * pass), pass = 1 where
 * `preclamp` (the forward result computed with PB_FLAG_NO_CLAMP, or NULL = no clamp) lies in [0,1].
 * Workspace: pb_deconv_vjp_workspace_bytes. */
PB_API size_t pb_deconv_vjp_workspace_bytes(int B, int C, int H, int W, int ksize, int engine);
PB_API int pb_deconv_vjp_f32(const float* grad_out, const float* preclamp, float* grad_img, int B, int C,
                      int H, int W, const float* kernel, int ksize, double alpha, double beta, int engine,
                      void* workspace, size_t workspace_bytes, void* stream);

/* The rest of the backward pass: what torch.autograd computes through the blur estimator
 * (blur_estimation.py:18-79) and the kernel argument of inverse_filtering_rank3.  One workspace size
 * serves the three calls.
 *   pb_estimate_trace_f32: forward trace of the estimator on img: trace_f = device float[B][24]
 *     (7 directional maxima of the normalised gray image, the sign of cos gx - sin gy at their arg-max
 *     pixels, min, max of the gray image and how many pixels attain them), trace_pos = device
 *     int32[B][8] (arg-max pixel y*W+x per angle).
 *   pb_kernel_grad_f32: gradient of <grad_out, inverse_filtering_rank3(img, kernel)> with respect to
 *     the kernel taps -> kernel_grad device float[B][ksize][ksize] (preclamp as in pb_deconv_vjp_f32).
 *   pb_estimator_vjp_f32: given mbar = device float[B][7], the gradient with respect to the 7 maxima
 *     (the scalar chain kernel taps -> sigma, rho -> maxima is left to the caller), ADDS the gradient
 *     with respect to img into grad_img: arg-max scatter, transposed spectral derivative, range
 *     normalisation (blur_estimation.py:96-134), channel mean. */
PB_API size_t pb_backward_workspace_bytes(int B, int C, int H, int W, int ksize, int engine);
PB_API int pb_estimate_trace_f32(const float* img, int B, int C, int H, int W, float* trace_f, int* trace_pos,
                          void* workspace, size_t workspace_bytes, void* stream);
/* The same with flags = 0 or PB_FLAG_DISCARD_SATURATION: the arg-max search then leaves out the pixels whose gray value
 * is > 0.99 (get_saturation_mask + compute_gradients, blur_estimation.py:83-88, 112-119). */
PB_API int pb_estimate_trace_ex_f32(const float* img, int B, int C, int H, int W, uint32_t flags, float* trace_f,
                             int* trace_pos, void* workspace, size_t workspace_bytes, void* stream);
PB_API int pb_kernel_grad_f32(const float* img, const float* grad_out, const float* preclamp, int B, int C,
                       int H, int W, const float* kernel, int ksize, double alpha, double beta, int engine,
                       float* kernel_grad, void* workspace, size_t workspace_bytes, void* stream);
PB_API int pb_estimator_vjp_f32(const float* img, const float* mbar, const float* trace_f, const int* trace_pos,
                         float* grad_img, int B, int C, int H, int W, void* workspace, size_t workspace_bytes,
                         void* stream);

/* edgetaper.edgetaper (edgetaper.py:26-33) on an already padded image. */
PB_API int pb_edgetaper_f32(const float* img, float* out, int B, int C, int H, int W,
                     const float* kernel, int ksize, int n_tapers, uint32_t flags,
                     void* workspace, size_t workspace_bytes, void* stream);

/* filters.bilateral_filter (filters.py:107-148), 5x5, sigma_spatial=5, sigma_color=0.1. */
PB_API int pb_bilateral_f32(const float* img, float* out, int B, int C, int H, int W,
                     float sigma_spatial, float sigma_color, void* stream);
/* Vector-Jacobian product of pb_bilateral_f32 (what torch.autograd computes over filters.py:107-148): grad_img =
 * (d out / d img)^T grad_out, every window tap's range weight differentiated; grad_img is overwritten. */
PB_API int pb_bilateral_vjp_f32(const float* img, const float* grad_out, float* grad_img, int B, int C, int H, int W,
                         float sigma_spatial, float sigma_color, void* stream);

/* domain_transform.recursive_filter (domain_transform.py:6-85); joint may be NULL. */
PB_API int pb_recursive_filter_f32(const float* img, const float* joint, float* out, int B, int C, int H,
                            int W, float sigma_s, float sigma_r, int num_iterations,
                            void* workspace, size_t workspace_bytes, void* stream);

/* 8-bit I/O around the hot path, what main.py does on the host (main.py:80 img_as_float32, :146
 * img_as_ubyte), fused with the layout change: (B,H,W,C) uint8 <-> (B,C,H,W) float32 in [0,1]. */
PB_API int pb_u8hwc_to_f32nchw(const uint8_t* in, float* out, int B, int H, int W, int C, void* stream);
PB_API int pb_f32nchw_to_u8hwc(const float* in, uint8_t* out, int B, int C, int H, int W, void* stream);

/* Patch decomposition of PolyblurDeblurring.forward (deblurring.py:269-340) on the device.
 * Geometry (all in pixels): the (h, w) image -- already cropped to even sides, :273-279; plane_stride / row_stride
 * in floats address it inside a larger allocation -- is centre-padded (replicate, :282-287, :368-377) by pad_top /
 * pad_left to the patch grid of ny x nx patches of ph x pw stepping by step_h / step_w.
 *   pb_patch_extract_f32: patches[(p * B + b)][c][y][x], p = iy * nx + ix  (the order of the reference's torch.cat)
 *   pb_patch_blend_f32:   out (B,C,h,w) = clamp(sum_p patches_p w / (sum_p w + 1e-8), 0, 1), w[y][x] = win_y[y] win_x[x]
 *                         (:312-339 overlap-add with build_window :349-366, summed in patch order, then cropped). */
PB_API int pb_patch_extract_f32(const float* img, size_t plane_stride, size_t row_stride, float* patches, int B, int C,
                         int h, int w, int ph, int pw, int step_h, int step_w, int ny, int nx, int pad_top,
                         int pad_left, void* stream);
PB_API int pb_patch_blend_f32(const float* patches, const float* win_y, const float* win_x, float* out, int B, int C,
                       int h, int w, int ph, int pw, int step_h, int step_w, int ny, int nx, int pad_top,
                       int pad_left, void* stream);

/* Domain-transform normalized convolution: the reference's native prototype
 * normalized_convolution(I, sigma_s, sigma_r, num_iterations) (polyblur/domain_transform/NC.cpp:143-204,
 * exported at :210), here for any batch size and channel count.
 * Needs 2 * B*H*W*4 + 2 * B*C*H*W*4 bytes of workspace (+ 256-byte alignment slack per block). */
PB_API int pb_normalized_convolution_f32(const float* img, float* out, int B, int C, int H, int W, float sigma_s,
                                  float sigma_r, int num_iterations, void* workspace,
                                  size_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* POLYBLUR_B200_H */
