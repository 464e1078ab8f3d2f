// Host-callable launchers shared between the translation units of libpolyblur_sm100.so.
#pragma once
#include "common.cuh"
#include "fft2.cuh"

#define PB_FFT_THREADS 256
#define PB_STATS_STRIDE 16   // per image: [0] min (ordered), [1] max (ordered), [2..8] 7 maxima

namespace pb {

// ---- optional per-kernel timing with CUDA events (api.cu); used by bench.py ---------------
enum ProfClass { PROF_SETUP = 0, PROF_COLS, PROF_ROWS, PROF_PARAMS, PROF_DECONV_SPATIAL,
                 PROF_FFT_ROWS_FWD, PROF_FFT_COLS, PROF_FFT_ROWS_INV, PROF_OTHER, PROF_DECONV_NARROW,
                 PROF_NCLASSES };

// Deconvolution engine classes.  k_params classifies every image on the device from the extent
// of its significant taps and appends it to the class's list; each engine's kernel then takes
// its work grid-stride from that list (an empty class costs one tiny launch, no host sync).
#define PB_CLS_N11 0      // taps within 3 x 3: register-rolling kernel (deconv_narrow.cu)
#define PB_CLS_N22 1      // taps within 5 x 5: register-rolling kernel
#define PB_CLS_TILED 2    // shared-memory tiled Horner stencil (deconv.cu), radius > 4: 1 CTA / SM
#define PB_CLS_FFT 3      // blur-independent on-chip FFT engine (deconv_fft.cu)
#define PB_CLS_TILED4 4   // tiled stencil, radius <= 4: half the shared memory, 2 CTAs / SM
#define PB_NCLS 5
#define PB_CLS_COUNT_STRIDE 16   // ints reserved for the counters in front of the lists
struct ProfScope {
    int idx;
    cudaStream_t stream;
    ProfScope(int cls, cudaStream_t s);
    ~ProfScope();
};

// estimate.cu
void keys_weights_host(float* out210);
int upload_constants(cudaStream_t stream);
int launch_twiddles(float2* tw, int n, cudaStream_t stream);
int launch_init_stats(unsigned* stats, int B, cudaStream_t stream);
int fft_batch_for(int n, int max_nb);
int launch_cols(bool est, const float* img, float* gray, float* gy, unsigned* stats, int nimg, int C,
                int H, int W, const FftPlan& planH, const float2* twH, cudaStream_t stream);
int launch_rows(bool est, const float* plane_in, const float* gy, float* gx, unsigned* stats, int nimg,
                int H, int W, const FftPlan& planW, const float2* twW, int discard_saturation,
                cudaStream_t stream);
int launch_params(const unsigned* stats, ImgKernel* kern, float* est, const float* th, const float* sg,
                  const float* rh, const float* kin, float* kout, int mode, int B, int ksize, float cc,
                  float bb, float tap_thr, int engine_req, int fft_radius_min, int* cls, cudaStream_t stream);

// estimate2.cu (fast path: lengths with prime factors <= 13)
bool fft2_supported(int H, int W);
int launch_fft2_omega(float* omega, const Fft2Plan& plan, cudaStream_t stream);
int launch_fft2_stage_tw(float2* stw, const Fft2Plan& plan, cudaStream_t stream);
int launch_rows2(bool est, const float* img, float* gray, float* gx, unsigned* stats, int nimg, int C,
                 int H, int W, const Fft2Plan& planW, const float2* twW, const float* omegaW,
                 const float* qrange, cudaStream_t stream);
int launch_cols2(bool est, const float* plane_in, const float* gx, float* gy, unsigned* stats, int nimg,
                 int H, int W, const Fft2Plan& planH, const float2* twH, const float* omegaH,
                 int discard_saturation, const float* mask_src, cudaStream_t stream);

// quantile.cu (q > 0 range normalisation)
size_t quantile_workspace_bytes(int B);
int launch_quantile_range(const float* img, float* gray, int B, int C, int H, int W, double q, void* ws,
                          float** qrange_out, cudaStream_t stream);

// stages.cu (optional stages: prefilters, halo masking, edgetaper)
int launch_bilateral(const float* img, float* out, int planes, int H, int W, float sigma_spatial,
                     float sigma_color, cudaStream_t stream);
size_t rf_workspace_bytes(int B, int H, int W);
int launch_recursive_filter(const float* in, const float* joint, float* out, int B, int C, int H, int W,
                            double sigma_s, double sigma_r, int num_iterations, void* ws, cudaStream_t stream);
int launch_residual_add(float* dec, const float* cur, const float* smooth, size_t n, cudaStream_t stream);
int launch_halo_norm(const float* gx, const float* gy, float* partial, float* nM, int planes, size_t plane,
                     cudaStream_t stream);
int launch_halo_apply(float* imout, const float* img, size_t img_plane, int img_pitch, int img_off, const float* gx,
                      const float* gy, const float* ox, const float* nM, int planes, int H, int W,
                      cudaStream_t stream);
size_t edgetaper_scratch_bytes(int B, int Hp, int Wp);
int launch_edgetaper_weights(const ImgKernel* kern, void* scratch, int B, int Hp, int Wp, int batch_max,
                             float** v_out, cudaStream_t stream);
// backward pass through the estimator (backward.cu)
#define PB_BW_TRACE_STRIDE 24   // per image: 7 maxima, 7 signs, min, max, #min, #max
int launch_bw_trace(const float* img, float* g, float* gn, unsigned* stats, unsigned long long* keys, int B, int C,
                    int H, int W, cudaStream_t stream);
int launch_bw_dirmax(const float* gx, const float* gy, const float* g, unsigned long long* keys, unsigned* stats,
                     float* trace_f, int* trace_pos, int B, int H, int W, cudaStream_t stream);
int launch_bw_kernel_grad(const float* gout, const float* preclamp, const float* V, float* kbar, int B, int C, int H,
                          int W, int ks, cudaStream_t stream);
int launch_bw_scatter(const float* mbar, const float* trace_f, const int* trace_pos, float* sgx, float* sgy, int B,
                      int H, int W, cudaStream_t stream);
int launch_bw_norm(const float* dx, const float* dy, const float* img, const float* trace_f, double* sums, float* gin,
                   int B, int C, int H, int W, cudaStream_t stream);
int launch_vjp_embed(const float* gout, const float* preclamp, float* z, int planes, int H, int W, int pad,
                     cudaStream_t stream);
int launch_vjp_fold(const float* t, float* gin, int planes, int H, int W, int pad, cudaStream_t stream);
int launch_flip_kernels(const float* k, float* kf, int B, int ksize, cudaStream_t stream);
int launch_pad_replicate(const float* img, float* a, float* b, int planes, int H, int W, int pad,
                         cudaStream_t stream);
int launch_edgetaper_passes(float* a, float* b, const ImgKernel* kern, const float* v, int B, int C, int Hp, int Wp,
                            int n_tapers, float** result, cudaStream_t stream);

// nc.cu (domain-transform normalized convolution)
size_t nc_workspace_bytes(int B, int C, int H, int W);
int launch_normalized_convolution(const float* img, float* out, int B, int C, int H, int W, double sigma_s,
                                  double sigma_r, int num_iterations, void* ws, cudaStream_t stream);

// io.cu (8-bit HWC <-> float32 NCHW)
int launch_u8_to_f32(const unsigned char* in, float* out, int B, int H, int W, int C, cudaStream_t stream);
int launch_f32_to_u8(const float* in, unsigned char* out, int B, int C, int H, int W, cudaStream_t stream);

// deconv_narrow.cu
int launch_deconv_narrow(int cls, const float* img, float* out, const ImgKernel* kern, const int* list,
                         const int* count, int B, int C, int H, int W, float a3, float a2, float a1, float b0,
                         const SrcGeom& G, cudaStream_t stream);

// deconv_fft.cu (blur-independent on-chip FFT engine)
struct FftEngineLayout {
    int NX, NY;
    size_t off_twX, off_twY, off_stwX, off_stwY, off_slotX, off_slotY, off_freqY, off_Z, total;
};
struct FftEngineTables {
    int NX, NY;
    Fft2Plan planX, planY;
    float2 *twX, *twY;       // master tables exp(-2 pi i k / N) (kernel spectrum)
    float2 *stwX, *stwY;     // stage-twiddle tables of the two plans
    int *slotX, *slotY, *freqY;
    float2* Z;
};
int fft_engine_length(int n);
bool fft_engine_supported(int H, int W, int pad);
size_t fft_engine_workspace(int B, int C, int H, int W, int pad, FftEngineLayout* L);
int fft_engine_prepare(char* base, const FftEngineLayout& L, FftEngineTables* T, cudaStream_t stream);
int launch_deconv_fft(const float* img, float* out, const ImgKernel* kern, const int* list, const int* count,
                      int B, int C, int H, int W, const FftEngineTables& T, float a3, float a2, float a1,
                      float b0, const SrcGeom& G, cudaStream_t stream);

// deconv.cu
int launch_deconv_spatial(const float* img, float* out, const ImgKernel* kern, const int* list, const int* count,
                          int B, int C, int H, int W, float a3, float a2, float a1, float b0, const SrcGeom& G,
                          int max_radius, cudaStream_t stream);

}  // namespace pb
