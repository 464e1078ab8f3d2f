// extern "C" entry points of libpolyblur_sm100.so (see include/polyblur_b200.h).
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <vector>

#include "kernels.cuh"

namespace pb {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int check_cuda(cudaError_t e, const char* what) {
    if (e == cudaSuccess) return PB_OK;
    set_error("CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
    return PB_ERR_CUDA;
}

// ---- profiler ------------------------------------------------------------------------------
struct ProfRec { int cls; cudaEvent_t a, b; };
// process-global, guarded by g_prof_mu (host threads may drive the library concurrently on their own streams)
static std::mutex g_prof_mu;
static bool g_prof_on = false;
static std::vector<ProfRec> g_prof;
static std::vector<cudaEvent_t> g_event_pool;
static const char* kProfNames[PROF_NCLASSES] = {"setup", "k_cols", "k_rows", "k_params", "k_deconv_spatial",
                                                "k_fft_rows_fwd", "k_fft_cols", "k_fft_rows_inv", "other",
                                                "k_deconv_narrow"};

static cudaEvent_t get_event() {
    if (!g_event_pool.empty()) {
        cudaEvent_t e = g_event_pool.back();
        g_event_pool.pop_back();
        return e;
    }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
}

ProfScope::ProfScope(int cls, cudaStream_t s) : idx(-1), stream(s) {
    std::lock_guard<std::mutex> lock(g_prof_mu);
    if (!g_prof_on) return;
    ProfRec r{cls, get_event(), get_event()};
    cudaEventRecord(r.a, s);
    idx = (int)g_prof.size();
    g_prof.push_back(r);
}
ProfScope::~ProfScope() {
    if (idx < 0) return;
    std::lock_guard<std::mutex> lock(g_prof_mu);
    if (idx < (int)g_prof.size()) cudaEventRecord(g_prof[idx].b, stream);
}

int make_fft_plan(int n, FftPlan* plan) {
    if (n < 1) return PB_ERR_ARG;
    plan->n = n;
    plan->ns = 0;
    int m = n;
    static const int pref[] = {8, 7, 5, 4, 3, 2};
    for (int r : pref) {
        while (m % r == 0) {
            if (plan->ns >= PB_MAX_STAGES) return PB_ERR_ARG;
            plan->radix[plan->ns++] = r;
            m /= r;
        }
    }
    for (int p = 11; m > 1; p += 2) {
        while (m % p == 0) {
            if (plan->ns >= PB_MAX_STAGES) return PB_ERR_ARG;
            plan->radix[plan->ns++] = p;
            m /= p;
        }
        if ((long long)p * p > m && m > 1) {   // m is prime
            if (plan->ns >= PB_MAX_STAGES) return PB_ERR_ARG;
            plan->radix[plan->ns++] = m;
            m = 1;
        }
    }
    return PB_OK;
}

// ---- workspace layout -------------------------------------------------------------------------
struct Workspace {
    size_t off_stats, off_kern, off_cls, off_twH, off_twW, off_omH, off_omW, off_gray, off_gy, off_tmp, off_fft, total;
    // optional stages (allocated only when their flag is set)
    size_t off_smooth, off_rf, off_padA, off_padB, off_et, off_g0x, off_g0y, off_ox, off_nm, off_ghat, off_q;
    bool has_fft;
    FftEngineLayout fft;
};

// Radius (max |tap offset|) from which AUTO hands an image to the FFT engine.  Measured on B200,
// 32 x 3 x 1080 x 1920, one deconvolution: radius 1 narrow 0.56 ms, radius 2 narrow 1.21 ms, radius 3
// tiled 4.6-5.6 ms, radius 4 tiled 6.0 ms, radius 5 tiled 22 ms; FFT engine 2.79 ms for any radius.
#define PB_FFT_RADIUS_MIN 3

static Workspace layout(int B, int C, int H, int W, int n_iter, int ksize = PB_KS, int engine = PB_ENGINE_AUTO,
                        uint32_t flags = 0, double q = 0.0) {
    Workspace w;
    size_t o = 0;
    auto take = [&](size_t bytes) {
        size_t at = o;
        o = align_up(o + bytes, 256);
        return at;
    };
    const size_t plane = (size_t)H * W;
    const size_t nimg = (size_t)B * (C > 0 ? C : 1);   // stage entry points treat channels as images
    w.off_stats = take(nimg * PB_STATS_STRIDE * sizeof(unsigned));
    w.off_kern = take(nimg * sizeof(ImgKernel));
    w.off_cls = take((PB_CLS_COUNT_STRIDE + PB_NCLS * nimg) * sizeof(int));
    w.off_twH = take((size_t)H * sizeof(float2));
    w.off_twW = take((size_t)W * sizeof(float2));
    w.off_omH = take((size_t)H * sizeof(float));
    w.off_omW = take((size_t)W * sizeof(float));
    w.off_gray = take((size_t)B * plane * sizeof(float));
    w.off_gy = take((size_t)B * plane * sizeof(float));
    w.off_tmp = take(n_iter >= 2 ? (size_t)B * C * plane * sizeof(float) : 0);
    w.has_fft = engine != PB_ENGINE_SPATIAL && fft_engine_supported(H, W, ksize / 2);
    w.off_fft = o;
    if (w.has_fft) take(fft_engine_workspace(B, C, H, W, ksize / 2, &w.fft));
    const size_t img_bytes = (size_t)B * C * plane * sizeof(float);
    const int pad = ksize / 2;
    const size_t padded_bytes = (size_t)B * C * (H + 2 * pad) * (W + 2 * pad) * sizeof(float);
    w.off_smooth = take((flags & (PB_FLAG_PREFILTER | PB_FLAG_PREFILTER_RF)) ? img_bytes : 0);
    w.off_rf = take((flags & PB_FLAG_PREFILTER_RF) ? rf_workspace_bytes(B, H, W) : 0);
    w.off_padA = take((flags & PB_FLAG_EDGETAPER) ? padded_bytes : 0);
    w.off_padB = take((flags & PB_FLAG_EDGETAPER) ? padded_bytes : 0);
    w.off_et = take((flags & PB_FLAG_EDGETAPER) ? edgetaper_scratch_bytes(B, H + 2 * pad, W + 2 * pad) : 0);
    w.off_g0x = take((flags & PB_FLAG_REMOVE_HALO) ? img_bytes : 0);
    w.off_g0y = take((flags & PB_FLAG_REMOVE_HALO) ? img_bytes : 0);
    w.off_ox = take((flags & PB_FLAG_REMOVE_HALO) ? img_bytes : 0);
    w.off_nm = take((flags & PB_FLAG_REMOVE_HALO) ? (size_t)B * C * (64 + 1) * sizeof(float) : 0);
    w.off_ghat = take(q > 0 ? (size_t)B * plane * sizeof(float) : 0);
    w.off_q = take(q > 0 ? quantile_workspace_bytes(B) : 0);
    w.total = o;
    return w;
}

static int check_shape(int B, int C, int H, int W) {
    if (B < 1 || C < 1 || H < 1 || W < 1) {
        set_error("bad shape B=%d C=%d H=%d W=%d", B, C, H, W);
        return PB_ERR_ARG;
    }
    if ((size_t)H * W > (size_t)1 << 31) {
        set_error("plane of %d x %d exceeds 2^31 pixels", H, W);
        return PB_ERR_ARG;
    }
    return PB_OK;
}

static int check_ws(const void* ws, size_t have, size_t need) {
    if (!ws || have < need) {
        set_error("workspace too small: have %zu bytes, need %zu", have, need);
        return PB_ERR_WORKSPACE;
    }
    if (((uintptr_t)ws & 255u) != 0) {
        set_error("workspace must be 256-byte aligned");
        return PB_ERR_WORKSPACE;
    }
    return PB_OK;
}

struct Tables {
    FftPlan planH, planW;
    float2 *twH, *twW;
    bool fast;                 // both lengths run on the fft2 core (estimate2.cu)
    Fft2Plan planH2, planW2;
    float *omH, *omW;
};

// Fills the table pointers and queues the generation of the tables (launch_table_jobs runs the queue);
// the any-length fallback core keeps its own master-twiddle launches.
static int prepare_tables(char* ws, const Workspace& L, int H, int W, Tables* t, TableJobs* jobs, cudaStream_t stream) {
    if (make_fft_plan(H, &t->planH) || make_fft_plan(W, &t->planW)) {
        set_error("cannot plan FFT for %d x %d", H, W);
        return PB_ERR_ARG;
    }
    t->twH = reinterpret_cast<float2*>(ws + L.off_twH);
    t->twW = reinterpret_cast<float2*>(ws + L.off_twW);
    t->fast = fft2_supported(H, W);
    t->omH = reinterpret_cast<float*>(ws + L.off_omH);
    t->omW = reinterpret_cast<float*>(ws + L.off_omW);
    if (t->fast) {
        // the fft2 core reads per-stage twiddle tables (< n entries each): they take the place of
        // the master tables of the any-length core
        make_fft2_plan(H, &t->planH2);
        make_fft2_plan(W, &t->planW2);
        jobs->add(TJ_STAGE_TW, t->planH2.tw_total, t->twH, nullptr, &t->planH2);
        jobs->add(TJ_STAGE_TW, t->planW2.tw_total, t->twW, nullptr, &t->planW2);
        jobs->add(TJ_OMEGA, H, t->omH, nullptr, &t->planH2);
        jobs->add(TJ_OMEGA, W, t->omW, nullptr, &t->planW2);
        return PB_OK;
    }
    (void)stream;
    jobs->add(TJ_TWIDDLES, H, t->twH, nullptr, nullptr);
    jobs->add(TJ_TWIDDLES, W, t->twW, nullptr, nullptr);
    return PB_OK;
}

static void poly_coeffs(double alpha, double beta, float* o) {
    // deblurring.py:160-162 (Python doubles, rounded to float32 when they meet the tensor)
    o[0] = (float)(alpha / 2 - beta + 2);
    o[1] = (float)(3 * beta - alpha - 6);
    o[2] = (float)(5 - 3 * beta + alpha / 2);
    o[3] = (float)beta;
}

// The same polynomial in powers of D = K - I.  a3 + a2 + a1 + b = 1 for every (alpha, beta), so
//     a3 K^3 + a2 K^2 + a1 K + b  =  I + c1 D + c2 D^2 + c3 D^3,   c1 = 3 a3 + 2 a2 + a1, c2 = 3 a3 + a2, c3 = a3.
// The engines evaluate this form: for the near-delta kernels Polyblur meets, D (*) v is small, so the
// Horner steps never subtract O(5) quantities that nearly cancel (the K form loses ~2 bits to that;
// measured: 4.1e-6 -> see profiles/r01_parity_report.jsonl).  Algebraically identical.
static void poly_coeffs_d(double alpha, double beta, float* o) {
    const double a3 = alpha / 2 - beta + 2, a2 = 3 * beta - alpha - 6, a1 = 5 - 3 * beta + alpha / 2;
    o[0] = (float)a3;                       // c3
    o[1] = (float)(3 * a3 + a2);            // c2
    o[2] = (float)(3 * a3 + 2 * a2 + a1);   // c1
    o[3] = 1.0f;                            // identity term
}

static int estimate_into(const float* img, int B, int C, int H, int W, double c, double b, double q, uint32_t flags,
                         float* est, char* ws, const Workspace& L, const Tables& T, int ksize,
                         float tap_thr, int engine, int fft_radius_min, cudaStream_t stream) {
    unsigned* stats = reinterpret_cast<unsigned*>(ws + L.off_stats);
    ImgKernel* kern = reinterpret_cast<ImgKernel*>(ws + L.off_kern);
    float* gray = reinterpret_cast<float*>(ws + L.off_gray);
    float* gy = reinterpret_cast<float*>(ws + L.off_gy);
    int* cls = reinterpret_cast<int*>(ws + L.off_cls);
    int rc;
    if ((rc = launch_init_stats(stats, B, stream))) return rc;
    if (q > 0) {
        // quantile normalisation: un-normalised gray -> quantiles -> the rows kernel normalises while
        // loading and writes the normalised plane (its min / max are then 0 and 1, so k_params'
        // division by the range is the identity); needs the fft2 path
        if (!T.fast) {
            set_error("q > 0 needs image sides whose prime factors are <= 13 (got %d x %d)", H, W);
            return PB_ERR_UNSUPPORTED;
        }
        float* ghat = reinterpret_cast<float*>(ws + L.off_ghat);
        float* qrange = nullptr;
        if ((rc = launch_quantile_range(img, gray, B, C, H, W, q, ws + L.off_q, &qrange, stream))) return rc;
        if ((rc = launch_rows2(true, gray, ghat, gy, stats, B, 1, H, W, T.planW2, T.twW, T.omW, qrange, stream))) return rc;
        if ((rc = launch_cols2(true, ghat, gy, nullptr, stats, B, H, W, T.planH2, T.twH, T.omH,
                               (flags & PB_FLAG_DISCARD_SATURATION) ? 1 : 0, gray, stream)))
            return rc;
        return launch_params(stats, kern, est, nullptr, nullptr, nullptr, nullptr, nullptr, 0, B, ksize,
                             (float)(c * c), (float)(b * b), tap_thr, engine, fft_radius_min, cls, stream);
    }
    if (T.fast) {
        // rows first (fully coalesced read of the iterate), `gy` holds d g / d x here
        if ((rc = launch_rows2(true, img, gray, gy, stats, B, C, H, W, T.planW2, T.twW, T.omW, nullptr, stream))) return rc;
        if ((rc = launch_cols2(true, gray, gy, nullptr, stats, B, H, W, T.planH2, T.twH, T.omH,
                               (flags & PB_FLAG_DISCARD_SATURATION) ? 1 : 0, nullptr, stream)))
            return rc;
        return launch_params(stats, kern, est, nullptr, nullptr, nullptr, nullptr, nullptr, 0, B, ksize,
                             (float)(c * c), (float)(b * b), tap_thr, engine, fft_radius_min, cls, stream);
    }
    if ((rc = launch_cols(true, img, gray, gy, stats, B, C, H, W, T.planH, T.twH, stream))) return rc;
    if ((rc = launch_rows(true, gray, gy, nullptr, stats, B, H, W, T.planW, T.twW,
                          (flags & PB_FLAG_DISCARD_SATURATION) ? 1 : 0, stream)))
        return rc;
    return launch_params(stats, kern, est, nullptr, nullptr, nullptr, nullptr, nullptr, 0, B, ksize,
                         (float)(c * c), (float)(b * b), tap_thr, engine, fft_radius_min, cls, stream);
}

// filters.fourier_gradients of a stack of planes; gx and / or gy may be NULL.
static int gradients_into(const float* planes, float* gx, float* gy, int nplanes, int H, int W, const Tables& T,
                          cudaStream_t stream) {
    int rc;
    if (T.fast) {
        if (gx && (rc = launch_rows2(false, planes, nullptr, gx, nullptr, nplanes, 1, H, W, T.planW2, T.twW, T.omW, nullptr, stream)))
            return rc;
        if (gy && (rc = launch_cols2(false, planes, nullptr, gy, nullptr, nplanes, H, W, T.planH2, T.twH, T.omH, 0, nullptr, stream)))
            return rc;
        return PB_OK;
    }
    if (gy && (rc = launch_cols(false, planes, nullptr, gy, nullptr, nplanes, 1, H, W, T.planH, T.twH, stream))) return rc;
    if (gx && (rc = launch_rows(false, planes, nullptr, gx, nullptr, nplanes, H, W, T.planW, T.twW, 0, stream))) return rc;
    return PB_OK;
}

static SrcGeom default_geom(int H, int W) {
    SrcGeom g;
    g.Hin = H;
    g.Win = W;
    g.off = 0;
    g.pad = -1;
    g.clamp_out = 1;
    return g;
}

// Runs every deconvolution engine over its class of images (lists filled by k_params).
static int deconv_all(const float* img, float* out, int B, int C, int H, int W, const float* coef, char* ws,
                      const Workspace& L, const FftEngineTables* F, const SrcGeom& G, cudaStream_t stream) {
    const ImgKernel* kern = reinterpret_cast<const ImgKernel*>(ws + L.off_kern);
    const int* cls = reinterpret_cast<const int*>(ws + L.off_cls);
    int rc;
    for (int k = PB_CLS_N11; k <= PB_CLS_N22; ++k)
        if ((rc = launch_deconv_narrow(k, img, out, kern, cls + PB_CLS_COUNT_STRIDE + k * B, cls + k, B, C, H, W,
                                       coef[0], coef[1], coef[2], coef[3], G, stream)))
            return rc;
    if ((rc = launch_deconv_spatial(img, out, kern, cls + PB_CLS_COUNT_STRIDE + PB_CLS_TILED4 * B, cls + PB_CLS_TILED4,
                                    B, C, H, W, coef[0], coef[1], coef[2], coef[3], G, 4, stream)))
        return rc;
    if ((rc = launch_deconv_spatial(img, out, kern, cls + PB_CLS_COUNT_STRIDE + PB_CLS_TILED * B, cls + PB_CLS_TILED,
                                    B, C, H, W, coef[0], coef[1], coef[2], coef[3], G, PB_PAD, stream)))
        return rc;
    if (F)
        return launch_deconv_fft(img, out, kern, cls + PB_CLS_COUNT_STRIDE + PB_CLS_FFT * B, cls + PB_CLS_FFT, B, C,
                                 H, W, *F, coef[0], coef[1], coef[2], coef[3], G, stream);
    return PB_OK;
}

// inverse_filtering_rank3 (deblurring.py:211-239) for the kernels k_params left in the workspace:
// [edgetaper on the explicitly padded image] -> polynomial on the torus -> crop -> [halo masking] -> clamp.
// grads_of_tapered: halo masking without an explicit grad_img takes the gradients of the image it is handed
// (deblurring.py:200-203), which inverse_filtering_rank3 has already cropped out of the padded AND tapered
// plane (deblurring.py:237-238): g0x / g0y / nM are then filled here, after the taper.
static int deconv_with_options(const float* src, float* dst, int B, int C, int H, int W, const float* coef, int ksize,
                               uint32_t flags, const float* g0x, const float* g0y, const float* nM, float* ox,
                               char* ws, const Workspace& L, const Tables& T, const FftEngineTables* F,
                               cudaStream_t stream, bool grads_of_tapered = false) {
    const bool halo = (flags & PB_FLAG_REMOVE_HALO) != 0;
    const bool taper = (flags & PB_FLAG_EDGETAPER) != 0;
    const int pad = ksize / 2;
    const int Hp = H + 2 * pad, Wp = W + 2 * pad;
    const int planes = B * C;
    const ImgKernel* kern = reinterpret_cast<const ImgKernel*>(ws + L.off_kern);
    int rc;
    // [edgetaping] taper the explicitly padded image; the engines then read that padded plane
    SrcGeom G = default_geom(H, W);
    const float* dec_in = src;
    if (taper) {
        float* padA = reinterpret_cast<float*>(ws + L.off_padA);
        float* padB = reinterpret_cast<float*>(ws + L.off_padB);
        float *v = nullptr, *tapered = nullptr;
        if ((rc = launch_pad_replicate(src, padA, padB, planes, H, W, pad, stream))) return rc;
        if ((rc = launch_edgetaper_weights(kern, ws + L.off_et, B, Hp, Wp, (flags & PB_FLAG_EDGETAPER_BATCHMAX) ? 1 : 0,
                                           &v, stream)))
            return rc;
        if ((rc = launch_edgetaper_passes(padA, padB, kern, v, B, C, Hp, Wp, 3, &tapered, stream))) return rc;
        G.Hin = Hp;
        G.Win = Wp;
        G.off = pad;
        G.pad = 0;
        dec_in = tapered;
        if (halo && grads_of_tapered) {
            // dst is free until the engines write it: it holds the cropped tapered image for a moment
            float* wx = reinterpret_cast<float*>(ws + L.off_g0x);
            float* wy = reinterpret_cast<float*>(ws + L.off_g0y);
            float* wn = reinterpret_cast<float*>(ws + L.off_nm);
            if ((rc = launch_crop(tapered, dst, planes, H, W, pad, stream))) return rc;
            if ((rc = gradients_into(dst, wx, wy, planes, H, W, T, stream))) return rc;
            if ((rc = launch_halo_norm(wx, wy, wn + planes, wn, planes, (size_t)H * W, stream))) return rc;
        }
    }
    G.clamp_out = (halo || (flags & PB_FLAG_NO_CLAMP)) ? 0 : 1;
    if ((rc = deconv_all(dec_in, dst, B, C, H, W, coef, ws, L, F, G, stream))) return rc;
    if (halo) {
        // halo_masking (deblurring.py:193-208): only d imout / dx is needed (M uses gy*gy, :174)
        if ((rc = gradients_into(dst, ox, nullptr, planes, H, W, T, stream))) return rc;
        if ((rc = launch_halo_apply(dst, dec_in, (size_t)G.Hin * G.Win, G.Win, G.off, g0x, g0y, ox, nM, planes, H, W,
                                    stream)))
            return rc;
    }
    return PB_OK;
}

}  // namespace pb

using namespace pb;

extern "C" {

int pb_version(void) { return PB_VERSION; }

const char* pb_last_error(void) { return g_err; }

void pb_default_params(pb_params* p) {
    if (!p) return;
    memset(p, 0, sizeof(*p));
    p->n_iter = 1;
    p->c = 0.352;
    p->b = 0.768;
    p->alpha = 2;
    p->beta = 3;
    p->sigma_s = 2.0;
    p->sigma_r = 0.8;
    p->ker_size = 25;
    p->q = 0.0;
    p->flags = 0;
    p->engine = PB_ENGINE_AUTO;
    p->tap_rel_threshold = 0.0f;
    p->chunk_images = 0;
}

void pb_polynomial_coefficients(double alpha, double beta, float* out4) { poly_coeffs(alpha, beta, out4); }

void pb_keys_weights(float* out210) { keys_weights_host(out210); }

int pb_fft_plan(int n, int* radices) {
    // the register-radix core (fft2.cuh) runs every length whose prime factors are <= 13; the any-length
    // Stockham core (fft.cuh) is the fallback
    Fft2Plan p2;
    if (n >= 2 && radices && make_fft2_plan(n, &p2) == 0) {
        for (int i = 0; i < p2.ns; ++i) radices[i] = p2.radix[i];
        return p2.ns;
    }
    FftPlan p;
    if (!radices || make_fft_plan(n, &p) != PB_OK) {
        set_error("cannot plan FFT of length %d", n);
        return PB_ERR_ARG;
    }
    for (int i = 0; i < p.ns; ++i) radices[i] = p.radix[i];
    return p.ns;
}

// images one engine pass of pb_polyblur_f32 works on: pb_params.chunk_images bounds it (and with it the workspace)
static int group_size(int B, const pb_params* p) {
    return (p && p->chunk_images > 0 && p->chunk_images < B) ? p->chunk_images : B;
}

size_t pb_workspace_bytes(int B, int C, int H, int W, const pb_params* p) {
    if (B < 1 || C < 1 || H < 1 || W < 1) return 0;
    B = group_size(B, p);
    return layout(B, C, H, W, p ? p->n_iter : 1, p ? p->ker_size : PB_KS, p ? p->engine : PB_ENGINE_AUTO,
                  p ? p->flags : 0, p ? p->q : 0.0).total;
}

static int validate_params(const pb_params* p) {
    if (!p) {
        set_error("params is NULL");
        return PB_ERR_ARG;
    }
    if (p->n_iter < 0) {
        set_error("n_iter must be >= 0");
        return PB_ERR_ARG;
    }
    if (p->ker_size < 1 || p->ker_size > PB_KSIZE_MAX || (p->ker_size & 1) == 0) {
        set_error("ker_size must be odd and <= %d (got %d)", PB_KSIZE_MAX, p->ker_size);
        return PB_ERR_ARG;
    }
    if (p->q < 0.0 || p->q >= 0.5) {
        set_error("q must be in [0, 0.5)");
        return PB_ERR_ARG;
    }
    if ((p->flags & PB_FLAG_PREFILTER) && (p->flags & PB_FLAG_PREFILTER_RF)) {
        set_error("choose one prefilter: PB_FLAG_PREFILTER (bilateral) or PB_FLAG_PREFILTER_RF");
        return PB_ERR_ARG;
    }
    return PB_OK;
}

static int polyblur_group(const float* in, float* out, int B, int C, int H, int W, const pb_params* p, void* workspace,
                          size_t workspace_bytes, float* est_out, int est_batch, void* stream_);

int pb_polyblur_f32(const float* in, float* out, int B, int C, int H, int W, const pb_params* p,
                    void* workspace, size_t workspace_bytes, float* est_out, void* stream_) {
    int rc;
    if ((rc = check_shape(B, C, H, W))) return rc;
    if ((rc = validate_params(p))) return rc;
    const int G = group_size(B, p);
    if (G == B) return polyblur_group(in, out, B, C, H, W, p, workspace, workspace_bytes, est_out, B, stream_);
    // chunk_images: the batch goes through in groups of G images, each running the whole loop in a workspace sized
    // for G (images are independent; the one batch-coupled option is the reference's edgetaper normalisation)
    if (p->flags & PB_FLAG_EDGETAPER_BATCHMAX) {
        set_error("chunk_images cannot be combined with PB_FLAG_EDGETAPER_BATCHMAX (a batch-global maximum)");
        return PB_ERR_ARG;
    }
    const size_t per = (size_t)C * H * W;
    for (int b0 = 0; b0 < B; b0 += G) {
        const int n = (B - b0 < G) ? B - b0 : G;
        float* est = est_out ? est_out + (size_t)b0 * PB_EST_STRIDE : nullptr;
        if ((rc = polyblur_group(in + b0 * per, out + b0 * per, n, C, H, W, p, workspace, workspace_bytes, est, B, stream_)))
            return rc;
    }
    return PB_OK;
}

// est_batch: batch size of the est_out array (rows of one iteration are est_batch records apart)
static int polyblur_group(const float* in, float* out, int B, int C, int H, int W, const pb_params* p, void* workspace,
                          size_t workspace_bytes, float* est_out, int est_batch, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    int rc;
    if (!in || !out || in == out) {
        set_error("in/out must be distinct non-null device pointers");
        return PB_ERR_ARG;
    }
    const size_t bytes = (size_t)B * C * H * W * sizeof(float);
    if (p->n_iter == 0) {
        PB_CUDA_TRY(cudaMemcpyAsync(out, in, bytes, cudaMemcpyDeviceToDevice, stream));
        return PB_OK;
    }
    const Workspace L = layout(B, C, H, W, p->n_iter, p->ker_size, p->engine, p->flags, p->q);
    if ((rc = check_ws(workspace, workspace_bytes, L.total))) return rc;
    if (p->engine == PB_ENGINE_FFT && !L.has_fft) {
        set_error("the FFT engine does not support %d x %d (ker_size %d)", H, W, p->ker_size);
        return PB_ERR_UNSUPPORTED;
    }
    char* ws = static_cast<char*>(workspace);
    Tables T;
    FftEngineTables F;
    if ((rc = upload_constants(stream))) return rc;
    TableJobs jobs;
    if ((rc = prepare_tables(ws, L, H, W, &T, &jobs, stream))) return rc;
    if (L.has_fft && (rc = fft_engine_prepare(ws + L.off_fft, L.fft, &F, &jobs))) return rc;
    if ((rc = launch_table_jobs(jobs, stream))) return rc;
    float coef[4];
    poly_coeffs_d(p->alpha, p->beta, coef);
    const float thr = p->tap_rel_threshold > 0 ? p->tap_rel_threshold : 1e-8f;
    float* tmp = reinterpret_cast<float*>(ws + L.off_tmp);
    const bool prefilter = (p->flags & (PB_FLAG_PREFILTER | PB_FLAG_PREFILTER_RF)) != 0;
    const bool halo = (p->flags & PB_FLAG_REMOVE_HALO) != 0;
    const int planes = B * C;
    const size_t plane = (size_t)H * W;
    float* smooth = reinterpret_cast<float*>(ws + L.off_smooth);
    float* g0x = reinterpret_cast<float*>(ws + L.off_g0x);
    float* g0y = reinterpret_cast<float*>(ws + L.off_g0y);
    float* ox = reinterpret_cast<float*>(ws + L.off_ox);
    float* nM = reinterpret_cast<float*>(ws + L.off_nm);
    if (halo) {
        // grad_img of the ORIGINAL input, all channels, once (deblurring.py:61) and its energy per plane
        if ((rc = gradients_into(in, g0x, g0y, planes, H, W, T, stream))) return rc;
        if ((rc = launch_halo_norm(g0x, g0y, nM + planes, nM, planes, plane, stream))) return rc;
    }
    const float* cur = in;
    for (int it = 0; it < p->n_iter; ++it) {
        float* dst = ((p->n_iter - 1 - it) & 1) ? tmp : out;
        float* est = est_out ? est_out + (size_t)it * est_batch * PB_EST_STRIDE : nullptr;
        if ((rc = estimate_into(cur, B, C, H, W, p->c, p->b, p->q, p->flags, est, ws, L, T, p->ker_size, thr, p->engine,
                                L.has_fft ? PB_FFT_RADIUS_MIN : (1 << 30), stream)))
            return rc;
        // [prefiltering] deconvolve the smooth component only (deblurring.py:80-84)
        const float* src = cur;
        if (prefilter) {
            if (p->flags & PB_FLAG_PREFILTER_RF) {
                if ((rc = launch_recursive_filter(cur, nullptr, smooth, B, C, H, W, p->sigma_s, p->sigma_r, 1,
                                                  ws + L.off_rf, stream)))
                    return rc;
            } else if ((rc = launch_bilateral(cur, smooth, planes, H, W, 5.0f, 0.1f, stream))) {
                return rc;
            }
            src = smooth;
        }
        if ((rc = deconv_with_options(src, dst, B, C, H, W, coef, p->ker_size, p->flags, g0x, g0y, nM, ox, ws, L, T,
                                      L.has_fft ? &F : nullptr, stream)))
            return rc;
        if (prefilter && (rc = launch_residual_add(dst, cur, smooth, (size_t)planes * plane, stream))) return rc;
        cur = dst;
    }
    return PB_OK;
}

int pb_fourier_gradients_f32(const float* img, float* gx, float* gy, int B, int C, int H, int W,
                             void* workspace, size_t workspace_bytes, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    int rc;
    if ((rc = check_shape(B, C, H, W))) return rc;
    if (!img || !gx || !gy) {
        set_error("null pointer");
        return PB_ERR_ARG;
    }
    const Workspace L = layout(B, C, H, W, 1);
    if ((rc = check_ws(workspace, workspace_bytes, L.total))) return rc;
    char* ws = static_cast<char*>(workspace);
    Tables T;
    TableJobs jobs;
    if ((rc = prepare_tables(ws, L, H, W, &T, &jobs, stream))) return rc;
    if ((rc = launch_table_jobs(jobs, stream))) return rc;
    if (T.fast) {
        if ((rc = launch_rows2(false, img, nullptr, gx, nullptr, B * C, 1, H, W, T.planW2, T.twW, T.omW, nullptr, stream)))
            return rc;
        return launch_cols2(false, img, nullptr, gy, nullptr, B * C, H, W, T.planH2, T.twH, T.omH, 0, nullptr, stream);
    }
    if ((rc = launch_cols(false, img, nullptr, gy, nullptr, B * C, 1, H, W, T.planH, T.twH, stream))) return rc;
    return launch_rows(false, img, nullptr, gx, nullptr, B * C, H, W, T.planW, T.twW, 0, stream);
}

int pb_estimate_f32(const float* img, int B, int C, int H, int W, double c, double b, double q,
                    uint32_t flags, float* est, void* workspace, size_t workspace_bytes, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    int rc;
    if ((rc = check_shape(B, C, H, W))) return rc;
    if (!img || !est) {
        set_error("null pointer");
        return PB_ERR_ARG;
    }
    if (q < 0.0 || q >= 0.5) {
        set_error("q must be in [0, 0.5)");
        return PB_ERR_ARG;
    }
    const Workspace L = layout(B, C, H, W, 1, PB_KS, PB_ENGINE_SPATIAL, 0, q);
    if ((rc = check_ws(workspace, workspace_bytes, L.total))) return rc;
    char* ws = static_cast<char*>(workspace);
    Tables T;
    if ((rc = upload_constants(stream))) return rc;
    TableJobs jobs;
    if ((rc = prepare_tables(ws, L, H, W, &T, &jobs, stream))) return rc;
    if ((rc = launch_table_jobs(jobs, stream))) return rc;
    return estimate_into(img, B, C, H, W, c, b, q, flags, est, ws, L, T, PB_KS, 1e-8f, PB_ENGINE_SPATIAL, 1 << 30, stream);
}

int pb_make_kernel_f32(const float* theta, const float* sigma, const float* rho, int B, int ksize,
                       float* kernel, void* workspace, size_t workspace_bytes, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (B < 1 || !theta || !sigma || !rho || !kernel || ksize < 1 || ksize > PB_KS || !(ksize & 1)) {
        set_error("bad arguments to pb_make_kernel_f32");
        return PB_ERR_ARG;
    }
    const size_t need = (size_t)B * sizeof(ImgKernel);
    int rc;
    if ((rc = check_ws(workspace, workspace_bytes, need))) return rc;
    return launch_params(nullptr, static_cast<ImgKernel*>(workspace), nullptr, theta, sigma, rho, nullptr,
                         kernel, 1, B, ksize, 0.f, 0.f, 1e-8f, PB_ENGINE_SPATIAL, 1 << 30, nullptr, stream);
}

int pb_deconv_f32(const float* img, float* out, int B, int C, int H, int W, const float* kernel, int ksize,
                  double alpha, double beta, int engine, void* workspace, size_t workspace_bytes,
                  void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    int rc;
    if ((rc = check_shape(B, C, H, W))) return rc;
    if (!img || !out || !kernel || img == out || ksize < 1 || ksize > PB_KS || !(ksize & 1)) {
        set_error("bad arguments to pb_deconv_f32 (ksize must be odd and <= 25)");
        return PB_ERR_ARG;
    }
    const Workspace L = layout(B, C, H, W, 1, ksize, engine);
    if ((rc = check_ws(workspace, workspace_bytes, L.total))) return rc;
    if (engine == PB_ENGINE_FFT && !L.has_fft) {
        set_error("the FFT engine does not support %d x %d (ker_size %d)", H, W, ksize);
        return PB_ERR_UNSUPPORTED;
    }
    char* ws = static_cast<char*>(workspace);
    ImgKernel* kern = reinterpret_cast<ImgKernel*>(ws + L.off_kern);
    int* cls = reinterpret_cast<int*>(ws + L.off_cls);
    FftEngineTables F;
    TableJobs jobs;
    if (L.has_fft && (rc = fft_engine_prepare(ws + L.off_fft, L.fft, &F, &jobs))) return rc;
    if ((rc = launch_table_jobs(jobs, stream))) return rc;
    if ((rc = launch_params(nullptr, kern, nullptr, nullptr, nullptr, nullptr, kernel, nullptr, 2, B, ksize,
                            0.f, 0.f, 1e-8f, engine, L.has_fft ? PB_FFT_RADIUS_MIN : (1 << 30), cls, stream)))
        return rc;
    float coef[4];
    poly_coeffs_d(alpha, beta, coef);
    return deconv_all(img, out, B, C, H, W, coef, ws, L, L.has_fft ? &F : nullptr, default_geom(H, W), stream);
}

int pb_deconv_ex_f32(const float* img, float* out, int B, int C, int H, int W, const float* kernel, int ksize,
                     double alpha, double beta, int engine, uint32_t flags, const float* grad_x, const float* grad_y,
                     void* workspace, size_t workspace_bytes, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    int rc;
    if ((rc = check_shape(B, C, H, W))) return rc;
    if (!img || !out || !kernel || img == out || ksize < 1 || ksize > PB_KS || !(ksize & 1)) {
        set_error("bad arguments to pb_deconv_ex_f32 (ksize must be odd and <= 25)");
        return PB_ERR_ARG;
    }
    if ((grad_x == nullptr) != (grad_y == nullptr)) {
        set_error("grad_x and grad_y must be given together");
        return PB_ERR_ARG;
    }
    flags &= PB_FLAG_REMOVE_HALO | PB_FLAG_EDGETAPER | PB_FLAG_EDGETAPER_BATCHMAX | PB_FLAG_NO_CLAMP;
    if ((flags & PB_FLAG_NO_CLAMP) && (flags & PB_FLAG_REMOVE_HALO)) {
        set_error("PB_FLAG_NO_CLAMP cannot be combined with halo masking (which clamps itself)");
        return PB_ERR_ARG;
    }
    const Workspace L = layout(B, C, H, W, 1, ksize, engine, flags);
    if ((rc = check_ws(workspace, workspace_bytes, L.total))) return rc;
    if (engine == PB_ENGINE_FFT && !L.has_fft) {
        set_error("the FFT engine does not support %d x %d (ker_size %d)", H, W, ksize);
        return PB_ERR_UNSUPPORTED;
    }
    char* ws = static_cast<char*>(workspace);
    ImgKernel* kern = reinterpret_cast<ImgKernel*>(ws + L.off_kern);
    int* cls = reinterpret_cast<int*>(ws + L.off_cls);
    Tables T;
    FftEngineTables F;
    TableJobs jobs;
    if ((rc = prepare_tables(ws, L, H, W, &T, &jobs, stream))) return rc;
    if (L.has_fft && (rc = fft_engine_prepare(ws + L.off_fft, L.fft, &F, &jobs))) return rc;
    if ((rc = launch_table_jobs(jobs, stream))) return rc;
    if ((rc = launch_params(nullptr, kern, nullptr, nullptr, nullptr, nullptr, kernel, nullptr, 2, B, ksize, 0.f, 0.f,
                            1e-8f, engine, L.has_fft ? PB_FFT_RADIUS_MIN : (1 << 30), cls, stream)))
        return rc;
    float* g0x = reinterpret_cast<float*>(ws + L.off_g0x);
    float* g0y = reinterpret_cast<float*>(ws + L.off_g0y);
    float* ox = reinterpret_cast<float*>(ws + L.off_ox);
    float* nM = reinterpret_cast<float*>(ws + L.off_nm);
    const float *gx = g0x, *gy = g0y;
    bool from_tapered = false;
    if (flags & PB_FLAG_REMOVE_HALO) {
        // grad_img defaults to the gradients of img itself (deblurring.py:200-203)
        if (grad_x) {
            gx = grad_x;
            gy = grad_y;
        } else if (flags & PB_FLAG_EDGETAPER) {
            from_tapered = true;                 // gradients of the cropped tapered image, computed after the taper
        } else if ((rc = gradients_into(img, g0x, g0y, B * C, H, W, T, stream))) {
            return rc;
        }
        if (!from_tapered && (rc = launch_halo_norm(gx, gy, nM + B * C, nM, B * C, (size_t)H * W, stream))) return rc;
    }
    float coef[4];
    poly_coeffs_d(alpha, beta, coef);
    return deconv_with_options(img, out, B, C, H, W, coef, ksize, flags, gx, gy, nM, ox, ws, L, T,
                               L.has_fft ? &F : nullptr, stream, from_tapered);
}

// workspace of the backward pass: the engines' workspace for a (H+2P) x (W+2P) "image", the
// embedded gradient, the filtered plane and the rotated kernels
struct VjpLayout {
    Workspace eng;
    size_t off_z, off_t, off_k, total;
};
static VjpLayout vjp_layout(int B, int C, int H, int W, int ksize, int engine) {
    VjpLayout v;
    const int pad = ksize / 2;
    const int Hp = H + 2 * pad, Wp = W + 2 * pad;
    v.eng = layout(B, C, Hp, Wp, 1, ksize, engine);
    size_t o = align_up(v.eng.total, 256);
    const size_t plane_bytes = align_up((size_t)B * C * Hp * Wp * sizeof(float), 256);
    v.off_z = o;
    o += plane_bytes;
    v.off_t = o;
    o += plane_bytes;
    v.off_k = o;
    o += align_up((size_t)B * ksize * ksize * sizeof(float), 256);
    v.total = o;
    return v;
}

size_t pb_deconv_vjp_workspace_bytes(int B, int C, int H, int W, int ksize, int engine) {
    if (B < 1 || C < 1 || H < 1 || W < 1 || ksize < 1 || ksize > PB_KS || !(ksize & 1)) return 0;
    return vjp_layout(B, C, H, W, ksize, engine).total;
}

int pb_deconv_vjp_f32(const float* grad_out, const float* preclamp, float* grad_img, int B, int C, int H, int W,
                      const float* kernel, int ksize, double alpha, double beta, int engine, void* workspace,
                      size_t workspace_bytes, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    int rc;
    if ((rc = check_shape(B, C, H, W))) return rc;
    if (!grad_out || !grad_img || !kernel || ksize < 1 || ksize > PB_KS || !(ksize & 1)) {
        set_error("bad arguments to pb_deconv_vjp_f32 (ksize must be odd and <= 25)");
        return PB_ERR_ARG;
    }
    const int pad = ksize / 2;
    const int Hp = H + 2 * pad, Wp = W + 2 * pad;
    if ((rc = check_shape(B, C, Hp, Wp))) return rc;
    const VjpLayout V = vjp_layout(B, C, H, W, ksize, engine);
    if ((rc = check_ws(workspace, workspace_bytes, V.total))) return rc;
    const Workspace& L = V.eng;
    if (engine == PB_ENGINE_FFT && !L.has_fft) {
        set_error("the FFT engine does not support %d x %d (ker_size %d)", Hp, Wp, ksize);
        return PB_ERR_UNSUPPORTED;
    }
    char* ws = static_cast<char*>(workspace);
    ImgKernel* kern = reinterpret_cast<ImgKernel*>(ws + L.off_kern);
    int* cls = reinterpret_cast<int*>(ws + L.off_cls);
    float* z = reinterpret_cast<float*>(ws + V.off_z);
    float* t = reinterpret_cast<float*>(ws + V.off_t);
    float* kf = reinterpret_cast<float*>(ws + V.off_k);
    FftEngineTables F;
    TableJobs jobs;
    if (L.has_fft && (rc = fft_engine_prepare(ws + L.off_fft, L.fft, &F, &jobs))) return rc;
    if ((rc = launch_table_jobs(jobs, stream))) return rc;
    if ((rc = launch_flip_kernels(kernel, kf, B, ksize, stream))) return rc;
    if ((rc = launch_params(nullptr, kern, nullptr, nullptr, nullptr, nullptr, kf, nullptr, 2, B, ksize, 0.f, 0.f,
                            1e-8f, engine, L.has_fft ? PB_FFT_RADIUS_MIN : (1 << 30), cls, stream)))
        return rc;
    if ((rc = launch_vjp_embed(grad_out, preclamp, z, B * C, H, W, pad, stream))) return rc;
    // the embedded plane is its own torus: pure wrap-around source, every padded position is an output
    SrcGeom G;
    G.Hin = Hp;
    G.Win = Wp;
    G.off = 0;
    G.pad = 0;
    G.clamp_out = 0;
    float coef[4];
    poly_coeffs_d(alpha, beta, coef);
    if ((rc = deconv_all(z, t, B, C, Hp, Wp, coef, ws, L, L.has_fft ? &F : nullptr, G, stream))) return rc;
    return launch_vjp_fold(t, grad_img, B * C, H, W, pad, stream);
}

// ---- backward pass through the estimator and the kernel argument (backward.cu) -----------------
struct BwLayout {
    Workspace eng;                       // engines + tables, sized for the (H+2P) x (W+2P) frame
    size_t off_pad, off_v, off_q[4], off_keys, off_stats, off_sums, total;
};
static BwLayout bw_layout(int B, int C, int H, int W, int ksize, int engine) {
    BwLayout v;
    const int pad = ksize / 2;
    const int Hp = H + 2 * pad, Wp = W + 2 * pad;
    v.eng = layout(B, C, Hp, Wp, 1, ksize, engine);
    size_t o = align_up(v.eng.total, 256);
    auto take = [&](size_t bytes) {
        size_t at = o;
        o = align_up(o + bytes, 256);
        return at;
    };
    v.off_pad = take((size_t)B * C * Hp * Wp * sizeof(float));
    v.off_v = take((size_t)B * C * Hp * Wp * sizeof(float));
    for (int i = 0; i < 4; ++i) v.off_q[i] = take((size_t)B * H * W * sizeof(float));
    v.off_keys = take((size_t)B * 7 * sizeof(unsigned long long));
    v.off_stats = take((size_t)B * 4 * sizeof(unsigned));
    v.off_sums = take((size_t)B * 2 * sizeof(double));
    v.total = o;
    return v;
}

size_t pb_backward_workspace_bytes(int B, int C, int H, int W, int ksize, int engine) {
    if (B < 1 || C < 1 || H < 1 || W < 1 || ksize < 1 || ksize > PB_KS || !(ksize & 1)) return 0;
    return bw_layout(B, C, H, W, ksize, engine).total;
}

int pb_estimate_trace_f32(const float* img, int B, int C, int H, int W, float* trace_f, int* trace_pos,
                          void* workspace, size_t workspace_bytes, void* stream_) {
    return pb_estimate_trace_ex_f32(img, B, C, H, W, 0, trace_f, trace_pos, workspace, workspace_bytes, stream_);
}

int pb_estimate_trace_ex_f32(const float* img, int B, int C, int H, int W, uint32_t flags, float* trace_f,
                             int* trace_pos, void* workspace, size_t workspace_bytes, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (flags & ~(uint32_t)PB_FLAG_DISCARD_SATURATION) {
        set_error("pb_estimate_trace_ex_f32 understands PB_FLAG_DISCARD_SATURATION only");
        return PB_ERR_UNSUPPORTED;
    }
    int rc;
    if ((rc = check_shape(B, C, H, W))) return rc;
    if (!img || !trace_f || !trace_pos) {
        set_error("null pointer");
        return PB_ERR_ARG;
    }
    const BwLayout V = bw_layout(B, C, H, W, PB_KS, PB_ENGINE_AUTO);
    if ((rc = check_ws(workspace, workspace_bytes, V.total))) return rc;
    char* ws = static_cast<char*>(workspace);
    Tables T;
    if ((rc = upload_constants(stream))) return rc;
    TableJobs jobs;
    if ((rc = prepare_tables(ws, V.eng, H, W, &T, &jobs, stream))) return rc;
    if ((rc = launch_table_jobs(jobs, stream))) return rc;
    float* g = reinterpret_cast<float*>(ws + V.off_q[0]);
    float* gn = reinterpret_cast<float*>(ws + V.off_q[1]);
    float* gx = reinterpret_cast<float*>(ws + V.off_q[2]);
    float* gy = reinterpret_cast<float*>(ws + V.off_q[3]);
    unsigned* stats = reinterpret_cast<unsigned*>(ws + V.off_stats);
    unsigned long long* keys = reinterpret_cast<unsigned long long*>(ws + V.off_keys);
    if ((rc = launch_bw_trace(img, g, gn, stats, keys, B, C, H, W, stream))) return rc;
    if ((rc = gradients_into(gn, gx, gy, B, H, W, T, stream))) return rc;
    return launch_bw_dirmax(gx, gy, g, keys, stats, trace_f, trace_pos, B, H, W,
                            (flags & PB_FLAG_DISCARD_SATURATION) ? 1 : 0, stream);
}

int pb_kernel_grad_f32(const float* img, const float* grad_out, const float* preclamp, int B, int C, int H, int W,
                       const float* kernel, int ksize, double alpha, double beta, int engine, float* kernel_grad,
                       void* workspace, size_t workspace_bytes, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    int rc;
    if ((rc = check_shape(B, C, H, W))) return rc;
    if (!img || !grad_out || !kernel || !kernel_grad || ksize < 1 || ksize > PB_KS || !(ksize & 1)) {
        set_error("bad arguments to pb_kernel_grad_f32 (ksize must be odd and <= 25)");
        return PB_ERR_ARG;
    }
    const int pad = ksize / 2;
    const int Hp = H + 2 * pad, Wp = W + 2 * pad;
    if ((rc = check_shape(B, C, Hp, Wp))) return rc;
    const BwLayout V = bw_layout(B, C, H, W, ksize, engine);
    if ((rc = check_ws(workspace, workspace_bytes, V.total))) return rc;
    const Workspace& L = V.eng;
    if (engine == PB_ENGINE_FFT && !L.has_fft) {
        set_error("the FFT engine does not support %d x %d (ker_size %d)", Hp, Wp, ksize);
        return PB_ERR_UNSUPPORTED;
    }
    char* ws = static_cast<char*>(workspace);
    ImgKernel* kern = reinterpret_cast<ImgKernel*>(ws + L.off_kern);
    int* cls = reinterpret_cast<int*>(ws + L.off_cls);
    float* xp = reinterpret_cast<float*>(ws + V.off_pad);
    float* v = reinterpret_cast<float*>(ws + V.off_v);
    FftEngineTables F;
    TableJobs jobs;
    if (L.has_fft && (rc = fft_engine_prepare(ws + L.off_fft, L.fft, &F, &jobs))) return rc;
    if ((rc = launch_table_jobs(jobs, stream))) return rc;
    if ((rc = launch_params(nullptr, kern, nullptr, nullptr, nullptr, nullptr, kernel, nullptr, 2, B, ksize, 0.f, 0.f,
                            1e-8f, engine, L.has_fft ? PB_FFT_RADIUS_MIN : (1 << 30), cls, stream)))
        return rc;
    if ((rc = launch_pad_replicate(img, xp, xp, B * C, H, W, pad, stream))) return rc;
    // V = dP/dK (*) pad(x) on the padded torus: dP/dD = c1 + 2 c2 D + 3 c3 D^2 in the engines' basis D = K - I
    float coef[4], dq[4];
    poly_coeffs_d(alpha, beta, coef);
    dq[0] = 0.0f;
    dq[1] = 3.0f * coef[0];
    dq[2] = 2.0f * coef[1];
    dq[3] = coef[2];
    SrcGeom G;
    G.Hin = Hp;
    G.Win = Wp;
    G.off = 0;
    G.pad = 0;
    G.clamp_out = 0;
    if ((rc = deconv_all(xp, v, B, C, Hp, Wp, dq, ws, L, L.has_fft ? &F : nullptr, G, stream))) return rc;
    return launch_bw_kernel_grad(grad_out, preclamp, v, kernel_grad, B, C, H, W, ksize, stream);
}

int pb_estimator_vjp_f32(const float* img, const float* mbar, const float* trace_f, const int* trace_pos,
                         float* grad_img, int B, int C, int H, int W, void* workspace, size_t workspace_bytes,
                         void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    int rc;
    if ((rc = check_shape(B, C, H, W))) return rc;
    if (!img || !mbar || !trace_f || !trace_pos || !grad_img) {
        set_error("null pointer");
        return PB_ERR_ARG;
    }
    const BwLayout V = bw_layout(B, C, H, W, PB_KS, PB_ENGINE_AUTO);
    if ((rc = check_ws(workspace, workspace_bytes, V.total))) return rc;
    char* ws = static_cast<char*>(workspace);
    Tables T;
    if ((rc = upload_constants(stream))) return rc;
    TableJobs jobs;
    if ((rc = prepare_tables(ws, V.eng, H, W, &T, &jobs, stream))) return rc;
    if ((rc = launch_table_jobs(jobs, stream))) return rc;
    float* sgx = reinterpret_cast<float*>(ws + V.off_q[0]);
    float* sgy = reinterpret_cast<float*>(ws + V.off_q[1]);
    float* dx = reinterpret_cast<float*>(ws + V.off_q[2]);
    float* dy = reinterpret_cast<float*>(ws + V.off_q[3]);
    double* sums = reinterpret_cast<double*>(ws + V.off_sums);
    if ((rc = launch_bw_scatter(mbar, trace_f, trace_pos, sgx, sgy, B, H, W, stream))) return rc;
    // D^T = -D: the sign is applied where the two planes are combined
    if ((rc = gradients_into(sgx, dx, nullptr, B, H, W, T, stream))) return rc;
    if ((rc = gradients_into(sgy, nullptr, dy, B, H, W, T, stream))) return rc;
    return launch_bw_norm(dx, dy, img, trace_f, sums, grad_img, B, C, H, W, stream);
}

int pb_profile_begin(void) {
    std::lock_guard<std::mutex> lock(g_prof_mu);
    for (auto& r : g_prof) {
        g_event_pool.push_back(r.a);
        g_event_pool.push_back(r.b);
    }
    g_prof.clear();
    g_prof_on = true;
    return PB_OK;
}

int pb_profile_end(float* ms_per_class, int* launches_per_class, int max_classes) {
    std::lock_guard<std::mutex> lock(g_prof_mu);
    g_prof_on = false;
    for (int i = 0; i < max_classes; ++i) {
        if (ms_per_class) ms_per_class[i] = 0.f;
        if (launches_per_class) launches_per_class[i] = 0;
    }
    for (auto& r : g_prof) {
        PB_CUDA_TRY(cudaEventSynchronize(r.b));
        float ms = 0.f;
        PB_CUDA_TRY(cudaEventElapsedTime(&ms, r.a, r.b));
        if (r.cls < max_classes) {
            if (ms_per_class) ms_per_class[r.cls] += ms;
            if (launches_per_class) launches_per_class[r.cls] += 1;
        }
    }
    return PROF_NCLASSES;
}

const char* pb_profile_class_name(int i) { return (i >= 0 && i < PROF_NCLASSES) ? kProfNames[i] : ""; }

int pb_edgetaper_f32(const float* img, float* out, int B, int C, int H, int W, const float* kernel, int ksize,
                     int n_tapers, uint32_t flags, void* workspace, size_t workspace_bytes, void* stream_) {
    // img / out are the already padded images (H, W = padded sizes), like edgetaper.edgetaper
    cudaStream_t stream = (cudaStream_t)stream_;
    int rc;
    if ((rc = check_shape(B, C, H, W))) return rc;
    if (!img || !out || !kernel || img == out || ksize < 1 || ksize > PB_KS || !(ksize & 1) || n_tapers < 0) {
        set_error("bad arguments to pb_edgetaper_f32");
        return PB_ERR_ARG;
    }
    const size_t img_bytes = (size_t)B * C * H * W * sizeof(float);
    const size_t o_kern = 0;
    const size_t o_et = align_up((size_t)B * sizeof(ImgKernel), 256);
    const size_t o_tmp = o_et + align_up(edgetaper_scratch_bytes(B, H, W), 256);
    if ((rc = check_ws(workspace, workspace_bytes, o_tmp + img_bytes))) return rc;
    char* ws = static_cast<char*>(workspace);
    ImgKernel* kern = reinterpret_cast<ImgKernel*>(ws + o_kern);
    float* tmp = reinterpret_cast<float*>(ws + o_tmp);
    if ((rc = launch_params(nullptr, kern, nullptr, nullptr, nullptr, nullptr, kernel, nullptr, 2, B, ksize, 0.f, 0.f,
                            1e-8f, PB_ENGINE_SPATIAL, 1 << 30, nullptr, stream)))
        return rc;
    float *v = nullptr, *res = nullptr;
    if ((rc = launch_edgetaper_weights(kern, ws + o_et, B, H, W, (flags & PB_FLAG_EDGETAPER_BATCHMAX) ? 1 : 0, &v,
                                       stream)))
        return rc;
    // ping-pong so that the last pass lands in `out`
    PB_CUDA_TRY(cudaMemcpyAsync((n_tapers & 1) ? tmp : out, img, img_bytes, cudaMemcpyDeviceToDevice, stream));
    if (n_tapers == 0) return PB_OK;
    return launch_edgetaper_passes((n_tapers & 1) ? tmp : out, (n_tapers & 1) ? out : tmp, kern, v, B, C, H, W, n_tapers,
                                   &res, stream);
}

int pb_bilateral_f32(const float* img, float* out, int B, int C, int H, int W, float sigma_spatial,
                     float sigma_color, void* stream_) {
    int rc;
    if ((rc = check_shape(B, C, H, W))) return rc;
    if (!img || !out || img == out) {
        set_error("img/out must be distinct non-null device pointers");
        return PB_ERR_ARG;
    }
    return launch_bilateral(img, out, B * C, H, W, sigma_spatial, sigma_color, (cudaStream_t)stream_);
}

int pb_bilateral_vjp_f32(const float* img, const float* grad_out, float* grad_img, int B, int C, int H, int W,
                         float sigma_spatial, float sigma_color, void* stream_) {
    int rc;
    if ((rc = check_shape(B, C, H, W))) return rc;
    if (!img || !grad_out || !grad_img || grad_img == img || grad_img == grad_out) {
        set_error("img / grad_out / grad_img must be non-null device pointers, grad_img distinct from the inputs");
        return PB_ERR_ARG;
    }
    return launch_bilateral_vjp(img, grad_out, grad_img, B * C, H, W, sigma_spatial, sigma_color, (cudaStream_t)stream_);
}

int pb_recursive_filter_f32(const float* img, const float* joint, float* out, int B, int C, int H, int W,
                            float sigma_s, float sigma_r, int num_iterations, void* workspace,
                            size_t workspace_bytes, void* stream_) {
    int rc;
    if ((rc = check_shape(B, C, H, W))) return rc;
    if (!img || !out || img == out || num_iterations < 1) {
        set_error("bad arguments to pb_recursive_filter_f32");
        return PB_ERR_ARG;
    }
    if ((rc = check_ws(workspace, workspace_bytes, rf_workspace_bytes(B, H, W)))) return rc;
    return launch_recursive_filter(img, joint, out, B, C, H, W, (double)sigma_s, (double)sigma_r, num_iterations,
                                   workspace, (cudaStream_t)stream_);
}

int pb_normalized_convolution_f32(const float* img, float* out, int B, int C, int H, int W, float sigma_s,
                                  float sigma_r, int num_iterations, void* workspace, size_t workspace_bytes,
                                  void* stream_) {
    int rc;
    if ((rc = check_shape(B, C, H, W))) return rc;
    if (!img || !out || img == out || num_iterations < 1) {
        set_error("bad arguments to pb_normalized_convolution_f32");
        return PB_ERR_ARG;
    }
    if ((rc = check_ws(workspace, workspace_bytes, nc_workspace_bytes(B, C, H, W)))) return rc;
    return launch_normalized_convolution(img, out, B, C, H, W, (double)sigma_s, (double)sigma_r, num_iterations,
                                         workspace, (cudaStream_t)stream_);
}

static int check_patch_geometry(int B, int C, int h, int w, int ph, int pw, int step_h, int step_w, int ny, int nx,
                                int pad_top, int pad_left) {
    if (B < 1 || C < 1 || h < 1 || w < 1 || ph < 1 || pw < 1 || step_h < 1 || step_w < 1 || ny < 1 || nx < 1 ||
        pad_top < 0 || pad_left < 0 || step_h > ph || step_w > pw) {
        set_error("bad patch geometry");
        return PB_ERR_ARG;
    }
    // the patch grid must cover the padded image
    if ((ny - 1) * step_h + ph < h + pad_top || (nx - 1) * step_w + pw < w + pad_left) {
        set_error("patch grid %d x %d does not cover the padded %d x %d image", ny, nx, h, w);
        return PB_ERR_ARG;
    }
    return PB_OK;
}

int pb_patch_extract_f32(const float* img, size_t plane_stride, size_t row_stride, float* patches, int B, int C, int h,
                         int w, int ph, int pw, int step_h, int step_w, int ny, int nx, int pad_top, int pad_left,
                         void* stream_) {
    int rc;
    if ((rc = check_patch_geometry(B, C, h, w, ph, pw, step_h, step_w, ny, nx, pad_top, pad_left))) return rc;
    if (!img || !patches) {
        set_error("null pointer");
        return PB_ERR_ARG;
    }
    return launch_patch_extract(img, plane_stride, row_stride, patches, B, C, h, w, ph, pw, step_h, step_w, ny, nx, pad_top,
                                pad_left, (cudaStream_t)stream_);
}

int pb_patch_blend_f32(const float* patches, const float* win_y, const float* win_x, float* out, int B, int C, int h, int w,
                       int ph, int pw, int step_h, int step_w, int ny, int nx, int pad_top, int pad_left, void* stream_) {
    int rc;
    if ((rc = check_patch_geometry(B, C, h, w, ph, pw, step_h, step_w, ny, nx, pad_top, pad_left))) return rc;
    if (!patches || !win_y || !win_x || !out) {
        set_error("null pointer");
        return PB_ERR_ARG;
    }
    return launch_patch_blend(patches, win_y, win_x, out, B, C, h, w, ph, pw, step_h, step_w, ny, nx, pad_top, pad_left,
                              (cudaStream_t)stream_);
}

int pb_u8hwc_to_f32nchw(const uint8_t* in, float* out, int B, int H, int W, int C, void* stream_) {
    int rc;
    if ((rc = check_shape(B, C, H, W))) return rc;
    if (!in || !out) {
        set_error("null pointer");
        return PB_ERR_ARG;
    }
    return launch_u8_to_f32(in, out, B, H, W, C, (cudaStream_t)stream_);
}

int pb_f32nchw_to_u8hwc(const float* in, uint8_t* out, int B, int C, int H, int W, void* stream_) {
    int rc;
    if ((rc = check_shape(B, C, H, W))) return rc;
    if (!in || !out) {
        set_error("null pointer");
        return PB_ERR_ARG;
    }
    return launch_f32_to_u8(in, out, B, C, H, W, (cudaStream_t)stream_);
}

}  // extern "C"
