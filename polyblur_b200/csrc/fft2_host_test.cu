// CPU unit test of the fft2.cuh core: the __host__ __device__ stage functions run with a single
// simulated thread.  Prints one line per length:  n  stages  err_dif  err_roundtrip  perm_ok
// (errors relative to the spectrum's max magnitude, against a float64 naive DFT).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "fft2_static.cuh"

using namespace pb;

struct ArraySrc {
    const float2* a;
    int stride;
    float2 operator()(int f, int i) const { return a[f * stride + i]; }
};
struct ArrayDst {
    float2* a;
    int stride;
    void operator()(int f, int i, float2 v) const { a[f * stride + i] = v; }
};

// compile-time plans against the run-time core on the same data: must agree to rounding (same
// operations, possibly different contraction)
template <class P>
int check_static() {
    Fft2Plan plan;
    if (make_fft2_plan(P::n, &plan) != 0 || !P::matches(plan)) {
        printf("static %d: radix order differs from the planner\n", P::n);
        return 1;
    }
    const int n = P::n, nb = 2, stride = n + 1;
    std::vector<float2> tw(plan.tw_total);
    for (int k = 0; k < plan.tw_total; ++k) tw[k] = fft2_stage_twiddle(k, plan);
    std::vector<float2> a(nb * stride), b;
    srand(n + 7);
    for (auto& v : a) v = make_float2(rand() / (float)RAND_MAX - 0.5f, rand() / (float)RAND_MAX - 0.5f);
    b = a;
    std::vector<float> h(stride * nb + n);
    for (size_t i = 0; i < h.size(); ++i) h[i] = 0.5f + (float)((i * 2654435761u) % 1000) / 1000.0f;
    double worst = 0;
    {
        std::vector<float2> u = a, v = a;
        fft2_forward_dif(u.data(), stride, nb, plan, tw.data(), 0, 1);
        s_forward_dif<P>(v.data(), stride, nb, tw.data(), 0, 1);
        for (size_t i = 0; i < u.size(); ++i) worst = fmax(worst, fmax(fabs(u[i].x - v[i].x), fabs(u[i].y - v[i].y)));
        fft2_forward_dit(u.data(), stride, nb, plan, tw.data(), 0, 1);
        s_forward_dit<P>(v.data(), stride, nb, tw.data(), 0, 1);
        for (size_t i = 0; i < u.size(); ++i) worst = fmax(worst, fmax(fabs(u[i].x - v[i].x), fabs(u[i].y - v[i].y)) / n);
    }
    {
        std::vector<float2> u = a, v = a;
        fft2_forward_mul_inverse(u.data(), stride, nb, plan, tw.data(), 0, 1, h.data(), 1);
        s_forward_mul_inverse<P, 1>(v.data(), stride, nb, tw.data(), 0, 1, h.data());
        for (size_t i = 0; i < u.size(); ++i) worst = fmax(worst, fmax(fabs(u[i].x - v[i].x), fabs(u[i].y - v[i].y)) / n);
        u = a; v = a;
        fft2_forward_mul_inverse(u.data(), stride, nb, plan, tw.data(), 0, 1, h.data(), 2);
        s_forward_mul_inverse<P, 2>(v.data(), stride, nb, tw.data(), 0, 1, h.data());
        for (size_t i = 0; i < u.size(); ++i) worst = fmax(worst, fmax(fabs(u[i].x - v[i].x), fabs(u[i].y - v[i].y)) / n);
    }
    printf("static %d %.3e\n", n, worst);
    return worst > 1e-5 ? 1 : 0;
}

int main(int argc, char** argv) {
    int worst = 0;
    if (argc > 1 && std::string(argv[1]) == "static") {
        worst |= check_static<PlanW1920>() | check_static<PlanH1080>() | check_static<PlanX2016>() |
                 check_static<PlanY1152>() | check_static<PlanW3840>() | check_static<PlanH2160>() |
                 check_static<PlanX4000>() | check_static<PlanY2304>();
        return worst;
    }
    for (int a = 1; a < argc; ++a) {
        const int n = atoi(argv[a]);
        Fft2Plan plan;
        if (make_fft2_plan(n, &plan) != 0) {
            printf("%d noplan\n", n);
            continue;
        }
        long long prod = 1;
        for (int s = 0; s < plan.ns; ++s) prod *= plan.radix[s];
        std::vector<float2> tw(plan.tw_total);
        for (int k = 0; k < plan.tw_total; ++k) tw[k] = fft2_stage_twiddle(k, plan);
        const int nb = 2;
        const int stride = n + 3;
        std::vector<float2> x(nb * stride), x0;
        srand(n);
        for (auto& v : x) v = make_float2(rand() / (float)RAND_MAX - 0.5f, rand() / (float)RAND_MAX - 0.5f);
        x0 = x;
        fft2_forward_dif(x.data(), stride, nb, plan, tw.data(), 0, 1);
        // naive DFT in double for sequence 1 (subsampled for long n)
        double err = 0, mag = 0;
        const int step = n > 4096 ? 37 : 1;
        bool perm_ok = (prod == n);
        for (int p = 0; p < n; p += step) {
            const int k = fft2_freq_of_slot(p, plan);
            if (fft2_slot_of_freq(k, plan) != p) perm_ok = false;
            double re = 0, im = 0;
            for (int m = 0; m < n; ++m) {
                double ang = -2.0 * M_PI * (double)(((long long)m * k) % n) / (double)n;
                re += x0[stride + m].x * cos(ang) - x0[stride + m].y * sin(ang);
                im += x0[stride + m].x * sin(ang) + x0[stride + m].y * cos(ang);
            }
            err = fmax(err, fmax(fabs(re - x[stride + p].x), fabs(im - x[stride + p].y)));
            mag = fmax(mag, hypot(re, im));
        }
        // inverse through the swap trick
        for (auto& v : x) v = make_float2(v.y, v.x);
        fft2_forward_dit(x.data(), stride, nb, plan, tw.data(), 0, 1);
        double rt = 0;
        for (int i = 0; i < nb * stride; ++i) {
            if (i % stride >= n) continue;
            rt = fmax(rt, fabs(x[i].y / n - x0[i].x));
            rt = fmax(rt, fabs(x[i].x / n - x0[i].y));
        }
        // forward -> multiply -> inverse with the fused middle stage against the unfused pair
        {
            std::vector<float> h(stride * nb + n);
            for (size_t i = 0; i < h.size(); ++i) h[i] = 0.5f + (float)((i * 2654435761u) % 1000) / 1000.0f;
            for (int mode = 1; mode <= 2; ++mode) {
                std::vector<float2> a = x0, b = x0;
                // sequences are addressed with `stride`, the per-sequence table of mode 2 with n
                fft2_forward_dif(a.data(), stride, nb, plan, tw.data(), 0, 1);
                fft2_forward_dit(a.data(), stride, nb, plan, tw.data(), 0, 1, h.data(), mode);
                fft2_forward_mul_inverse(b.data(), stride, nb, plan, tw.data(), 0, 1, h.data(), mode);
                for (int i = 0; i < nb * stride; ++i) {
                    if (i % stride >= n) continue;
                    rt = fmax(rt, fabs(a[i].x - b[i].x) / n);
                    rt = fmax(rt, fabs(a[i].y - b[i].y) / n);
                }
            }
        }
        // the same round trip through the fused first / last stages (source / sink functors)
        if (plan.ns >= 2) {
            std::vector<float2> work(nb * stride), res(nb * stride), swapped(nb * stride);
            ArraySrc src{x0.data(), stride};
            fft2_forward_dif_from(work.data(), stride, nb, plan, tw.data(), 0, 1, src);
            for (auto& v : work) v = make_float2(v.y, v.x);
            ArrayDst dst{res.data(), stride};
            fft2_forward_dit_to(work.data(), stride, nb, plan, tw.data(), 0, 1, dst);
            for (int i = 0; i < nb * stride; ++i) {
                if (i % stride >= n) continue;
                rt = fmax(rt, fabs(res[i].y / n - x0[i].x));
                rt = fmax(rt, fabs(res[i].x / n - x0[i].y));
            }
        }
        printf("%d %d %.3e %.3e %d", n, plan.ns, err / (mag > 0 ? mag : 1), rt, perm_ok ? 1 : 0);
        for (int s = 0; s < plan.ns; ++s) printf(" r%d", plan.radix[s]);
        printf("\n");
        if (!perm_ok || err / (mag > 0 ? mag : 1) > 2e-6 || rt > 2e-6) worst = 1;
    }
    return worst;
}
