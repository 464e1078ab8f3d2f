/* Plain-C consumer of the C ABI (no Python, no torch): compiles against include/polyblur_b200.h, loads the
 * shared library with dlopen and exercises every host-side entry point; the compute entry points must
 * refuse to run without a CUDA device (no CPU fallback) or with bad arguments.
 *     gcc -std=c99 -Iinclude tests/c/abi_smoke.c -ldl -o abi_smoke && ./abi_smoke polyblur_b200/libpolyblur_sm100.so */
#include <dlfcn.h>
#include <math.h>
#include <stdio.h>
#include <string.h>

#include "polyblur_b200.h"

#define LOAD(name)                                                       \
    *(void**)(&f_##name) = dlsym(h, #name);                              \
    if (!f_##name) {                                                     \
        fprintf(stderr, "missing symbol %s\n", #name);                   \
        return 2;                                                        \
    }

int main(int argc, char** argv) {
    if (argc < 2) return 64;
    void* h = dlopen(argv[1], RTLD_NOW | RTLD_LOCAL);
    if (!h) {
        fprintf(stderr, "dlopen: %s\n", dlerror());
        return 2;
    }
    int (*f_pb_version)(void);
    const char* (*f_pb_last_error)(void);
    void (*f_pb_default_params)(pb_params*);
    void (*f_pb_polynomial_coefficients)(double, double, float*);
    void (*f_pb_keys_weights)(float*);
    int (*f_pb_fft_plan)(int, int*);
    size_t (*f_pb_workspace_bytes)(int, int, int, int, const pb_params*);
    size_t (*f_pb_deconv_vjp_workspace_bytes)(int, int, int, int, int, int);
    size_t (*f_pb_backward_workspace_bytes)(int, int, int, int, int, int);
    int (*f_pb_polyblur_f32)(const float*, float*, int, int, int, int, const pb_params*, void*, size_t, float*, void*);
    LOAD(pb_version) LOAD(pb_last_error) LOAD(pb_default_params) LOAD(pb_polynomial_coefficients)
    LOAD(pb_keys_weights) LOAD(pb_fft_plan) LOAD(pb_workspace_bytes) LOAD(pb_deconv_vjp_workspace_bytes)
    LOAD(pb_backward_workspace_bytes) LOAD(pb_polyblur_f32)

    if (f_pb_version() < 1) return 3;
    pb_params p;
    f_pb_default_params(&p);
    if (p.n_iter != 1 || p.ker_size != 25 || fabs(p.c - 0.352) > 1e-12 || fabs(p.b - 0.768) > 1e-12) return 4;
    if (sizeof(pb_params) != 80) return 5;

    float a[4];
    f_pb_polynomial_coefficients(6.0, 1.0, a);          /* deblurring.py:160-162: a3 = 4, a2 = -9, a1 = 5, b = 1 */
    if (a[0] != 4.0f || a[1] != -9.0f || a[2] != 5.0f || a[3] != 1.0f) return 6;

    float w[210];
    f_pb_keys_weights(w);
    for (int i = 0; i < 30; ++i) {                       /* rows sum to sum / (sum + 1e-5) */
        float s = 0;
        for (int j = 0; j < 7; ++j) s += w[i * 7 + j];
        if (fabsf(s - 1.0f) > 2e-5f) return 7;
    }
    int radices[32];
    int ns = f_pb_fft_plan(1920, radices);
    long prod = 1;
    for (int i = 0; i < ns; ++i) prod *= radices[i];
    if (ns < 1 || prod != 1920) return 8;

    p.n_iter = 3;
    size_t ws = f_pb_workspace_bytes(32, 3, 1080, 1920, &p);
    if (ws < (size_t)32 * 1080 * 1920 * 4) return 9;     /* at least the gray plane */
    if (f_pb_workspace_bytes(0, 3, 1080, 1920, &p) != 0) return 10;
    if (f_pb_deconv_vjp_workspace_bytes(2, 3, 64, 64, 25, 0) == 0 || f_pb_backward_workspace_bytes(2, 3, 64, 64, 25, 0) == 0)
        return 11;
    if (f_pb_deconv_vjp_workspace_bytes(2, 3, 64, 64, 24, 0) != 0) return 12;   /* even kernel size */

    /* null pointers are rejected before anything touches the device */
    int rc = f_pb_polyblur_f32(NULL, NULL, 1, 3, 8, 8, &p, NULL, 0, NULL, NULL);
    if (rc >= 0 || strlen(f_pb_last_error()) == 0) return 13;
    printf("abi smoke ok: version %d, workspace for 32x3x1080x1920 n_iter=3: %zu bytes, plan(1920) = %d stages\n",
           f_pb_version(), ws, ns);
    dlclose(h);
    return 0;
}
