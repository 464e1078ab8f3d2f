"""ctypes binding of libpolyblur_sm100.so (the C ABI in include/polyblur_b200.h).

There is no CPU fallback: if the shared library is missing, or no CUDA device is
available when a compute entry point is called, this module raises.  The library is
built in-tree by ``python __graft_entry__.py`` / ``make -C polyblur_b200/csrc``.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# PB_LIB_PATH selects another build of the same library (kernel tuning experiments only)
LIB_PATH = os.environ.get("PB_LIB_PATH") or os.path.join(_HERE, "libpolyblur_sm100.so")

PB_EST_STRIDE = 12
PB_OK, PB_ERR_ARG, PB_ERR_WORKSPACE, PB_ERR_CUDA, PB_ERR_UNSUPPORTED = 0, -1, -2, -3, -4
FLAG_REMOVE_HALO = 0x01
FLAG_EDGETAPER = 0x02
FLAG_PREFILTER = 0x04
FLAG_PREFILTER_RF = 0x08
FLAG_DISCARD_SATURATION = 0x10
FLAG_EDGETAPER_BATCHMAX = 0x20
FLAG_NO_CLAMP = 0x40
ENGINE_AUTO, ENGINE_SPATIAL, ENGINE_FFT = 0, 1, 2


class PbParams(C.Structure):
    """Mirror of ``struct pb_params``."""
    _fields_ = [
        ("c", C.c_double), ("b", C.c_double), ("alpha", C.c_double), ("beta", C.c_double),
        ("sigma_s", C.c_double), ("sigma_r", C.c_double), ("q", C.c_double),
        ("n_iter", C.c_int32), ("ker_size", C.c_int32), ("flags", C.c_uint32),
        ("engine", C.c_int32), ("tap_rel_threshold", C.c_float), ("chunk_images", C.c_int32),
    ]


class PolyblurLibraryError(RuntimeError):
    pass


_lib = None

_P = C.c_void_p
_SIGS = {
    "pb_version": (C.c_int, []),
    "pb_last_error": (C.c_char_p, []),
    "pb_default_params": (None, [C.POINTER(PbParams)]),
    "pb_polynomial_coefficients": (None, [C.c_double, C.c_double, C.POINTER(C.c_float)]),
    "pb_keys_weights": (None, [C.POINTER(C.c_float)]),
    "pb_fft_plan": (C.c_int, [C.c_int, C.POINTER(C.c_int)]),
    "pb_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(PbParams)]),
    "pb_profile_begin": (C.c_int, []),
    "pb_profile_end": (C.c_int, [C.POINTER(C.c_float), C.POINTER(C.c_int), C.c_int]),
    "pb_profile_class_name": (C.c_char_p, [C.c_int]),
    "pb_polyblur_f32": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(PbParams), _P,
                                  C.c_size_t, _P, _P]),
    "pb_fourier_gradients_f32": (C.c_int, [_P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, _P, C.c_size_t, _P]),
    "pb_estimate_f32": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double,
                                  C.c_uint32, _P, _P, C.c_size_t, _P]),
    "pb_make_kernel_f32": (C.c_int, [_P, _P, _P, C.c_int, C.c_int, _P, _P, C.c_size_t, _P]),
    "pb_deconv_f32": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, C.c_int, _P, C.c_int, C.c_double,
                                C.c_double, C.c_int, _P, C.c_size_t, _P]),
    "pb_deconv_ex_f32": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, C.c_int, _P, C.c_int, C.c_double,
                                   C.c_double, C.c_int, C.c_uint32, _P, _P, _P, C.c_size_t, _P]),
    "pb_deconv_vjp_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    "pb_deconv_vjp_f32": (C.c_int, [_P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, _P, C.c_int, C.c_double,
                                    C.c_double, C.c_int, _P, C.c_size_t, _P]),
    "pb_backward_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    "pb_estimate_trace_f32": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P, _P, C.c_size_t, _P]),
    "pb_estimate_trace_ex_f32": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint32, _P, _P, _P, C.c_size_t,
                                           _P]),
    "pb_kernel_grad_f32": (C.c_int, [_P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, _P, C.c_int, C.c_double,
                                     C.c_double, C.c_int, _P, _P, C.c_size_t, _P]),
    "pb_estimator_vjp_f32": (C.c_int, [_P, _P, _P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, _P, C.c_size_t, _P]),
    "pb_edgetaper_f32": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, C.c_int, _P, C.c_int, C.c_int,
                                   C.c_uint32, _P, C.c_size_t, _P]),
    "pb_bilateral_f32": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, _P]),
    "pb_bilateral_vjp_f32": (C.c_int, [_P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, _P]),
    "pb_recursive_filter_f32": (C.c_int, [_P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float,
                                          C.c_float, C.c_int, _P, C.c_size_t, _P]),
    "pb_u8hwc_to_f32nchw": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, C.c_int, _P]),
    "pb_f32nchw_to_u8hwc": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, C.c_int, _P]),
    "pb_patch_extract_f32": (C.c_int, [_P, C.c_size_t, C.c_size_t, _P] + [C.c_int] * 12 + [_P]),
    "pb_patch_blend_f32": (C.c_int, [_P, _P, _P, _P] + [C.c_int] * 12 + [_P]),
    "pb_normalized_convolution_f32": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float,
                                                C.c_float, C.c_int, _P, C.c_size_t, _P]),
}
EXPORTS = tuple(_SIGS)


def lib():
    """Load (once) and return the shared library; raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise PolyblurLibraryError(
                f"{LIB_PATH} is missing: build it with `python __graft_entry__.py` "
                "(nvcc, sm_100a).  polyblur_b200 has no CPU fallback.")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def last_error() -> str:
    return lib().pb_last_error().decode("utf-8", "replace")


def check(rc: int, what: str) -> None:
    if rc == PB_OK:
        return
    msg = f"{what} failed ({rc}): {last_error()}"
    if rc == PB_ERR_UNSUPPORTED:
        raise NotImplementedError(msg)
    if rc == PB_ERR_ARG:
        raise ValueError(msg)
    raise PolyblurLibraryError(msg)


def require_cuda(t: torch.Tensor | None = None) -> torch.device:
    """Device the work will run on; raises when there is no GPU (no CPU fallback)."""
    if not torch.cuda.is_available():
        raise PolyblurLibraryError("polyblur_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    if t is not None and t.is_cuda:
        return t.device
    return torch.device("cuda", torch.cuda.current_device())


def default_params() -> PbParams:
    p = PbParams()
    lib().pb_default_params(C.byref(p))
    return p


def workspace(B: int, Cn: int, H: int, W: int, p: PbParams, device: torch.device) -> torch.Tensor:
    n = lib().pb_workspace_bytes(B, Cn, H, W, C.byref(p))
    if n == 0:
        raise ValueError(f"bad shape {(B, Cn, H, W)}")
    # torch's caching allocator returns 512-byte aligned blocks and is stream ordered
    return torch.empty(n, dtype=torch.uint8, device=device)


def stream_ptr(device: torch.device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def ptr(t: torch.Tensor | None) -> int | None:
    return None if t is None else t.data_ptr()


def profile_begin() -> None:
    check(lib().pb_profile_begin(), "pb_profile_begin")


def profile_end() -> dict:
    """-> {kernel class: (total ms, launches)} for every launch since profile_begin()."""
    n = 16
    ms = (C.c_float * n)()
    cnt = (C.c_int * n)()
    k = lib().pb_profile_end(ms, cnt, n)
    if k < 0:
        check(k, "pb_profile_end")
    return {lib().pb_profile_class_name(i).decode(): (float(ms[i]), int(cnt[i])) for i in range(k) if cnt[i]}
