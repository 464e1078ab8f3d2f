"""CPU-only checks of the C ABI: the library loads, exports every symbol the header
declares, and its host-only entry points agree with the oracle.  No compute calls."""
import ctypes
import os
import re

import numpy as np
import pytest

from oracle import polyblur_oracle as po

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as ge
    ge.build()
    from polyblur_b200 import _lib
    return _lib


def header_symbols():
    text = open(os.path.join(ROOT, "include", "polyblur_b200.h")).read()
    return sorted(set(re.findall(r"PB_API[^;(]*?\b(pb_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported(lib):
    syms = header_symbols()
    assert len(syms) >= 14
    handle = lib.lib()
    for s in syms:
        assert hasattr(handle, s), f"{s} declared in the header but not exported"
    assert sorted(lib.EXPORTS) == syms, "ctypes binding and header disagree"
    assert handle.pb_version() == 100


def test_params_struct_layout(lib):
    p = lib.default_params()
    assert (p.n_iter, p.c, p.b, p.alpha, p.beta, p.ker_size, p.q) == (1, 0.352, 0.768, 2.0, 3.0, 25, 0.0)
    assert (p.sigma_s, p.sigma_r, p.flags, p.engine) == (2.0, 0.8, 0, 0)
    assert ctypes.sizeof(lib.PbParams) == 80


def test_keys_weights_and_coefficients_match_oracle(lib):
    w = (ctypes.c_float * 210)()
    lib.lib().pb_keys_weights(w)
    ow, _ = po.keys_weights()
    assert np.array_equal(np.array(w, dtype=np.float32).reshape(30, 7), ow)
    c = (ctypes.c_float * 4)()
    for a, b in ((6, 1), (2, 3), (2, 4), (3.5, 0.25)):
        lib.lib().pb_polynomial_coefficients(a, b, c)
        np.testing.assert_allclose(list(c), po.polynomial_coefficients(a, b), rtol=1e-7)
        assert abs(sum(c) - 1.0) < 1e-6          # DC preserved


@pytest.mark.parametrize("n", [1, 2, 8, 25, 500, 700, 1080, 1104, 1920, 1944, 2160, 3840, 9000, 12000, 97, 9973])
def test_fft_plan_factors(lib, n):
    r = (ctypes.c_int * 32)()
    ns = lib.lib().pb_fft_plan(n, r)
    assert ns >= 0
    assert int(np.prod(list(r[:ns]), dtype=np.int64)) == n
    assert lib.lib().pb_fft_plan(0, r) < 0


def test_workspace_and_errors_without_gpu(lib):
    p = lib.default_params()
    p.n_iter = 3
    n = lib.lib().pb_workspace_bytes(32, 3, 1080, 1920, ctypes.byref(p))
    assert n >= 32 * 1080 * 1920 * 4 * (2 + 3)
    assert lib.lib().pb_workspace_bytes(0, 3, 8, 8, ctypes.byref(p)) == 0
    # argument validation happens before any CUDA call
    rc = lib.lib().pb_polyblur_f32(None, None, 1, 3, 8, 8, ctypes.byref(p), None, 0, None, None)
    assert rc == lib.PB_ERR_ARG and "in/out" in lib.last_error()
    # option validation: q outside [0, 0.5), both prefilters at once
    p.q = 0.7
    rc = lib.lib().pb_polyblur_f32(None, None, 1, 3, 8, 8, ctypes.byref(p), None, 0, None, None)
    assert rc == lib.PB_ERR_ARG and "q must be" in lib.last_error()
    p.q = 0.0
    p.flags = lib.FLAG_PREFILTER | lib.FLAG_PREFILTER_RF
    rc = lib.lib().pb_polyblur_f32(None, None, 1, 3, 8, 8, ctypes.byref(p), None, 0, None, None)
    assert rc == lib.PB_ERR_ARG
    # optional stages grow the workspace only when asked for
    p.flags = 0
    base = lib.lib().pb_workspace_bytes(4, 3, 120, 160, ctypes.byref(p))
    p.flags = lib.FLAG_REMOVE_HALO | lib.FLAG_EDGETAPER | lib.FLAG_PREFILTER_RF
    p.q = 1e-3
    assert lib.lib().pb_workspace_bytes(4, 3, 120, 160, ctypes.byref(p)) > base + 6 * 4 * 3 * 120 * 160 * 4


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import polyblur_b200
    from polyblur_b200._lib import PolyblurLibraryError
    with pytest.raises(PolyblurLibraryError):
        polyblur_b200.polyblur_deblurring(np.zeros((8, 8), np.float32), n_iter=1)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "polyblur_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.replace("# oracle", ""), f"{f} mentions the oracle"


def test_plain_c_consumer(tmp_path):
    """The boundary is usable from plain C: tests/c/abi_smoke.c compiles against include/polyblur_b200.h with
    gcc, dlopens the library and exercises the host-side entry points (no Python, no torch in the loop)."""
    import shutil
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    so = os.path.join(root, "polyblur_b200", "libpolyblur_sm100.so")
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    exe = str(tmp_path / "abi_smoke")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I" + os.path.join(root, "include"),
                    os.path.join(root, "tests", "c", "abi_smoke.c"), "-ldl", "-lm", "-o", exe], check=True)
    out = subprocess.run([exe, so], capture_output=True, text=True)
    assert out.returncode == 0, (out.returncode, out.stdout, out.stderr)
    assert "abi smoke ok" in out.stdout


def test_fft_plan_reports_the_register_radix_plan(lib):
    """1080p lengths take three stages of radices <= 16 (what the kernels run), not the fallback's plan."""
    r = (ctypes.c_int * 32)()
    for n, want in ((1920, 3), (1080, 3), (2016, 3), (1152, 3), (3840, 3), (2160, 3)):
        assert lib.lib().pb_fft_plan(n, r) == want
        assert max(r[:want]) <= 16
