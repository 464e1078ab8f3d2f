"""Pin the CPU oracle against the round-2 vectors of the live reference (tests/golden/make_golden_r2.py):
kernels that are not point-symmetric and the compiled NC.cpp.  CPU-only."""
import os

import numpy as np
import pytest

from oracle import polyblur_oracle as po

G = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def r2():
    return np.load(os.path.join(G, "round2.npz"))


def maxabs(a, b):
    return float(np.max(np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64))))


@pytest.mark.parametrize("tag", ["a", "b", "c", "d"])
def test_normalized_convolution_matches_compiled_nc_cpp(r2, tag):
    """oracle.normalized_convolution restates NC.cpp:143-204; the fixtures are outputs of that file compiled
    where it lies.  (Sequential float32 running sums on both sides; 2e-6 is rounding of the box averages.)"""
    ss, sr, n = r2[f"nc/{tag}/par"]
    out = po.normalized_convolution(r2[f"nc/{tag}/in"], float(ss), float(sr), int(n))
    assert maxabs(out, r2[f"nc/{tag}/out"]) < 4e-6


@pytest.mark.parametrize("ks", [25, 9])
def test_asymmetric_kernels_convolve_like_the_reference(r2, ks):
    x, k = r2[f"asym/in{ks}"], r2[f"asym/k{ks}"]
    assert maxabs(k, k[..., ::-1, ::-1]) > 0.01                  # really not point-symmetric
    for ab, (alpha, beta) in {"a6b1": (6, 1), "a2b3": (2, 3)}.items():
        y = po.inverse_filtering_rank3(x, k, alpha=alpha, b=beta)
        assert maxabs(y, r2[f"asym/deconv{ks}/{ab}/conv"]) < 4e-6
        yc = po.inverse_filtering_rank3(x, np.ascontiguousarray(k[..., ::-1, ::-1]), alpha=alpha, b=beta)
        assert maxabs(yc, r2[f"asym/deconv{ks}/{ab}/corr"]) < 4e-6
    # the spatial restatement the CUDA engines mirror is the same convolution (small crop: O(625 HW))
    xs = x[:, :1, :24, :31]
    a = po.inverse_filtering_rank3(xs, k, alpha=6, b=1, dtype=np.float64)
    b = po.inverse_filtering_rank3(xs, k, alpha=6, b=1, dtype=np.float64, spatial=True)
    assert maxabs(a, b) < 1e-12
    y = po.inverse_filtering_rank3(x, k, alpha=6, b=1, do_edgetaper=True)
    assert maxabs(y, r2[f"asym/deconv{ks}/taper"]) < 4e-6
    assert maxabs(po.edgetaper(po.pad_with_kernel(x, ks // 2), k), r2[f"asym/edgetaper{ks}"]) < 3e-6


def test_asymmetric_kernel_gradients_match_autograd(r2):
    gi, gk, _ = po.inverse_filtering_rank3_vjp(r2["asym/in25"], r2["asym/k25"], r2["asym/vjp/w"], alpha=6, b=1)
    assert maxabs(gi, r2["asym/vjp/gx"]) < 5e-6 * np.abs(r2["asym/vjp/gx"]).max()
    assert maxabs(gk, r2["asym/vjp/gk"]) < 5e-6 * np.abs(r2["asym/vjp/gk"]).max()


def test_halo_masking_default_gradients(r2):
    """remove_halo with grad_img=None: gradients of the cropped (tapered) padded image (deblurring.py:200-203, 237-238)."""
    x, k = r2["asym/in25"], r2["halo/k"]
    for tag, taper in (("plain", False), ("taper", True)):
        y = po.inverse_filtering_rank3(x, k, alpha=6, b=1, remove_halo=True, do_edgetaper=taper, grad_img=None)
        assert maxabs(y, r2[f"halo/{tag}"]) < 4e-6
