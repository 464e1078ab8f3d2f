"""Pin the CPU oracle against vectors produced by the live reference
(tests/golden/make_golden.py).  CPU-only."""
import os

import numpy as np
import pytest

from oracle import polyblur_oracle as po

G = os.path.join(os.path.dirname(__file__), "golden")


def load(name):
    return np.load(os.path.join(G, name))


def maxabs(a, b):
    return float(np.max(np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64))))


def test_fourier_gradients_match_reference():
    st = load("stages.npz")
    for key in ("grad", "grad_odd"):
        gx, gy = po.fourier_gradients(st[key + "/in"])
        assert maxabs(gx, st[key + "/gx"]) < 3e-6
        assert maxabs(gy, st[key + "/gy"]) < 3e-6
        gx1, gy1 = po.fourier_gradients_1d(st[key + "/in"])
        assert maxabs(gx1, st[key + "/gx"]) < 3e-6
        assert maxabs(gy1, st[key + "/gy"]) < 3e-6


def test_kernels_match_reference():
    st = load("stages.npz")
    k = po.gaussian_kernel(st["kern/theta"], st["kern/sigma"], st["kern/rho"])
    assert k.shape == st["kern/k"].shape
    assert maxabs(k, st["kern/k"]) < 2e-7
    np.testing.assert_allclose(k.sum(axis=(-1, -2)), 1.0, atol=1e-6)


def test_direction_and_parameters_match_reference():
    st = load("stages.npz")
    m_n, m_o, theta, theta_deg, _ = po.find_direction(st["dir/mags"])
    np.testing.assert_allclose(m_n, st["dir/m_n"], rtol=1e-6)
    np.testing.assert_allclose(m_o, st["dir/m_o"], rtol=1e-6)
    np.testing.assert_allclose(theta, st["dir/theta"], rtol=1e-7)
    s, r = po.gaussian_parameters(m_n, m_o, 0.352, 0.768)
    np.testing.assert_allclose(s, st["dir/sigma"], rtol=2e-6)
    np.testing.assert_allclose(r, st["dir/rho"], rtol=2e-6)


@pytest.mark.parametrize("tag,alpha,beta", [("a6b1", 6, 1), ("a2b3", 2, 3)])
def test_deconvolution_matches_reference(tag, alpha, beta):
    st = load("stages.npz")
    out = po.inverse_filtering_rank3(st["deconv/in"], st["kern/k"], alpha=alpha, b=beta)
    assert maxabs(out, st["deconv/" + tag]) < 3e-6


def test_torus_restatement_equals_fft_path():
    st = load("stages.npz")
    x = st["deconv/in"][:3].astype(np.float64)
    k = st["kern/k"][:3].astype(np.float64)
    a = po.inverse_filtering_rank3(x, k, alpha=6, b=1, dtype=np.float64)
    b = po.inverse_filtering_rank3(x, k, alpha=6, b=1, dtype=np.float64, spatial=True)
    assert maxabs(a, b) < 1e-12


CASES = ["mosaic_rgb_48x64", "mosaic_gray_37x53", "white_rgb_40x41", "mosaic_c2_64x48",
         "tiny_rgb_8x8", "mosaic_rgb_96x120"]


@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("tag,n_iter,alpha,beta", [("a6b1n3", 3, 6, 1), ("a2b3n1", 1, 2, 3)])
def test_end_to_end_small_cases(name, tag, n_iter, alpha, beta):
    sc = load("small_cases.npz")
    tr = []
    out = po.polyblur_deblurring(sc[name + "/in"], n_iter=n_iter, alpha=alpha, beta=beta, trace=tr)
    th = np.stack([t["theta_deg"] for t in tr])
    sg = np.stack([t["sigma"] for t in tr])
    rh = np.stack([t["rho"] for t in tr])
    mg = np.stack([t["mags"] for t in tr])
    np.testing.assert_allclose(mg, sc[f"{name}/{tag}/mags"], rtol=2e-5, atol=2e-6)
    # theta only matters when the kernel is anisotropic
    aniso = np.abs(sc[f"{name}/{tag}/sigma"] - sc[f"{name}/{tag}/rho"]) > 1e-6
    assert np.array_equal(th[aniso], sc[f"{name}/{tag}/theta_deg"][aniso])
    np.testing.assert_allclose(sg, sc[f"{name}/{tag}/sigma"], rtol=5e-5)
    np.testing.assert_allclose(rh, sc[f"{name}/{tag}/rho"], rtol=5e-5)
    assert maxabs(out, sc[f"{name}/{tag}/out"]) < 1e-5


def test_peacock_kat(golden_dir):
    from PIL import Image
    kat = load("peacock_kat.npz")
    img = np.asarray(Image.open(os.path.join(golden_dir, "peacock_defocus.png"))).astype(np.float32) / 255
    tr = []
    out = po.polyblur_deblurring(img, n_iter=3, alpha=6, beta=1, trace=tr)
    assert out.shape == (500, 700, 3) and out.dtype == np.float32
    assert np.array_equal(np.stack([t["theta_deg"] for t in tr]), kat["theta_deg"])
    np.testing.assert_allclose(np.stack([t["sigma"] for t in tr]), kat["sigma"], rtol=2e-5)
    np.testing.assert_allclose(np.stack([t["rho"] for t in tr]), kat["rho"], rtol=2e-5)
    np.testing.assert_allclose(np.stack([t["mags"] for t in tr]), kat["mags"], rtol=1e-5)
    assert maxabs(out[::5, ::5], kat["out_sub"]) < 1e-5
    assert maxabs(out[200:264, 300:364], kat["out_crop"]) < 1e-5
    assert maxabs(out[:40, :40], kat["out_border"]) < 1e-5
    assert abs(out.astype(np.float64).mean() - float(kat["out_mean"])) < 1e-7
    # survey KAT
    assert abs(out.mean() - 0.3339409) < 1e-6


def test_options_match_reference():
    op = load("options.npz")
    x = op["in"]
    kw = dict(n_iter=2, alpha=6, beta=1)
    assert maxabs(po.polyblur_deblurring(x, **kw), op["default"]) < 1e-5
    assert maxabs(po.polyblur_deblurring(x, remove_halo=True, **kw), op["remove_halo"]) < 1e-5
    assert maxabs(po.polyblur_deblurring(x, edgetaping=True, **kw), op["edgetaping"]) < 1e-5
    assert maxabs(po.polyblur_deblurring(x, prefiltering=True, **kw), op["prefiltering"]) < 1e-5
    assert maxabs(po.polyblur_deblurring(x, discard_saturation=True, **kw), op["discard_saturation"]) < 1e-5
    assert maxabs(po.polyblur_deblurring(x[:1], q=1e-2, **kw), op["q1e-2_b1"]) < 1e-5
    assert maxabs(po.polyblur_deblurring(x[:1], remove_halo=True, edgetaping=True, prefiltering=True,
                                         discard_saturation=True, q=1e-2, **kw), op["all_b1"]) < 1e-5
    assert maxabs(po.polyblur_deblurring(x, remove_halo=True, edgetaping=True, prefiltering=True,
                                         discard_saturation=True, **kw), op["all_q0"]) < 1e-5
    # module surface defaults (b=0.468, beta=4, alpha=2)
    assert maxabs(po.polyblur_deblurring(x, n_iter=2, c=0.352, b=0.468, alpha=2, beta=4),
                  op["module_default"]) < 1e-5


def test_prefilters_match_reference():
    op = load("options.npz")
    x = op["in"]
    assert maxabs(po.bilateral_filter(x), op["bilateral"]) < 2e-6
    assert maxabs(po.recursive_filter(x, 2.0, 0.8, 1), op["rf_s2_r0.8_n1"]) < 2e-6
    assert maxabs(po.recursive_filter(x, 60, 0.4, 3), op["rf_s60_r0.4_n3"]) < 2e-6
    pad = po.pad_with_kernel(x, 12)
    assert maxabs(po.edgetaper(pad, op["edgetaper/k"]), op["edgetaper/out"]) < 3e-6


def test_ndarray_surface():
    nd = load("ndarray_api.npz")
    o = po.polyblur_deblurring(nd["hwc_in"], n_iter=2, alpha=6, beta=1)
    assert o.shape == nd["hwc_out"].shape and maxabs(o, nd["hwc_out"]) < 1e-5
    o = po.polyblur_deblurring(nd["hw_in"], n_iter=2, alpha=6, beta=1)
    assert o.shape == nd["hw_out"].shape and maxabs(o, nd["hw_out"]) < 1e-5


def test_fp64_truth_is_close_to_fp32():
    sc = load("small_cases.npz")
    x = sc["mosaic_rgb_96x120/in"]
    a = po.polyblur_deblurring(x, n_iter=3, alpha=6, beta=1, dtype=np.float64)
    assert maxabs(a, sc["mosaic_rgb_96x120/a6b1n3/out"]) < 1e-5


@pytest.mark.parametrize("name", CASES)
def test_torch_port_matches_reference(name):
    """The ATen-CPU port timed as the CPU baseline is pinned to the same golden vectors."""
    import torch
    from oracle import polyblur_oracle_torch as pt
    sc = load("small_cases.npz")
    x = torch.from_numpy(sc[name + "/in"])
    tr = []
    y = pt.polyblur_deblurring(x, n_iter=3, alpha=6, beta=1, trace=tr).numpy()
    assert maxabs(y, sc[f"{name}/a6b1n3/out"]) < 1e-6
    np.testing.assert_allclose(np.stack([t["sigma"].numpy() for t in tr]), sc[f"{name}/a6b1n3/sigma"], rtol=1e-6)
    y = pt.polyblur_deblurring(x, n_iter=1, alpha=2, beta=3).numpy()
    assert maxabs(y, sc[f"{name}/a2b3n1/out"]) < 1e-6


@pytest.mark.parametrize("tag,alpha,beta", [("a6b1", 6, 1), ("a2b4", 2, 4)])
def test_deconvolution_backward_oracle_matches_reference_autograd(tag, alpha, beta):
    """The numpy restatement of the backward pass (image and kernel gradients of inverse_filtering_rank3)
    against torch.autograd over the live reference (vjp.npz, tests/golden/make_golden_vjp.py)."""
    d = np.load(os.path.join(G, "vjp.npz"))
    gi, gk, pre = po.inverse_filtering_rank3_vjp(d["deconv_x"], d["deconv_kernels"], d["deconv_ybar"], alpha=alpha,
                                                 b=beta, dtype=np.float64)
    ref_i, ref_k = d[f"deconv_{tag}_grad"], d[f"deconv_{tag}_kernel_grad"]
    assert np.abs(np.clip(pre, 0, 1) - d[f"deconv_{tag}_y"]).max() < 5e-6
    assert np.abs(gi - ref_i).max() < 1e-5 * np.abs(ref_i).max()
    assert np.abs(gk - ref_k).max() < 1e-5 * np.abs(ref_k).max()
