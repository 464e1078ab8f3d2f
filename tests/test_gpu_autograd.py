"""GPU tests of the backward pass: the vector-Jacobian product of the deconvolution
against torch.autograd over the live reference (tests/golden/vjp.npz, made by make_golden_vjp.py) and
against the adjoint identity; the Polyblur loop with the blur estimates held constant.  Run with -m gpu."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

G = os.path.join(os.path.dirname(__file__), "golden")
ENGINES = {"auto": 0, "spatial": 1, "fft": 2}


@pytest.fixture(scope="module")
def pb():
    import polyblur_b200
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return polyblur_b200


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(G, "vjp.npz"))


def cu(x):
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / np.abs(b).max())


@pytest.mark.parametrize("engine", sorted(ENGINES))
@pytest.mark.parametrize("tag,alpha,beta", [("a6b1", 6, 1), ("a2b4", 2, 4)])
def test_deconvolution_vjp_golden(pb, gold, tag, alpha, beta, engine):
    """inverse_filtering_rank3 with a requires_grad image: forward and gradient against the reference's
    autograd, a wide and a narrow kernel in one batch, part of the result clamped."""
    x = cu(gold["deconv_x"]).requires_grad_(True)
    k = cu(gold["deconv_kernels"])
    y = pb.deblurring.inverse_filtering_rank3(x, k, alpha=alpha, b=beta, engine=ENGINES[engine])
    assert y.requires_grad
    assert np.abs(y.detach().cpu().numpy() - gold[f"deconv_{tag}_y"]).max() < 1e-5
    (y * cu(gold["deconv_ybar"])).sum().backward()
    ref = gold[f"deconv_{tag}_grad"]
    assert rel(x.grad.cpu().numpy(), ref) < 2e-5
    # the clamp mask is part of the golden: a gradient that ignored it would be far off
    from polyblur_b200 import autograd as ag
    nomask = ag.inverse_filtering_rank3_vjp(cu(gold["deconv_ybar"]), k, alpha, beta, preclamp=None, engine=ENGINES[engine])
    assert rel(nomask.cpu().numpy(), ref) > 1e-2


@pytest.mark.parametrize("shape", [(1, 1, 5, 7), (2, 3, 33, 1), (1, 2, 1, 40), (2, 3, 97, 131), (1, 3, 300, 260)])
@pytest.mark.parametrize("engine", ["auto", "fft"])
def test_deconvolution_vjp_is_the_adjoint(pb, shape, engine):
    """<A x, y> = <x, A^T y> for the unclamped operator, odd / tiny / one-pixel-wide shapes, wide kernels."""
    from polyblur_b200 import autograd as ag
    from oracle import polyblur_oracle as po
    B, C, H, W = shape
    rng = np.random.default_rng(H * 1000 + W)
    x = cu(rng.standard_normal(shape).astype(np.float32))
    ybar = cu(rng.standard_normal(shape).astype(np.float32))
    ks = np.stack([po.gaussian_filter_np((float(rng.uniform(0.4, 3.0)), float(rng.uniform(0.3, 1.5))),
                                         float(rng.uniform(0, 3))) for _ in range(B)])[:, None].astype(np.float32)
    k = cu(ks)
    e = ENGINES[engine]
    Ax = ag._deconv_noclamp(x, k, 6, 1, e)
    ATy = ag.inverse_filtering_rank3_vjp(ybar, k, 6, 1, preclamp=None, engine=e)
    lhs = float((Ax.double() * ybar.double()).sum())
    rhs = float((x.double() * ATy.double()).sum())
    scale = float(Ax.double().norm() * ybar.double().norm())
    assert abs(lhs - rhs) < 1e-5 * scale


@pytest.mark.parametrize("engine", sorted(ENGINES))
@pytest.mark.parametrize("tag,alpha,beta", [("a6b1", 6, 1), ("a2b4", 2, 4)])
def test_kernel_gradient_golden(pb, gold, tag, alpha, beta, engine):
    """d <ybar, inverse_filtering_rank3(x, k)> / d k against autograd over the reference."""
    from polyblur_b200 import autograd as ag
    x, k, ybar = cu(gold["deconv_x"]), cu(gold["deconv_kernels"]), cu(gold["deconv_ybar"])
    e = ENGINES[engine]
    v = ag._deconv_noclamp(x, k, alpha, beta, e)
    kb = ag.kernel_grad(x, ybar, v, k, alpha, beta, e)
    assert rel(kb.cpu().numpy(), gold[f"deconv_{tag}_kernel_grad"]) < 1e-4


@pytest.mark.parametrize("shape,ks", [((2, 3, 45, 77), 25), ((1, 1, 64, 31), 9), ((2, 4, 19, 120), 15),
                                      ((1, 2, 130, 96), 25), ((1, 3, 7, 9), 9)])
@pytest.mark.parametrize("engine", ["auto", "fft"])
def test_deconvolution_backward_vs_oracle(pb, shape, ks, engine):
    """Image and kernel gradients of one deconvolution against the numpy backward oracle (float64) on odd,
    small and non-square shapes, 1-4 channels and smaller kernel supports; part of the result is clamped."""
    from oracle import polyblur_oracle as po
    from polyblur_b200 import autograd as ag
    B, C, H, W = shape
    rng = np.random.default_rng(B * 7 + H * 13 + W + ks)
    x = np.clip(rng.random(shape, dtype=np.float32) * 1.4 - 0.2, 0, 1).astype(np.float32)
    full = np.stack([po.gaussian_filter_np((float(rng.uniform(0.4, 2.2)), float(rng.uniform(0.3, 1.2))),
                                           float(rng.uniform(0, 3))) for _ in range(B)])[:, None]
    c0 = 12 - ks // 2
    k = full[..., c0:c0 + ks, c0:c0 + ks]
    k = (k / k.sum(axis=(-1, -2), keepdims=True)).astype(np.float32)
    ybar = rng.standard_normal(shape).astype(np.float32)
    _, _, pre = po.inverse_filtering_rank3_vjp(x, k, ybar, alpha=6, b=1, dtype=np.float64)
    # pixels whose unclamped value sits within rounding of 0 or 1 may fall on either side of the clamp mask:
    # give them no upstream gradient
    ybar[(np.abs(pre) < 5e-6) | (np.abs(pre - 1) < 5e-6)] = 0
    gi, gk, pre = po.inverse_filtering_rank3_vjp(x, k, ybar, alpha=6, b=1, dtype=np.float64)
    e = ENGINES[engine]
    xd, kd, yd = cu(x), cu(k), cu(ybar)
    v = ag._deconv_noclamp(xd, kd, 6, 1, e)
    assert np.abs(v.cpu().numpy() - pre).max() < 1e-5
    got_i = ag.inverse_filtering_rank3_vjp(yd, kd, 6, 1, preclamp=v, engine=e).cpu().numpy()
    got_k = ag.kernel_grad(xd, yd, v, kd, 6, 1, e).cpu().numpy()
    assert rel(got_i, gi) < 2e-5
    assert rel(got_k, gk) < 1e-4


def test_estimator_vjp_golden(pb, gold):
    """Gradient of <gaussian_blur_estimation(x), kbar> with respect to x: arg-max pixels, transposed spectral
    derivative, range normalisation with tied maxima, against autograd over the reference."""
    from polyblur_b200 import autograd as ag
    x, kbar = cu(gold["est_x"]), cu(gold["est_kbar"])
    k = pb.blur_estimation.gaussian_blur_estimation(x, q=0.0, c=0.352, b=0.768)
    assert np.abs(k.cpu().numpy() - gold["est_kernel"]).max() < 1e-6
    tf, tp = ag.estimate_trace(x)
    mbar = ag._maxima_grad(tf[:, :7], kbar, 0.352, 0.768, 25)
    gin = torch.zeros_like(x)
    ag.estimator_vjp(x, mbar, tf, tp, gin)
    ref = gold["est_grad"]
    assert float(np.abs(ref).max()) > 0
    assert rel(gin.cpu().numpy(), ref) < 2e-4


def test_polyblur_gradient(pb, gold):
    """Two iterations.  Default (estimate_grad=True): the reference's own gradient, estimator in the graph.
    estimate_grad=False: autograd over the reference with its estimator under no_grad.  Forward equals the
    no-grad path either way; the module surface routes the same way."""
    x0 = cu(gold["chain_x"])
    ybar = cu(gold["chain_ybar"])
    with torch.no_grad():
        y0 = pb.polyblur_deblurring(x0, n_iter=2, alpha=6, beta=1)
    for flag, key, tol in ((True, "chain_full_grad", 5e-4), (False, "chain_grad", 1e-4)):
        x = x0.clone().requires_grad_(True)
        y = pb.polyblur_deblurring(x, n_iter=2, alpha=6, beta=1, estimate_grad=flag)
        assert float((y.detach() - y0).abs().max()) < 2e-6
        assert np.abs(y.detach().cpu().numpy() - gold["chain_y"]).max() < 1e-5
        (y * ybar).sum().backward()
        assert rel(x.grad.cpu().numpy(), gold[key]) < tol, key
    x = x0.clone().requires_grad_(True)
    (pb.polyblur_deblurring(x, n_iter=1, alpha=6, beta=1) * ybar).sum().backward()
    assert rel(x.grad.cpu().numpy(), gold["chain_full_grad_1iter"]) < 5e-4
    xm = x0.clone().requires_grad_(True)
    ym = pb.PolyblurDeblurring()(xm, n_iter=1, c=0.352, b=0.768, alpha=6, beta=1)
    (ym * ybar).sum().backward()
    assert rel(xm.grad.cpu().numpy(), gold["chain_full_grad_1iter"]) < 5e-4
    # CPU tensors: the gradient comes back on the input's device
    xc = torch.from_numpy(gold["chain_x"]).requires_grad_(True)
    yc = pb.polyblur_deblurring(xc, n_iter=1, alpha=6, beta=1)
    yc.sum().backward()
    assert xc.grad is not None and xc.grad.device.type == "cpu" and bool(torch.isfinite(xc.grad).all())


def test_gradient_unsupported_options_raise(pb, gold):
    x = cu(gold["chain_x"]).requires_grad_(True)
    for kw in (dict(edgetaping=True), dict(q=0.01), dict(remove_halo=True, edgetaping=True)):
        with pytest.raises(NotImplementedError):
            pb.polyblur_deblurring(x, n_iter=1, **kw)
    with pytest.raises(NotImplementedError):
        pb.PolyblurDeblurring(patch_decomposition=True, patch_size=32)(x, n_iter=1)
    with torch.no_grad():
        pb.polyblur_deblurring(x, n_iter=1, remove_halo=True)


def test_halo_masking_gradients_golden(pb):
    """remove_halo=True under autograd (polyblur/deblurring.py:171-239): the stage with its own and with another
    image's gradients (image, kernel taps and that other image receive gradient), and two iterations of the loop
    with the estimator in the graph, against torch.autograd over the reference (tests/golden/vjp_halo.npz)."""
    gold = np.load(os.path.join(G, "vjp_halo.npz"))
    x, x0 = cu(gold["stage_x"]), cu(gold["stage_x0"])
    k, ybar = cu(gold["stage_kernels"]), cu(gold["stage_ybar"])
    # forward of the differentiable composite = the fused no-grad path = the reference
    with torch.no_grad():
        y_fused = pb.deblurring.inverse_filtering_rank3(x, k, alpha=6, b=1, remove_halo=True)
    xr, kr = x.clone().requires_grad_(True), k.clone().requires_grad_(True)
    y = pb.deblurring.inverse_filtering_rank3(xr, kr, alpha=6, b=1, remove_halo=True)
    assert float((y.detach() - y_fused).abs().max()) < 2e-6
    assert np.abs(y.detach().cpu().numpy() - gold["stage_self_y"]).max() < 1e-5
    (y * ybar).sum().backward()
    assert rel(xr.grad.cpu().numpy(), gold["stage_self_grad"]) < 5e-5
    assert rel(kr.grad.cpu().numpy(), gold["stage_self_kernel_grad"]) < 2e-4
    # the mask from another image's gradients: that image gets gradient too
    xr, x0r = x.clone().requires_grad_(True), x0.clone().requires_grad_(True)
    y = pb.deblurring.inverse_filtering_rank3(xr, k, alpha=6, b=1, remove_halo=True,
                                              grad_img=pb.filters.fourier_gradients(x0r))
    assert np.abs(y.detach().cpu().numpy() - gold["stage_other_y"]).max() < 1e-5
    (y * ybar).sum().backward()
    assert rel(xr.grad.cpu().numpy(), gold["stage_other_grad"]) < 5e-5
    assert rel(x0r.grad.cpu().numpy(), gold["stage_other_grad_x0"]) < 2e-4
    # the loop
    x, ybar = cu(gold["loop_x"]), cu(gold["loop_ybar"])
    with torch.no_grad():
        y_fused = pb.polyblur_deblurring(x, n_iter=2, alpha=6, beta=1, remove_halo=True)
    xr = x.clone().requires_grad_(True)
    y = pb.polyblur_deblurring(xr, n_iter=2, alpha=6, beta=1, remove_halo=True)
    assert float((y.detach() - y_fused).abs().max()) < 5e-6
    assert np.abs(y.detach().cpu().numpy() - gold["loop_y"]).max() < 1e-5
    (y * ybar).sum().backward()
    assert rel(xr.grad.cpu().numpy(), gold["loop_grad"]) < 5e-4
    # the module surface routes the same way
    xm = x.clone().requires_grad_(True)
    ym = pb.PolyblurDeblurring()(xm, n_iter=2, c=0.352, b=0.768, alpha=6, beta=1, remove_halo=True)
    (ym * ybar).sum().backward()
    assert rel(xm.grad.cpu().numpy(), gold["loop_grad"]) < 5e-4


def test_saturation_mask_gradients_golden(pb):
    """discard_saturation=True under autograd (blur_estimation.py:83-88, 112-119): the estimator's arg-max search leaves
    the saturated pixels out in the trace as in the forward pass; two iterations of the loop, and one with halo
    masking on top, against torch.autograd over the reference (tests/golden/vjp_halo.npz)."""
    gold = np.load(os.path.join(G, "vjp_halo.npz"))
    x, ybar = cu(gold["sat_x"]), cu(gold["loop_ybar"])
    with torch.no_grad():
        y_fused = pb.polyblur_deblurring(x, n_iter=2, alpha=6, beta=1, discard_saturation=True)
    xr = x.clone().requires_grad_(True)
    y = pb.polyblur_deblurring(xr, n_iter=2, alpha=6, beta=1, discard_saturation=True)
    assert float((y.detach() - y_fused).abs().max()) < 2e-6
    assert np.abs(y.detach().cpu().numpy() - gold["sat_y"]).max() < 1e-5
    (y * ybar).sum().backward()
    assert rel(xr.grad.cpu().numpy(), gold["sat_grad"]) < 5e-4
    # (the mask matters in this case: without it an arg-max pixel sits elsewhere and the gradient is another one)
    xn = x.clone().requires_grad_(True)
    (pb.polyblur_deblurring(xn, n_iter=2, alpha=6, beta=1) * ybar).sum().backward()
    assert rel(xn.grad.cpu().numpy(), gold["sat_grad"]) > 1e-2
    xr = x.clone().requires_grad_(True)
    y = pb.polyblur_deblurring(xr, n_iter=1, alpha=6, beta=1, discard_saturation=True, remove_halo=True)
    assert np.abs(y.detach().cpu().numpy() - gold["sat_halo_y"]).max() < 1e-5
    (y * ybar).sum().backward()
    assert rel(xr.grad.cpu().numpy(), gold["sat_halo_grad"]) < 5e-4


def test_bilateral_prefilter_gradients_golden(pb):
    """prefiltering=True under autograd (filters.py:107-148, deblurring.py:80-84, 99-110): the 5x5 bilateral filter's own
    backward kernel (pb_bilateral_vjp_f32: samples and range weights differentiated), two iterations of the loop, and one
    with halo masking on top.  The reference's filter scales an autograd-saved tensor in place and raises under
    autograd; the goldens are torch.autograd over the same arithmetic with that statement out of place
    (tests/golden/make_golden_vjp_halo.py checks the restatement bit-identical in the forward direction)."""
    gold = np.load(os.path.join(G, "vjp_halo.npz"))
    x, ybar = cu(gold["bil_x"]), cu(gold["loop_ybar"])
    xr = x.clone().requires_grad_(True)
    y = pb.filters.bilateral_filter(xr)
    assert np.abs(y.detach().cpu().numpy() - gold["bil_y"]).max() < 2e-6
    (y * ybar).sum().backward()
    assert rel(xr.grad.cpu().numpy(), gold["bil_grad"]) < 5e-5
    # adjoint identity on another shape (ragged sides, one channel): <J u, v> = <u, J^T v> to first order
    g = torch.Generator(device="cuda").manual_seed(5)
    a = torch.rand(1, 1, 37, 53, device="cuda", generator=g)
    u = torch.randn(a.shape, device="cuda", generator=g)
    v = torch.randn(a.shape, device="cuda", generator=g)
    ar = a.clone().requires_grad_(True)
    (pb.filters.bilateral_filter(ar) * v).sum().backward()
    eps = 1e-3
    with torch.no_grad():
        jv = (pb.filters.bilateral_filter(a + eps * u) - pb.filters.bilateral_filter(a - eps * u)) / (2 * eps)
    lhs, rhs = float((jv * v).sum()), float((ar.grad * u).sum())
    assert abs(lhs - rhs) < 2e-3 * max(1.0, abs(lhs)), (lhs, rhs)
    # the loop
    with torch.no_grad():
        y_fused = pb.polyblur_deblurring(x, n_iter=2, alpha=6, beta=1, prefiltering=True)
    xr = x.clone().requires_grad_(True)
    y = pb.polyblur_deblurring(xr, n_iter=2, alpha=6, beta=1, prefiltering=True)
    assert float((y.detach() - y_fused).abs().max()) < 5e-6
    assert np.abs(y.detach().cpu().numpy() - gold["pre_y"]).max() < 1e-5
    (y * ybar).sum().backward()
    assert rel(xr.grad.cpu().numpy(), gold["pre_grad"]) < 5e-4
    xr = x.clone().requires_grad_(True)
    y = pb.polyblur_deblurring(xr, n_iter=1, alpha=6, beta=1, prefiltering=True, remove_halo=True)
    assert np.abs(y.detach().cpu().numpy() - gold["pre_halo_y"]).max() < 1e-5
    (y * ybar).sum().backward()
    assert rel(xr.grad.cpu().numpy(), gold["pre_halo_grad"]) < 5e-4
    with pytest.raises(NotImplementedError):
        pb.polyblur_deblurring(x.clone().requires_grad_(True), n_iter=1, prefiltering=True, prefilter="rf")
    # CPU tensors go through the same composite: the gradient comes back on the input's device
    xc = torch.from_numpy(gold["bil_x"]).requires_grad_(True)
    yc = pb.polyblur_deblurring(xc, n_iter=1, alpha=6, beta=1, prefiltering=True, remove_halo=True)
    assert yc.device.type == "cpu" and np.abs(yc.detach().numpy() - gold["pre_halo_y"]).max() < 1e-5
    (yc * torch.from_numpy(gold["loop_ybar"])).sum().backward()
    assert xc.grad.device.type == "cpu" and rel(xc.grad.numpy(), gold["pre_halo_grad"]) < 5e-4
