"""GPU parity tests added in round 2 (run with -m gpu): kernels that are not point-symmetric, the native
normalized convolution pinned to the compiled NC.cpp, the named BASELINE configurations at full size (4K at
n_iter = 3, the 12000 x 9000 image), every ``method=``, and the sharded path against the one-GPU path."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from oracle import polyblur_oracle as po

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
G = os.path.join(HERE, "golden")
ENGINE_AUTO, ENGINE_SPATIAL, ENGINE_FFT = 0, 1, 2


def maxabs(a, b):
    return float(np.max(np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64))))


def cu(x):
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


@pytest.fixture(scope="module")
def pb():
    import polyblur_b200
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return polyblur_b200


@pytest.fixture(scope="module")
def r2():
    return np.load(os.path.join(G, "round2.npz"))


# ---- kernels with K[-d] != K[d]: the reference's p2o / fft2 product is a convolution ----------------------------
@pytest.mark.parametrize("engine", [ENGINE_AUTO, ENGINE_SPATIAL, ENGINE_FFT])
@pytest.mark.parametrize("ks", [25, 9])
@pytest.mark.parametrize("ab,alpha,beta", [("a6b1", 6, 1), ("a2b3", 2, 3)])
def test_asymmetric_kernel_deconvolution_golden(pb, r2, engine, ks, ab, alpha, beta):
    """inverse_filtering_rank3 (deblurring.py:211-239) for a motion streak and a shifted Gaussian, every engine
    request (the FFT engine's real spectrum only serves point-symmetric kernels: these images are routed to
    the spatial engines on the device), with and without ``correlate`` (deblurring.py:229-230)."""
    x, k = cu(r2[f"asym/in{ks}"]), cu(r2[f"asym/k{ks}"])
    for corr in (False, True):
        got = pb.deblurring.inverse_filtering_rank3(x, k, alpha=alpha, b=beta, correlate=corr, engine=engine)
        assert maxabs(got.cpu().numpy(), r2[f"asym/deconv{ks}/{ab}/{'corr' if corr else 'conv'}"]) < 5e-6
    # the two differ by O(1): a correlation in place of the convolution cannot pass
    assert maxabs(r2[f"asym/deconv{ks}/{ab}/conv"], r2[f"asym/deconv{ks}/{ab}/corr"]) > 0.1


@pytest.mark.parametrize("ks", [25, 9])
def test_asymmetric_kernel_edgetaper_golden(pb, r2, ks):
    x, k = cu(r2[f"asym/in{ks}"]), cu(r2[f"asym/k{ks}"])
    xp = pb.utils.pad_with_kernel(x, k)
    got = pb.edgetaper.edgetaper(xp, k, n_tapers=3)
    assert maxabs(got.cpu().numpy(), r2[f"asym/edgetaper{ks}"]) < 3e-6
    got = pb.deblurring.inverse_filtering_rank3(x, k, alpha=6, b=1, do_edgetaper=True)
    assert maxabs(got.cpu().numpy(), r2[f"asym/deconv{ks}/taper"]) < 5e-6


@pytest.mark.parametrize("engine", [ENGINE_AUTO, ENGINE_SPATIAL])
def test_asymmetric_kernel_gradients_golden(pb, r2, engine):
    """d <w, deconv(x, k)> / dx and / dk from torch.autograd over the reference, asymmetric taps: the image
    gradient filters with the rotated kernel, the tap gradient is a correlation read at the kernel's offsets."""
    x = cu(r2["asym/in25"]).requires_grad_(True)
    k = cu(r2["asym/k25"]).requires_grad_(True)
    w = cu(r2["asym/vjp/w"])
    y = pb.deblurring.inverse_filtering_rank3(x, k, alpha=6, b=1, engine=engine)
    (y * w).sum().backward()
    gx, gk = r2["asym/vjp/gx"], r2["asym/vjp/gk"]
    assert maxabs(x.grad.cpu().numpy(), gx) < 2e-5 * np.abs(gx).max()
    assert maxabs(k.grad.cpu().numpy(), gk) < 1e-4 * np.abs(gk).max()
    # a kernel shared by the batch collects both images' terms
    k1 = cu(r2["asym/k25"][:1]).requires_grad_(True)
    y = pb.deblurring.inverse_filtering_rank3(cu(r2["asym/in25"]), k1, alpha=6, b=1, engine=engine)
    (y * w).sum().backward()
    _, gk1, _ = po.inverse_filtering_rank3_vjp(r2["asym/in25"], r2["asym/k25"][:1], r2["asym/vjp/w"], alpha=6, b=1)
    gk1 = gk1.sum(axis=0, keepdims=True)
    assert k1.grad.shape == (1, 1, 25, 25) and maxabs(k1.grad.cpu().numpy(), gk1) < 1e-4 * np.abs(gk1).max()


def test_stage_functions_refuse_to_drop_autograd_history(pb):
    x = torch.rand(1, 3, 24, 32, device="cuda", requires_grad=True)
    k = torch.from_numpy(po.gaussian_filter_np((1.5, 0.8), 0.3))[None, None].cuda()
    for call in (lambda: pb.edgetaper.edgetaper(x, k),
                 lambda: pb.domain_transform.recursive_filter(x), lambda: pb.blur_estimation.gaussian_blur_estimation(x)):
        with pytest.raises(NotImplementedError):
            call()
    # fourier_gradients is differentiable: D^T = -D
    gx, gy = pb.filters.fourier_gradients(x)
    a, b = torch.rand_like(gx), torch.rand_like(gy)
    ((gx * a).sum() + (gy * b).sum()).backward()
    ax, _ = pb.filters.fourier_gradients(a)
    _, by = pb.filters.fourier_gradients(b)
    assert maxabs(x.grad.cpu().numpy(), (-(ax + by)).cpu().numpy()) < 1e-6
    xr = x.detach().cpu().numpy().astype(np.float64)
    # adjoint identity against the float64 oracle: <D x, a> = <x, D^T a>
    ox, oy = po.fourier_gradients(xr, np.float64)
    lhs = float((ox * a.cpu().numpy()).sum() + (oy * b.cpu().numpy()).sum())
    rhs = float((xr * x.grad.cpu().numpy()).sum())
    assert abs(lhs - rhs) < 1e-3 * max(1.0, abs(lhs))


def test_halo_gradients_default_golden(pb, r2):
    x, k = cu(r2["asym/in25"]), cu(r2["halo/k"])
    for tag, taper in (("plain", False), ("taper", True)):
        for engine in (ENGINE_AUTO, ENGINE_SPATIAL, ENGINE_FFT):
            got = pb.deblurring.inverse_filtering_rank3(x, k, alpha=6, b=1, remove_halo=True, do_edgetaper=taper,
                                                        engine=engine)
            assert maxabs(got.cpu().numpy(), r2[f"halo/{tag}"]) < 5e-6


def test_halo_gradients_default_to_the_tapered_crop(pb):
    """remove_halo + do_edgetaper without grad_img: the reference takes the gradients of crop(tapered padded
    image) (deblurring.py:237-238, 200-203), which differs from the input in the border band."""
    from tests.test_gpu_parity import mosaic
    x = mosaic(2, 3, 90, 130, seed=5, sigma=(1.5, 0.8), theta_deg=20.0)
    k = po.gaussian_kernel(np.array([0.4, 1.9], np.float32), np.array([1.4, 2.2], np.float32),
                           np.array([0.7, 1.1], np.float32))
    ref = po.inverse_filtering_rank3(x, k, alpha=6, b=1, remove_halo=True, do_edgetaper=True, grad_img=None)
    got = pb.deblurring.inverse_filtering_rank3(cu(x), cu(k), alpha=6, b=1, remove_halo=True, do_edgetaper=True)
    assert maxabs(got.cpu().numpy(), ref) < 5e-6


# ---- native normalized convolution pinned to the compiled NC.cpp --------------------------------------------------
@pytest.mark.parametrize("tag", ["a", "b", "c", "d"])
def test_normalized_convolution_golden_from_nc_cpp(pb, r2, tag):
    """csrc/nc.cu against outputs of the reference's NC.cpp:143-204, JIT-compiled by tests/golden/make_golden_r2.py."""
    ss, sr, n = r2[f"nc/{tag}/par"]
    got = pb.domain_transform.normalized_convolution(cu(r2[f"nc/{tag}/in"]), float(ss), float(sr), int(n))
    assert maxabs(got.cpu().numpy(), r2[f"nc/{tag}/out"]) < 5e-6


# ---- every method= returns the 'fft' result -----------------------------------------------------
def test_every_method_gives_the_fft_result(pb):
    sc = np.load(os.path.join(G, "small_cases.npz"))
    x = cu(sc["mosaic_rgb_48x64/in"])
    ref = sc["mosaic_rgb_48x64/a6b1n3/out"]
    outs = {m: pb.polyblur_deblurring(x, n_iter=3, alpha=6, beta=1, method=m) for m in ("fft", "direct", "direct_separable")}
    assert maxabs(outs["fft"].cpu().numpy(), ref) < 1e-5
    assert torch.equal(outs["fft"], outs["direct"]) and torch.equal(outs["fft"], outs["direct_separable"])
    st = np.load(os.path.join(G, "stages.npz"))
    a = pb.deblurring.inverse_filtering_rank3(cu(st["deconv/in"]), cu(st["kern/k"]), alpha=6, b=1, method="direct")
    b = pb.deblurring.inverse_filtering_rank3(cu(st["deconv/in"]), cu(st["kern/k"]), alpha=6, b=1, method="fft")
    assert torch.equal(a, b) and maxabs(a.cpu().numpy(), st["deconv/a6b1"]) < 3e-6
    mod = pb.PolyblurDeblurring()
    assert torch.equal(mod(x, n_iter=3, alpha=6, beta=1, b=0.768, method="direct_separable"), outs["fft"])
    with pytest.raises(ValueError):
        pb.polyblur_deblurring(x, method="winograd")


# ---- BASELINE configurations at full size ---------------------------------------------------------------------------
def _check_against_torch_oracle(pb, x, n_iter, tol=1e-5):
    from oracle import polyblur_oracle_torch as pot
    ref = pot.polyblur_deblurring(torch.from_numpy(x), n_iter=n_iter, alpha=6, beta=1).numpy()
    out = pb.polyblur_deblurring(cu(x), n_iter=n_iter, alpha=6, beta=1).cpu().numpy()
    return maxabs(out, ref)


@pytest.mark.parametrize("kind", ["mosaic", "white"])
def test_4k_image_n_iter_3_vs_oracle(pb, kind):
    """One image of BASELINE configs 3 / 5 (3 x 2160 x 3840) at the metric's n_iter = 3, alpha = 6, beta = 1."""
    from polyblur_b200 import synthetic
    x = synthetic.make(kind, 1, 3, 2160, 3840).numpy()
    assert _check_against_torch_oracle(pb, x, 3) < 1e-5


@pytest.mark.parametrize("kind,n_iter", [("mosaic", 5), ("white", 5)])
def test_c4_12000x9000_vs_oracle(pb, kind, n_iter):
    """BASELINE config 4: a single 3 x 9000 x 12000 image, n_iter = 5, against the ATen-CPU restatement of the
    reference path (oracle/polyblur_oracle_torch.py, bit-identical to the reference on the golden inputs)."""
    from polyblur_b200 import synthetic
    x = synthetic.make(kind, 1, 3, 9000, 12000).numpy()
    assert _check_against_torch_oracle(pb, x, n_iter) < 1e-5


# ---- sharded run == one-GPU run, bit for bit ------------------------------------------------------------------------
def test_two_rank_gathered_output_is_bitwise_the_single_gpu_output():
    """torchrun --nproc-per-node 2: every rank deblurs its contiguous slice of the batch (sharding.shard_range),
    the slices are gathered (sharding.gather_outputs: NCCL when the box has two GPUs, gloo with both ranks on
    cuda:0 otherwise) and rank 0 compares with the whole batch processed by one call."""
    env = dict(os.environ)
    env["PYTHONPATH"] = os.path.dirname(HERE) + os.pathsep + env.get("PYTHONPATH", "")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29647", os.path.join(HERE, "mp_sharded_worker.py")]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "SHARDED_BITWISE_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


def test_chunk_images_bounds_the_workspace_and_keeps_the_result(pb):
    """pb_params.chunk_images: the batch runs in groups through the whole loop, in a workspace sized for one group."""
    import ctypes as C
    from polyblur_b200 import _lib, deblurring, synthetic
    x = synthetic.make("mosaic", 5, 3, 135, 240).cuda()
    whole, e0 = pb.polyblur_deblurring(x, n_iter=3, alpha=6, beta=1, return_estimates=True)
    grouped, e1 = pb.polyblur_deblurring(x, n_iter=3, alpha=6, beta=1, return_estimates=True, chunk_images=2)
    assert torch.equal(whole, grouped) and torch.equal(e0, e1)
    p_all = deblurring._make_params(3, 0.352, 0.768, 6, 1, 0.8, 2.0, 25, 0.0, False, False, False, False)
    p_grp = deblurring._make_params(3, 0.352, 0.768, 6, 1, 0.8, 2.0, 25, 0.0, False, False, False, False, chunk_images=2)
    n_all = _lib.lib().pb_workspace_bytes(5, 3, 135, 240, C.byref(p_all))
    n_grp = _lib.lib().pb_workspace_bytes(5, 3, 135, 240, C.byref(p_grp))
    assert n_grp < 0.6 * n_all


# ---- third-generation estimator kernels (csrc/estimate3.cu): edge shapes of the compile-time plans -------------
@pytest.mark.parametrize("shape,discard", [((2, 3, 1081, 1920), False),     # odd height: the last row pair has one row
                                           ((1, 3, 6, 1920), False),        # fewer row pairs than a CTA holds
                                           ((1, 3, 1080, 1090), False),     # width not a multiple of the 16-column tile
                                           ((1, 3, 1080, 34), True),        # narrow, with the saturation mask
                                           ((1, 3, 2160, 70), False),       # 4K column plan
                                           ((1, 3, 20, 3840), False)])      # 4K row plan
def test_estimator_gen3_edge_shapes(pb, shape, discard):
    """blur_estimation.py:18-65, 96-134 on shapes whose sides take the fused-stage kernels (1920 / 3840 wide rows,
    1080 / 2160 tall columns) with ragged other sides: directional maxima, direction index and sigma / rho
    against the CPU oracle."""
    rng = np.random.default_rng(7)
    B, C, H, W = shape
    x = rng.random(shape, dtype=np.float32)
    # a smooth component so that the estimate is not the clamped white-noise one, and some saturated pixels
    yy = np.linspace(0, 3, H, dtype=np.float32)[:, None]
    xx = np.linspace(0, 5, W, dtype=np.float32)[None, :]
    x = np.clip(0.25 * x + 0.5 + 0.45 * np.sin(yy * 2.1 + 0.3) * np.cos(xx * 1.7), 0, 1).astype(np.float32)
    tr = []
    po.gaussian_blur_estimation(x, c=0.352, b=0.768, q=0.0, discard_saturation=discard, trace=tr)
    e = pb.blur_estimation.estimate_parameters(cu(x), c=0.352, b=0.768, q=0.0, discard_saturation=discard)
    np.testing.assert_allclose(e["mags"].cpu().numpy(), tr[0]["mags"], rtol=3e-5, atol=3e-6)
    assert np.array_equal(e["theta_deg"].cpu().numpy().astype(np.int64), tr[0]["theta_deg"])
    np.testing.assert_allclose(e["sigma"].cpu().numpy(), tr[0]["sigma"], rtol=1e-4)
    np.testing.assert_allclose(e["rho"].cpu().numpy(), tr[0]["rho"], rtol=1e-4)
