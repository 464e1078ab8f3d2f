"""Host-side logic of the drop-in surfaces that needs no GPU: argument validation (raised before any
CUDA call), ndarray <-> tensor helpers, the patch-decomposition geometry, the synthetic inputs and the
CLI argument parser."""
import numpy as np
import pytest
import torch

import polyblur_b200 as pb
from oracle import polyblur_oracle as po


def test_argument_validation_happens_before_cuda():
    x = torch.rand(1, 3, 16, 16)
    with pytest.raises(ValueError):
        pb.polyblur_deblurring(x, method="bogus")
    with pytest.raises(ValueError):
        pb.polyblur_deblurring(x, n_angles=8)
    with pytest.raises(ValueError):
        pb.polyblur_deblurring(x, n_interpolated_angles=60)
    with pytest.raises(TypeError):
        pb.polyblur_deblurring(x.double())
    with pytest.raises(ValueError):
        pb.polyblur_deblurring(x[0])
    with pytest.raises(TypeError):
        pb.polyblur_deblurring([1, 2, 3])
    # n_iter = 0 returns the input object / the squeezed array without touching the GPU (deblurring.py:60,96)
    assert pb.polyblur_deblurring(x, n_iter=0) is x
    arr = np.random.default_rng(0).random((5, 7, 1), dtype=np.float32)
    assert pb.polyblur_deblurring(arr, n_iter=0).shape == (5, 7)
    with pytest.raises(TypeError):
        pb.io.deblur_uint8(np.zeros((4, 4, 3), np.float32))


def test_utils_match_oracle_restatement():
    rng = np.random.default_rng(1)
    a = rng.random((6, 9, 3), dtype=np.float32)
    t = pb.utils.to_tensor(a)
    assert t.shape == (3, 6, 9) and np.array_equal(t.numpy(), po.to_tensor(a))
    assert np.array_equal(pb.utils.to_array(t[None]), po.to_array(t[None].numpy()))
    g = rng.random((6, 9), dtype=np.float32)
    assert pb.utils.to_tensor(g).shape == (1, 6, 9) and pb.utils.to_array(pb.utils.to_tensor(g)).shape == (6, 9)
    x = torch.from_numpy(rng.random((2, 3, 8, 10), dtype=np.float32))
    k = torch.zeros(1, 1, 25, 25)
    p = pb.utils.pad_with_kernel(x, k)
    assert p.shape == (2, 3, 32, 34) and np.array_equal(p.numpy(), po.pad_with_kernel(x.numpy(), 12))
    assert torch.equal(pb.utils.crop_with_kernel(p, k), x)
    u8 = (rng.random((4, 5, 3)) * 255).astype(np.uint8)
    assert np.array_equal(pb.utils.to_ubyte(pb.utils.to_float(u8)), u8)         # main.py:146 img_as_ubyte rounds
    f = np.array([0.0, 0.999, 1.0, 0.5, 1.7, -0.2], np.float32)
    assert pb.utils.to_uint(f).tolist() == [0, 254, 255, 127, 255, 0]               # utils.py:41-45 truncates
    assert pb.utils.to_ubyte(f).tolist() == [0, 255, 255, 128, 255, 0]


def test_patch_geometry_helpers():
    mod = pb.PolyblurDeblurring(patch_decomposition=True, patch_size=64, patch_overlap=0.25, batch_size=3)
    assert len(list(mod.parameters())) == 0 and len(list(mod.buffers())) == 0
    x = torch.arange(2 * 3 * 150 * 200, dtype=torch.float32).reshape(2, 3, 150, 200)
    step = int(64 * 0.75)
    new_h = int(np.ceil((150 - 64) / step) * step) + 64
    new_w = int(np.ceil((200 - 64) / step) * step) + 64
    padded = mod.pad_with_new_size(x, (new_h, new_w), mode="replicate")
    assert padded.shape[-2:] == (new_h, new_w)
    assert torch.equal(mod.crop_with_old_size(padded, (150, 200)), x)
    w = mod.build_window((64, 64), "kaiser")
    assert w.shape == (64, 64) and float(w.max()) <= 1.0 and float(w.min()) > 0.0
    with pytest.raises(ValueError):
        mod.build_window((8, 8), "nope")


def test_filters_gaussian_filter_matches_oracle():
    for sig, th in (((2.5, 1.2), 0.5236), ((0.4, 0.4), 0.0), ((4.0, 0.3), 2.0)):
        k = pb.filters.gaussian_filter(sig, th, k_size=np.array([25, 25]))
        assert k.shape == (25, 25) and abs(float(k.sum()) - 1.0) < 1e-5
        assert np.allclose(k, po.gaussian_filter_np(sig, th), atol=1e-7)


def test_cli_parser_defaults_match_reference():
    from polyblur_b200 import main as cli
    a = cli.build_parser().parse_args(["--impath", "x.png"])
    # main.py:30-55
    assert (a.N, a.alpha, a.beta, a.q) == (3, 2, 3, 0)
    assert (a.sigma, a.rho, a.theta, a.sigma_n) == (3.0, 1.0, 0.0, 0.01)
    assert (a.patch_size, a.patch_overlap) == (400, 0.25)
    assert not (a.synthetic_degradation or a.do_prefiltering or a.do_halo_removal or a.do_edgetaping
                or a.do_patch_decomposition)
    assert cli.str2bool("Yes") is True and cli.str2bool("0") is False
    img = np.random.default_rng(2).random((40, 50, 3)).astype(np.float32)
    blurred = cli.synthetic_blur(img, 2.0, 1.0, 30.0, 0.0)
    assert blurred.shape == img.shape and abs(float(blurred.mean()) - float(img.mean())) < 1e-3


def test_pipeline_chunks_cover_the_batch():
    """The host pipelines' chunk schedule: every image exactly once, small chunks at both ends."""
    from polyblur_b200.sharding import pipeline_chunks
    for B in range(1, 70):
        for body in (1, 2, 3, 8):
            for ramp in ((1,), (1, 3), ()):
                s = pipeline_chunks(B, body, ramp)
                assert sum(s) == B and all(v > 0 for v in s)
                assert max(s) <= body
    assert pipeline_chunks(32, 2, (1,)) == [1] + [2] * 15 + [1]
    assert pipeline_chunks(32, 8, (1, 3)) == [1, 3, 8, 8, 8, 3, 1]


def test_maxima_gradient_chain_matches_reference_autograd():
    """The host-side scalar chain of the backward pass (autograd._maxima_grad: Keys interpolation, arg-min
    direction, affine model with clamping, Gaussian taps; blur_estimation.py:138-232) against torch.autograd
    over the reference's own functions (vjp.npz chainrule_*), on CPU tensors."""
    import os
    import numpy as np
    import torch
    from polyblur_b200 import autograd as ag
    d = np.load(os.path.join(os.path.dirname(__file__), "golden", "vjp.npz"))
    m, kb, ref = (torch.from_numpy(d[k]) for k in ("chainrule_m", "chainrule_kbar", "chainrule_mbar"))
    got = ag._maxima_grad(m, kb, 0.352, 0.768, 25)
    assert float(ref.abs().max()) > 0
    assert float((got - ref).abs().max()) < 1e-5 * float(ref.abs().max())
    sr = d["chainrule_sigma_rho"]
    # image 2 sits on the upper clamp of the affine model with both sigma and rho (4.0): no gradient at all;
    # image 1 has rho on the lower clamp (0.3) and still gets the sigma term
    assert np.allclose(sr[2], 4.0) and float(got[2].abs().max()) == 0.0
    assert np.isclose(sr[1, 1], 0.3) and float(got[1].abs().max()) > 0.0


def test_import_name_drop_in():
    """polyblur_b200.compat.install(): `import polyblur` and the reference's module names resolve to this package
    (polyblur/__init__.py:1), and a foreign `polyblur` already imported is not silently replaced."""
    import importlib
    import sys
    import types

    import polyblur_b200
    from polyblur_b200 import compat

    saved = {k: v for k, v in sys.modules.items() if k == "polyblur" or k.startswith("polyblur.")}
    for k in saved:
        del sys.modules[k]
    try:
        compat.install()
        import polyblur
        from polyblur import PolyblurDeblurring, polyblur_deblurring
        from polyblur.deblurring import inverse_filtering_rank3
        from polyblur.blur_estimation import gaussian_blur_estimation
        from polyblur.filters import fourier_gradients
        from polyblur.edgetaper import edgetaper
        from polyblur.domain_transform import recursive_filter
        from polyblur.utils import to_tensor, to_array
        assert polyblur is polyblur_b200
        assert polyblur_deblurring is polyblur_b200.polyblur_deblurring
        assert PolyblurDeblurring is polyblur_b200.PolyblurDeblurring
        assert inverse_filtering_rank3 is polyblur_b200.deblurring.inverse_filtering_rank3
        assert importlib.import_module("polyblur.filters") is polyblur_b200.filters
        for f in (gaussian_blur_estimation, fourier_gradients, edgetaper, recursive_filter, to_tensor, to_array):
            assert callable(f)
        compat.uninstall()
        assert "polyblur" not in sys.modules and "polyblur.deblurring" not in sys.modules
        sys.modules["polyblur"] = types.ModuleType("polyblur")          # somebody else's package of that name
        with pytest.raises(ImportError):
            compat.install()
        compat.install(force=True)
        assert sys.modules["polyblur"] is polyblur_b200
        compat.uninstall()
    finally:
        for k in [k for k in sys.modules if k == "polyblur" or k.startswith("polyblur.")]:
            del sys.modules[k]
        sys.modules.update(saved)
