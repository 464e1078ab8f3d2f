"""Generate the golden fixtures in this directory from the LIVE reference.

Run in the build container only (it needs /root/reference, which does not exist on
the GPU box):

    python tests/golden/make_golden.py

The reference (teboli/polyblur, commit 86ca0d0) is imported unmodified; the only shim
is a 3-line stub for ``skimage`` (absent here, used only by to_float/to_uint).  All
outputs are float32, method='fft' (the reference's default path).  The resulting
``*.npz`` files are committed; tests never import the reference.
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("POLYBLUR_REFERENCE", "/root/reference")

sk = types.ModuleType("skimage")
sk.img_as_float32 = lambda x: x
sys.modules["skimage"] = sk
sys.path.insert(0, REF)

import polyblur  # noqa: E402  (the reference)
from polyblur import blur_estimation, deblurring, domain_transform, edgetaper, filters, utils  # noqa: E402

torch.set_num_threads(1)  # batch-size / thread-count invariance is not bitwise in ATen; pin it


def mosaic(B, C, H, W, seed, sigma=(2.5, 1.2), theta_deg=30.0, block=12):
    """Blurred random-block mosaic (SURVEY.md 8d distribution M, smaller blocks)."""
    g = torch.Generator().manual_seed(seed)
    small = torch.rand(B, C, -(-H // block), -(-W // block), generator=g)
    img = small.repeat_interleave(block, -2).repeat_interleave(block, -1)[..., :H, :W].contiguous()
    k = filters.gaussian_filter(sigma, theta_deg * np.pi / 180, k_size=np.array([25, 25]))
    k = torch.from_numpy(k)[None, None].repeat(B, 1, 1, 1)
    return filters.convolve2d(img, k, method="fft").clamp(0, 1).contiguous()


def traced_run(x, n_iter, c, b, alpha, beta, q=0.0, discard_saturation=False):
    """Stage-by-stage replay of polyblur_deblurring (deblurring.py:68-88) with taps."""
    thetas = torch.linspace(0, 180, 7).unsqueeze(0).long()
    interp = torch.arange(0, 180, 6).unsqueeze(0).long()
    tr = dict(mags=[], interp=[], theta_deg=[], sigma=[], rho=[], kernel=[], out=[])
    cur = x
    for _ in range(n_iter):
        gray = cur.mean(dim=1, keepdims=True)
        mask = blur_estimation.get_saturation_mask(gray, discard_saturation)
        gn = blur_estimation.normalize(gray, q=q)
        grads = blur_estimation.compute_gradients(gn, mask=mask)
        mags = blur_estimation.compute_gradient_magnitudes(grads)
        mi = blur_estimation.cubic_interpolator(interp / 30, thetas / 30, mags)
        m_n, m_o, th = blur_estimation.find_maximal_blur_direction(mags, thetas, interp)
        sigma, rho = blur_estimation.compute_gaussian_parameters(m_n, m_o, c=c, b=b)
        kernel = blur_estimation.create_gaussian_filter(th, sigma, rho, ksize=25)
        cur = deblurring.inverse_filtering_rank3(cur, kernel, alpha=alpha, b=beta, method="fft")
        cur = cur.clip(0.0, 1.0)
        tr["mags"].append(mags.numpy())
        tr["interp"].append(mi.numpy())
        tr["theta_deg"].append(np.rint(th.numpy() * 180 / np.pi).astype(np.int64)[:, 0])
        tr["sigma"].append(sigma.numpy()[:, 0])
        tr["rho"].append(rho.numpy()[:, 0])
        tr["kernel"].append(kernel.numpy())
        tr["out"].append(cur.numpy())
    return cur, {k: np.stack(v) for k, v in tr.items()}


def main():
    from PIL import Image

    # ---- 1. peacock known-answer test (config 1) -------------------------------------
    png = os.path.join(REF, "pictures", "peacock_defocus.png")
    img = np.asarray(Image.open(png)).astype(np.float32) / 255
    out = polyblur.polyblur_deblurring(img, n_iter=3, alpha=6, beta=1)
    x = utils.to_tensor(img).unsqueeze(0)
    out2, tr = traced_run(x, 3, 0.352, 0.768, 6, 1)
    assert np.array_equal(out, utils.to_array(out2)), "manual replay != API"
    np.savez_compressed(
        os.path.join(HERE, "peacock_kat.npz"),
        mags=tr["mags"], interp=tr["interp"], theta_deg=tr["theta_deg"], sigma=tr["sigma"],
        rho=tr["rho"], kernel_center=tr["kernel"][:, :, 0, 12, 12],
        out_sub=out[::5, ::5].copy(), out_crop=out[200:264, 300:364].copy(),
        out_border=out[:40, :40].copy(),
        out_mean=np.float64(out.astype(np.float64).mean()),
        out_sumsq=np.float64((out.astype(np.float64) ** 2).sum()),
        out_min=out.min(), out_max=out.max(),
        iter_means=np.array([o.astype(np.float64).mean() for o in tr["out"]]),
    )

    # ---- 2. small full-output cases ---------------------------------------------------
    cases = {}
    g = torch.Generator().manual_seed(1234)
    inputs = {
        "mosaic_rgb_48x64": mosaic(2, 3, 48, 64, seed=1, block=8),
        "mosaic_gray_37x53": mosaic(1, 1, 37, 53, seed=2, sigma=(1.6, 0.6), theta_deg=110.0, block=7),
        "white_rgb_40x41": torch.rand(3, 3, 40, 41, generator=g),
        "mosaic_c2_64x48": mosaic(2, 2, 64, 48, seed=3, sigma=(3.0, 2.0), theta_deg=65.0, block=9),
        # pad 12 > image size: replicate pad still works in the reference
        "tiny_rgb_8x8": (0.5 * torch.rand(1, 3, 8, 8, generator=g)
                         + 0.05 * torch.arange(8.0).view(1, 1, 1, 8)).contiguous(),
        "mosaic_rgb_96x120": mosaic(2, 3, 96, 120, seed=5, sigma=(2.5, 1.2), theta_deg=30.0, block=12),
    }
    for name, x in inputs.items():
        cases[name + "/in"] = x.numpy()
        for tag, (n_iter, alpha, beta) in {"a6b1n3": (3, 6, 1), "a2b3n1": (1, 2, 3)}.items():
            y = polyblur.polyblur_deblurring(x, n_iter=n_iter, alpha=alpha, beta=beta)
            y2, tr = traced_run(x, n_iter, 0.352, 0.768, alpha, beta)
            assert torch.equal(y, y2)
            cases[f"{name}/{tag}/out"] = y.numpy()
            for k in ("mags", "theta_deg", "sigma", "rho"):
                cases[f"{name}/{tag}/{k}"] = tr[k]
            if tag == "a6b1n3":
                cases[f"{name}/{tag}/kernel0"] = tr["kernel"][0]
                cases[f"{name}/{tag}/interp"] = tr["interp"]
    np.savez_compressed(os.path.join(HERE, "small_cases.npz"), **cases)

    # ---- 3. stage-level vectors ---------------------------------------------------------
    st = {}
    x = inputs["mosaic_rgb_48x64"]
    gx, gy = filters.fourier_gradients(x)
    st["grad/in"] = x.numpy(); st["grad/gx"] = gx.numpy(); st["grad/gy"] = gy.numpy()
    xo = mosaic(1, 1, 35, 54, seed=7, block=6)       # odd H, even W
    gx, gy = filters.fourier_gradients(xo)
    st["grad_odd/in"] = xo.numpy(); st["grad_odd/gx"] = gx.numpy(); st["grad_odd/gy"] = gy.numpy()
    # kernels for a grid of parameters
    th = torch.tensor([0.0, 18.0, 24.0, 90.0, 150.0, 174.0, 45.0, 66.0]).view(-1, 1) * np.pi / 180
    sg = torch.tensor([0.3, 0.7, 2.88, 4.0, 1.39, 0.36, 4.0, 2.0]).view(-1, 1)
    rh = torch.tensor([0.3, 0.3, 1.76, 1.0, 0.80, 0.30, 4.0, 0.5]).view(-1, 1)
    st["kern/theta"] = th.numpy()[:, 0]; st["kern/sigma"] = sg.numpy()[:, 0]; st["kern/rho"] = rh.numpy()[:, 0]
    kk = blur_estimation.create_gaussian_filter(th, sg, rh, ksize=25)
    st["kern/k"] = kk.numpy()
    # deconvolution with given kernels (alpha=6, beta=1 and alpha=2, beta=3)
    xd = mosaic(8, 3, 40, 56, seed=8, block=8)
    st["deconv/in"] = xd.numpy()
    st["deconv/a6b1"] = deblurring.inverse_filtering_rank3(xd, kk, alpha=6, b=1, method="fft").numpy()
    st["deconv/a2b3"] = deblurring.inverse_filtering_rank3(xd, kk, alpha=2, b=3, method="fft").numpy()
    # estimator taps on the white-noise case (sigma = rho = 0.3 clamp)
    m = torch.tensor([[0.338, 0.428, 0.530, 0.499, 0.484, 0.395, 0.338],
                      [1.46, 1.50, 1.53, 1.47, 1.49, 1.52, 1.46],
                      [0.12, 0.10, 0.08, 0.11, 0.15, 0.14, 0.12]])
    thetas = torch.linspace(0, 180, 7).unsqueeze(0).long()
    interp = torch.arange(0, 180, 6).unsqueeze(0).long()
    mn, mo, tt = blur_estimation.find_maximal_blur_direction(m, thetas, interp)
    s_, r_ = blur_estimation.compute_gaussian_parameters(mn, mo, c=0.352, b=0.768)
    st["dir/mags"] = m.numpy(); st["dir/m_n"] = mn.numpy()[:, 0]; st["dir/m_o"] = mo.numpy()[:, 0]
    st["dir/theta"] = tt.numpy()[:, 0]; st["dir/sigma"] = s_.numpy()[:, 0]; st["dir/rho"] = r_.numpy()[:, 0]
    np.savez_compressed(os.path.join(HERE, "stages.npz"), **st)

    # ---- 4. optional flags on one small image -----------------------------------------
    op = {}
    x = mosaic(2, 3, 60, 72, seed=11, block=10)
    x = (x * 1.08).clamp(0, 1)                       # some saturated pixels for discard_saturation
    op["in"] = x.numpy()
    kw = dict(n_iter=2, alpha=6, beta=1)
    op["default"] = polyblur.polyblur_deblurring(x, **kw).numpy()
    op["remove_halo"] = polyblur.polyblur_deblurring(x, remove_halo=True, **kw).numpy()
    op["edgetaping"] = polyblur.polyblur_deblurring(x, edgetaping=True, **kw).numpy()
    op["prefiltering"] = polyblur.polyblur_deblurring(x, prefiltering=True, **kw).numpy()
    op["discard_saturation"] = polyblur.polyblur_deblurring(x, discard_saturation=True, **kw).numpy()
    # q > 0 only works for B == 1 in the reference: the (B,1,1) quantiles broadcast
    # against (B,1,H,W) as (1,B,1,1) (blur_estimation.py:104-109) and B > 1 raises.
    op["q1e-2_b1"] = polyblur.polyblur_deblurring(x[:1], q=1e-2, **kw).numpy()
    op["all_b1"] = polyblur.polyblur_deblurring(x[:1], remove_halo=True, edgetaping=True, prefiltering=True,
                                                discard_saturation=True, q=1e-2, **kw).numpy()
    op["all_q0"] = polyblur.polyblur_deblurring(x, remove_halo=True, edgetaping=True, prefiltering=True,
                                                discard_saturation=True, **kw).numpy()
    op["bilateral"] = filters.bilateral_filter(x).numpy()
    op["rf_s2_r0.8_n1"] = domain_transform.recursive_filter(x, sigma_s=2.0, sigma_r=0.8, num_iterations=1).numpy()
    op["rf_s60_r0.4_n3"] = domain_transform.recursive_filter(x, sigma_s=60, sigma_r=0.4, num_iterations=3).numpy()
    kk2 = blur_estimation.create_gaussian_filter(th[:2], sg[2:4], rh[2:4], ksize=25)
    op["edgetaper/k"] = kk2.numpy()
    op["edgetaper/out"] = edgetaper.edgetaper(utils.pad_with_kernel(x, kk2), kk2, method="fft").numpy()
    # module surface (different defaults: b=0.468, beta=4)
    mod = polyblur.PolyblurDeblurring()
    op["module_default"] = mod(x, n_iter=2).numpy()
    np.savez_compressed(os.path.join(HERE, "options.npz"), **op)

    # ---- 5. ndarray surface -----------------------------------------------------------
    nd = {}
    a = inputs["mosaic_rgb_48x64"][0].permute(1, 2, 0).numpy().copy()
    nd["hwc_in"] = a
    nd["hwc_out"] = polyblur.polyblur_deblurring(a, n_iter=2, alpha=6, beta=1)
    nd["hw_in"] = a[..., 0].copy()
    nd["hw_out"] = polyblur.polyblur_deblurring(a[..., 0].copy(), n_iter=2, alpha=6, beta=1)
    np.savez_compressed(os.path.join(HERE, "ndarray_api.npz"), **nd)

    tot = sum(os.path.getsize(os.path.join(HERE, f)) for f in os.listdir(HERE) if f.endswith(".npz"))
    print("golden written, total npz bytes:", tot)


if __name__ == "__main__":
    main()
