"""Round-2 golden fixtures, generated from the LIVE reference (build container only):

    python tests/golden/make_golden_r2.py        ->  tests/golden/round2.npz

* ``nc/*``     outputs of the reference's native normalized convolution, JIT-compiled from
               /root/reference/polyblur/domain_transform/NC.cpp:143-204 where it lies (B = 1, C = 3 as written
               there); pins oracle.normalized_convolution and csrc/nc.cu.
* ``halo/*``   inverse_filtering_rank3 with remove_halo=True and grad_img=None, with and without do_edgetaper.
* ``asym/*``   inverse_filtering_rank3 (deblurring.py:211-239, method='fft'), edgetaper (edgetaper.py:26-33) and the
               autograd gradients of the deconvolution for kernels that are NOT point-symmetric (a motion streak and a
               shifted anisotropic Gaussian): the p2o / fft2 product (filters.py:255-273) is a convolution, and
               ``correlate=True`` rotates the kernel by 180 degrees (deblurring.py:229-230).

The reference is imported unmodified (skimage stub as in make_golden.py); tests never import it.
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

from polyblur import deblurring, edgetaper, filters, utils  # noqa: E402  (the reference)

torch.set_num_threads(1)


def textured(B, C, H, W, seed):
    g = torch.Generator().manual_seed(seed)
    small = torch.rand(B, C, -(-H // 7), -(-W // 7), generator=g)
    img = small.repeat_interleave(7, -2).repeat_interleave(7, -1)[..., :H, :W]
    yy = torch.linspace(0, 5, H)[:, None]
    xx = torch.linspace(0, 7, W)[None, :]
    img = 0.25 + 0.5 * img + 0.1 * torch.sin(3 * xx + yy) * torch.cos(2 * yy - xx) + 0.02 * torch.rand(B, C, H, W, generator=g)
    return img.clamp(0, 1).contiguous()


def asym_kernels(ks):
    """(2,1,ks,ks): a diagonal motion streak through one quadrant and a Gaussian shifted off-centre."""
    h = ks // 2
    k0 = torch.zeros(ks, ks)
    for t in range(0, h - 1):
        k0[h - t // 2, h + t] = 1.0 + 0.15 * t            # up-right streak, brighter at the far end
    k0[h + 1, h - 1] = 0.35
    k0 /= k0.sum()
    yy, xx = torch.meshgrid(torch.arange(ks) - h, torch.arange(ks) - h, indexing="ij")
    u = (xx - 2.3) * 0.9 + (yy + 1.4) * 0.4
    v = -(xx - 2.3) * 0.4 + (yy + 1.4) * 0.9
    k1 = torch.exp(-0.5 * (u ** 2 / 2.0 ** 2 + v ** 2 / 0.8 ** 2))
    k1 /= k1.sum()
    return torch.stack([k0, k1])[:, None].contiguous()


def main():
    from torch.utils.cpp_extension import load
    G = {}
    # ---- normalized convolution from the compiled NC.cpp ---------------------------------------------------------
    nc = load(name="nc_ref", sources=[os.path.join(REF, "polyblur", "domain_transform", "NC.cpp")],
              build_directory=os.environ.get("NC_BUILD_DIR", "/tmp/nc_build"), verbose=False)
    os.makedirs(os.environ.get("NC_BUILD_DIR", "/tmp/nc_build"), exist_ok=True)
    for tag, shape, ss, sr, n, seed in [("a", (1, 3, 10, 14), 8.0, 0.5, 1, 1), ("b", (1, 3, 12, 9), 3.0, 0.3, 2, 2),
                                        ("c", (1, 3, 48, 40), 20.0, 0.4, 3, 3), ("d", (1, 3, 33, 57), 60.0, 0.4, 1, 4)]:
        x = textured(*shape, seed=seed)
        y = nc.normalized_convolution(x.clone(), ss, sr, n)
        G[f"nc/{tag}/in"] = x.numpy()
        G[f"nc/{tag}/par"] = np.array([ss, sr, n], np.float64)
        G[f"nc/{tag}/out"] = y.numpy()
        print("nc", tag, shape, float(y.mean()))

    # ---- asymmetric kernels -----------------------------------------------------------------------------------------
    x = textured(2, 3, 61, 83, seed=11)
    for ks in (25, 9):
        k = asym_kernels(ks)
        G[f"asym/k{ks}"] = k.numpy()
        G[f"asym/in{ks}"] = x.numpy()
        for ab, (alpha, beta) in {"a6b1": (6, 1), "a2b3": (2, 3)}.items():
            for corr in (False, True):
                y = deblurring.inverse_filtering_rank3(x, k, alpha=alpha, b=beta, correlate=corr, method="fft")
                G[f"asym/deconv{ks}/{ab}/{'corr' if corr else 'conv'}"] = y.numpy()
        y = deblurring.inverse_filtering_rank3(x, k, alpha=6, b=1, do_edgetaper=True, method="fft")
        G[f"asym/deconv{ks}/taper"] = y.numpy()
        xp = utils.pad_with_kernel(x, k, mode="replicate")
        G[f"asym/edgetaper{ks}"] = edgetaper.edgetaper(xp, k, n_tapers=3).numpy()
    # gradients of <w, deconv(x, k)> with respect to x and k (autograd over the reference)
    k = asym_kernels(25)
    xg = x.clone().requires_grad_(True)
    kg = k.clone().requires_grad_(True)
    w = textured(2, 3, 61, 83, seed=12) - 0.5
    y = deblurring.inverse_filtering_rank3(xg, kg, alpha=6, b=1, method="fft")
    (y * w).sum().backward()
    G["asym/vjp/w"] = w.numpy()
    G["asym/vjp/gx"] = xg.grad.numpy()
    G["asym/vjp/gk"] = kg.grad.numpy()
    print("vjp", float(xg.grad.abs().max()), float(kg.grad.abs().max()))

    # ---- halo masking without grad_img: gradients of the image inverse_filtering_rank3 hands over, i.e. of the crop of
    #      the padded (and, with do_edgetaper, tapered) image (deblurring.py:237-238, 200-203) ---------------------------
    from polyblur import blur_estimation
    kk = blur_estimation.create_gaussian_filter(torch.tensor([[0.4], [1.9]]), torch.tensor([[1.4], [2.2]]),
                                                torch.tensor([[0.7], [1.1]]), ksize=25)
    G["halo/k"] = kk.numpy()
    for taper in (False, True):
        y = deblurring.inverse_filtering_rank3(x, kk, alpha=6, b=1, remove_halo=True, do_edgetaper=taper, method="fft")
        G[f"halo/{'taper' if taper else 'plain'}"] = y.numpy()

    # ---- method= is accepted and 'direct' differs upstream (B = 1 only there): record the fft result the
    #      drop-in returns for every method ----------------------------------------------------------
    np.savez_compressed(os.path.join(HERE, "round2.npz"), **G)
    print("wrote round2.npz with", len(G), "arrays")


if __name__ == "__main__":
    main()
