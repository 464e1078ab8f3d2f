"""Command-line demo with the arguments of the reference's main.py (main.py:30-55), on the CUDA
engine:  python -m polyblur_b200.main --impath tests/golden/peacock_defocus.png --N 3 --alpha 6 --beta 1

Differences: Pillow instead of scikit-image for PNG I/O, no matplotlib window, ``--out`` for the
result path, and the 8-bit conversions run on the device (polyblur_b200.io)."""
from __future__ import annotations

import argparse
import os
import time

import numpy as np
import torch

from . import PolyblurDeblurring, filters, io as pbio, utils


def str2bool(v):
    v = str(v)
    if v.lower() in ('yes', 'true', 't', 'y', '1'):
        return True
    if v.lower() in ('no', 'false', 'f', 'n', '0'):
        return False
    raise argparse.ArgumentTypeError('Boolean value expected.')


def build_parser():
    ap = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    ap.add_argument('--impath', type=str, required=True, help='input image')
    ap.add_argument('--out', type=str, default=None, help='output PNG (default results/restored_alpha_A_beta_B.png)')
    ap.add_argument('--synthetic_degradation', type=str2bool, default=False)
    ap.add_argument('--sigma', type=float, default=3.0)
    ap.add_argument('--rho', type=float, default=1.0)
    ap.add_argument('--theta', type=float, default=0.0)
    ap.add_argument('--sigma_n', type=float, default=0.01)
    ap.add_argument('--N', type=int, default=3)
    ap.add_argument('--alpha', type=int, default=2)
    ap.add_argument('--beta', type=int, default=3)
    ap.add_argument('--q', type=float, default=0)
    ap.add_argument('--do_prefiltering', type=str2bool, default=False)
    ap.add_argument('--do_halo_removal', type=str2bool, default=False)
    ap.add_argument('--do_edgetaping', type=str2bool, default=False)
    ap.add_argument('--do_patch_decomposition', type=str2bool, default=False)
    ap.add_argument('--patch_size', type=int, default=400)
    ap.add_argument('--patch_overlap', type=float, default=0.25)
    return ap


def synthetic_blur(img, sigma, rho, theta_deg, sigma_n, seed=0):
    """main.py:89-96: circular blur with filters.gaussian_filter + Gaussian noise, on (H,W,C) float."""
    k = filters.gaussian_filter((sigma, rho), theta=theta_deg * np.pi / 180, k_size=np.array([25, 25]))
    H, W = img.shape[:2]
    kp = np.zeros((H, W), np.float64)
    kp[:25, :25] = k
    kp = np.roll(kp, (-12, -12), axis=(0, 1))
    K = np.fft.rfft2(kp)
    out = np.stack([np.fft.irfft2(np.fft.rfft2(img[..., c]) * np.conj(K), s=(H, W)) for c in range(img.shape[-1])], -1)
    out = out + sigma_n * np.random.default_rng(seed).standard_normal(out.shape)
    return np.clip(out, 0.0, 1.0).astype(np.float32)


def main(argv=None):
    from PIL import Image
    args = build_parser().parse_args(argv)
    im = Image.open(args.impath)
    if im.mode not in ("L", "RGB"):
        im = im.convert("RGB")                        # rgba2rgb, palettes
    img_u8 = np.asarray(im)
    print('Processing a (%d,%d) image.' % (img_u8.shape[1], img_u8.shape[0]))
    c, b = 0.362, 0.468                               # main.py:105-106
    kw = dict(n_iter=args.N, c=c, b=b, alpha=args.alpha, beta=args.beta, remove_halo=args.do_halo_removal,
              prefiltering=args.do_prefiltering, edgetaping=args.do_edgetaping, q=args.q)
    simple = not args.synthetic_degradation and not args.do_patch_decomposition
    for label in ('Mock run (library load, allocator warm-up).', 'Real run.'):
        print(label)
        start = time.time()
        if simple:
            out_u8 = pbio.deblur_uint8(img_u8, **kw)   # 8-bit in, 8-bit out, conversions on the device
        else:
            img = utils.to_float(img_u8)
            if img.ndim == 2:
                img = img[..., None]
            if args.synthetic_degradation:
                img = synthetic_blur(img, args.sigma, args.rho, args.theta, args.sigma_n)
            mod = PolyblurDeblurring(patch_decomposition=args.do_patch_decomposition, patch_size=args.patch_size,
                                     patch_overlap=args.patch_overlap, batch_size=20)
            x = utils.to_tensor(img).unsqueeze(0).cuda()
            with torch.no_grad():
                y = mod(x, **kw)
            out_u8 = utils.to_ubyte(utils.to_array(y))
        torch.cuda.synchronize()
        print('Restoration took %2.4f seconds' % (time.time() - start))
    out = args.out or os.path.join('results', 'restored_alpha_%d_beta_%d.png' % (args.alpha, args.beta))
    os.makedirs(os.path.dirname(out) or '.', exist_ok=True)
    Image.fromarray(np.squeeze(out_u8)).save(out)
    print('saved', out)
    return out


if __name__ == '__main__':
    main()
