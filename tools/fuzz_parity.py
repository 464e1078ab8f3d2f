#!/usr/bin/env python
"""Random-shape parity sweep (GPU): polyblur_deblurring against the CPU oracle for odd, prime,
tiny and non-square shapes, channel counts 1-4, both kinds of content, random options."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import polyblur_b200 as pb  # noqa: E402
from oracle import polyblur_oracle as po  # noqa: E402


def main(n_cases=40, seed=0, plan_sides=False):
    """plan_sides: one side is a length with a compile-time plan (1080 / 1920 / 2160 / 3840: the fused-stage kernels of
    csrc/estimate3.cu and the second-generation FFT passes), the other side random."""
    rng = np.random.default_rng(seed)
    worst = 0.0            # worst error against the float32 oracle among the cases inside 1e-5 of it
    worst_arb = 0.0        # worst error against the float64 restatement among the arbitrated cases
    n_arb = 0
    fails = []
    for case in range(n_cases):
        B = int(rng.integers(1, 4))
        C = int(rng.choice([1, 2, 3, 3, 3, 4]))
        H = int(rng.choice([1, 2, 3, 5, 8, 13, 24, 37, 64, 97, 100, 131, 180, 256]))
        W = int(rng.choice([1, 2, 4, 7, 8, 16, 25, 53, 64, 101, 120, 128, 200, 243]))
        if plan_sides:
            B = int(rng.integers(1, 3))
            C = int(rng.choice([1, 3, 3, 3]))
            other = int(rng.choice([2, 3, 9, 16, 31, 64, 100, 150, 257]))
            if rng.random() < 0.5:
                H, W = int(rng.choice([1080, 2160])), other
            else:
                H, W = other, int(rng.choice([1920, 3840]))
        if H * W < 4:
            H = 4
        n_iter = int(rng.integers(1, 3 if plan_sides else 5))
        kind = rng.choice(["noise", "blocks", "smooth"])
        x = rng.random((B, C, H, W), dtype=np.float32)
        if kind == "blocks":
            x = np.repeat(np.repeat(rng.random((B, C, -(-H // 9), -(-W // 9)), dtype=np.float32), 9, -2), 9, -1)[..., :H, :W]
            k = po.gaussian_filter_np((float(rng.uniform(0.5, 3)), float(rng.uniform(0.4, 1.5))), float(rng.uniform(0, 3)))
            x = np.clip(po.convolve2d_fft(x, np.broadcast_to(k[None, None], (B, 1, 25, 25))), 0, 1).astype(np.float32)
        elif kind == "smooth":
            x = (0.5 + 0.4 * np.sin(np.linspace(0, 9, W))[None, None, None, :] * np.cos(np.linspace(0, 5, H))[None, None, :, None]
                 + 0.05 * x).astype(np.float32)
        x = np.ascontiguousarray(np.clip(x, 0, 1))
        kw = {}
        if rng.random() < 0.3:
            kw["discard_saturation"] = True
        if rng.random() < 0.2:
            kw["ker_size"] = int(rng.choice([9, 15, 21]))
        if rng.random() < 0.2:
            kw["prefiltering"] = True
        if rng.random() < 0.15:
            kw["remove_halo"] = True
        if rng.random() < 0.15:
            kw["edgetaping"] = True
        ab = [(6, 1), (2, 3), (2, 4)][int(rng.integers(0, 3))]
        extra = {}
        try:
            with np.errstate(all="ignore"):
                ref = po.polyblur_deblurring(x, n_iter=n_iter, alpha=ab[0], beta=ab[1], **kw)
            out = pb.polyblur_deblurring(torch.from_numpy(x).cuda(), n_iter=n_iter, alpha=ab[0], beta=ab[1], **kw).cpu().numpy()
            if not np.isfinite(ref).all():
                status, err = "ref-nonfinite", float("nan")      # constant image: NaN in the reference too
            else:
                err = float(np.abs(out.astype(np.float64) - ref).max())
                status = "ok" if err < 1e-5 else "MISMATCH"
                if status != "ok":
                    # the float32 oracle carries its own rounding: arbitrate with the float64 restatement
                    with np.errstate(all="ignore"):
                        truth = po.polyblur_deblurring(x, n_iter=n_iter, alpha=ab[0], beta=ab[1], dtype=np.float64, **kw)
                    extra = {"err_vs_f64": float(np.abs(out.astype(np.float64) - truth).max()),
                             "oracle_f32_vs_f64": float(np.abs(ref.astype(np.float64) - truth).max())}
                    if extra["err_vs_f64"] < 1e-5:
                        status = "ok-vs-f64"
        except Exception as exc:                 # noqa: BLE001
            status, err = f"EXC {type(exc).__name__}: {exc}", float("nan")
        rec = {"case": case, "shape": [B, C, H, W], "n_iter": n_iter, "kind": str(kind), "ab": ab, **kw, "err": err,
               "status": status, **extra}
        print(json.dumps(rec), flush=True)
        if status == "ok":
            worst = max(worst, err)
        elif status == "ok-vs-f64":
            n_arb += 1
            worst_arb = max(worst_arb, extra["err_vs_f64"])
        elif status != "ref-nonfinite":
            fails.append(rec)
    print(json.dumps({"worst_ok_err": worst, "arbitrated": n_arb, "worst_arbitrated_err_vs_f64": worst_arb,
                      "failures": len(fails)}))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "plan-sides":
        main(n_cases=int(sys.argv[2]) if len(sys.argv) > 2 else 16, seed=1, plan_sides=True)
    else:
        main()
