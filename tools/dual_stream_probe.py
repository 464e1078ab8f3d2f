#!/usr/bin/env python
"""C2 step time with the batch split over S concurrent graph branches (each branch = one pb_polyblur_f32 over B / S images
with its own workspace, forked and joined by events inside one CUDA graph): do the other branches' kernels fill the tail
wave of every kernel?  Usage: dual_stream_probe.py [S ...]"""
import ctypes as C
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import polyblur_b200 as pb  # noqa: E402
from polyblur_b200 import _lib, synthetic  # noqa: E402


def main():
    B, H, W = 32, 1080, 1920
    dev = torch.device("cuda", 0)
    splits = [int(a) for a in sys.argv[1:]] or [1, 2, 4]
    for kind in ("mosaic", "white"):
        x = synthetic.make(kind, B, 3, H, W, device="cuda")
        ref = None
        for S in splits:
            n = B // S
            p = _lib.default_params()
            p.n_iter, p.alpha, p.beta = 3, 6, 1
            out = torch.empty_like(x)
            wss = [_lib.workspace(n, 3, H, W, p, dev) for _ in range(S)]
            streams = [torch.cuda.Stream(dev) for _ in range(S)]

            def enqueue():
                cur = torch.cuda.current_stream(dev)
                for s, st in enumerate(streams):
                    st.wait_stream(cur)
                    with torch.cuda.stream(st):
                        rc = _lib.lib().pb_polyblur_f32(x[s * n:(s + 1) * n].data_ptr(), out[s * n:(s + 1) * n].data_ptr(), n, 3, H, W,
                                                        C.byref(p), wss[s].data_ptr(), wss[s].numel(), None, st.cuda_stream)
                        _lib.check(rc, "pb_polyblur_f32")
                for st in streams:
                    cur.wait_stream(st)

            side = torch.cuda.Stream(dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                enqueue()
            torch.cuda.current_stream(dev).wait_stream(side)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                enqueue()
            for _ in range(3):
                g.replay()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                g.replay()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 10
            same = None
            if ref is None:
                ref = out.clone()
            else:
                same = bool(torch.equal(ref, out))
            print(json.dumps({"dist": kind, "branches": S, "ms": round(ms, 3), "Mpix_s": round(B * H * W / 1e3 / ms),
                              "bitwise_equal_to_first": same}), flush=True)
            del g, wss


if __name__ == "__main__":
    main()
