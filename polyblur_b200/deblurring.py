"""The two public surfaces of the reference, kept as drop-ins:

    polyblur_deblurring(img, n_iter, c, b, alpha, beta, ...)   polyblur/deblurring.py:23-96
    PolyblurDeblurring(nn.Module)                              polyblur/deblurring.py:250-394

Host code only: argument handling mirrors the reference, all arithmetic happens in
libpolyblur_sm100.so (hand-written sm_100a kernels) on the tensor's CUDA device, enqueued
on torch's current stream with no host synchronisation.  ndarray / CPU tensor input is
copied to the GPU and back.  There is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np
import torch
from torch import nn

from . import _lib, sharding, utils

_METHODS = ("fft", "direct", "direct_separable")


def _make_params(n_iter, c, b, alpha, beta, sigma_r, sigma_s, ker_size, q, remove_halo, edgetaping,
                 prefiltering, discard_saturation, engine=_lib.ENGINE_AUTO, prefilter="bilateral",
                 tap_rel_threshold=0.0, chunk_images=0):
    p = _lib.default_params()
    p.n_iter = int(n_iter)
    p.c, p.b, p.alpha, p.beta = float(c), float(b), float(alpha), float(beta)
    p.sigma_r, p.sigma_s, p.q = float(sigma_r), float(sigma_s), float(q)
    p.ker_size = int(ker_size)
    flags = 0
    if remove_halo:
        flags |= _lib.FLAG_REMOVE_HALO
    if edgetaping:
        flags |= _lib.FLAG_EDGETAPER | _lib.FLAG_EDGETAPER_BATCHMAX
    if prefiltering:
        flags |= _lib.FLAG_PREFILTER_RF if prefilter == "rf" else _lib.FLAG_PREFILTER
    if discard_saturation:
        flags |= _lib.FLAG_DISCARD_SATURATION
    p.flags = flags
    p.engine = int(engine)
    p.tap_rel_threshold = float(tap_rel_threshold)
    p.chunk_images = int(chunk_images)
    if p.chunk_images > 0 and edgetaping:
        p.flags &= ~_lib.FLAG_EDGETAPER_BATCHMAX        # groups cannot share the reference's batch-global maximum
    return p


def polyblur_device(x: torch.Tensor, p: "_lib.PbParams", out: torch.Tensor | None = None,
                    return_estimates: bool = False):
    """Run the whole Polyblur loop on a contiguous float32 CUDA tensor (B,C,H,W).

    Thin wrapper over ``pb_polyblur_f32``; everything is enqueued on the current stream.
    """
    assert x.is_cuda and x.dtype == torch.float32 and x.is_contiguous() and x.ndim == 4
    B, Cn, H, W = x.shape
    dev = x.device
    with torch.cuda.device(dev):
        if out is None:
            out = torch.empty_like(x)
        ws = _lib.workspace(B, Cn, H, W, p, dev)
        est = None
        if return_estimates:
            est = torch.empty(max(p.n_iter, 1), B, _lib.PB_EST_STRIDE, dtype=torch.float32, device=dev)
        rc = _lib.lib().pb_polyblur_f32(x.data_ptr(), out.data_ptr(), B, Cn, H, W, C.byref(p),
                                        ws.data_ptr(), ws.numel(), _lib.ptr(est), _lib.stream_ptr(dev))
        _lib.check(rc, "pb_polyblur_f32")
    return (out, est) if return_estimates else out


class GraphedPolyblur:
    """One ``pb_polyblur_f32`` enqueue for a fixed shape and parameter set, captured into a CUDA
    graph.  The loop makes ~20 kernel launches per iteration and never synchronises the host (the
    per-image engine choice happens on the device), so the whole call replays as one graph launch:
    worth ~0.2-0.3 ms per call, i.e. most of the time for small images.

        g = GraphedPolyblur((B, C, H, W), n_iter=3, alpha=6, beta=1)
        y = g(x)            # copies x into the static input, replays, returns the static output
    """

    def __init__(self, shape, device=None, params=None, uint8_io=False, **kw):
        B, Cn, H, W = shape
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        if params is not None:
            self.params = _lib.PbParams.from_buffer_copy(bytes(params))
        else:
            self.params = _make_params(kw.pop("n_iter", 1), kw.pop("c", 0.352), kw.pop("b", 0.768), kw.pop("alpha", 2),
                                       kw.pop("beta", 3), kw.pop("sigma_r", 0.8), kw.pop("sigma_s", 2.0),
                                       kw.pop("ker_size", 25), kw.pop("q", 0.0), kw.pop("remove_halo", False),
                                       kw.pop("edgetaping", False), kw.pop("prefiltering", False),
                                       kw.pop("discard_saturation", False), **kw)
        if self.params.n_iter < 1:
            raise ValueError("n_iter must be >= 1")
        self.uint8_io = bool(uint8_io)
        with torch.cuda.device(dev):
            self.x = torch.zeros(B, Cn, H, W, dtype=torch.float32, device=dev)
            self.out = torch.empty_like(self.x)
            self.ws = _lib.workspace(B, Cn, H, W, self.params, dev)
            self.x.uniform_(0, 1)                 # a non-constant image for the warm-up call
            if self.uint8_io:
                # 8-bit HWC in / out (csrc/io.cu) inside the same graph: x_u8 -> x -> Polyblur -> out -> out_u8
                self.x_u8 = (self.x.permute(0, 2, 3, 1) * 255).to(torch.uint8).contiguous()
                self.out_u8 = torch.empty_like(self.x_u8)
            side = torch.cuda.Stream(dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                self._enqueue(side.cuda_stream)    # warm-up: sets kernel attributes outside the capture
            torch.cuda.current_stream(dev).wait_stream(side)
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self._enqueue(torch.cuda.current_stream(dev).cuda_stream)

    def _enqueue(self, stream):
        B, Cn, H, W = self.x.shape
        if self.uint8_io:
            _lib.check(_lib.lib().pb_u8hwc_to_f32nchw(self.x_u8.data_ptr(), self.x.data_ptr(), B, H, W, Cn, stream),
                       "pb_u8hwc_to_f32nchw")
        rc = _lib.lib().pb_polyblur_f32(self.x.data_ptr(), self.out.data_ptr(), B, Cn, H, W, C.byref(self.params),
                                        self.ws.data_ptr(), self.ws.numel(), None, stream)
        _lib.check(rc, "pb_polyblur_f32")
        if self.uint8_io:
            _lib.check(_lib.lib().pb_f32nchw_to_u8hwc(self.out.data_ptr(), self.out_u8.data_ptr(), B, Cn, H, W, stream),
                       "pb_f32nchw_to_u8hwc")

    def __call__(self, x=None):
        if x is not None:
            (self.x_u8 if self.uint8_io and x.dtype == torch.uint8 else self.x).copy_(x, non_blocking=True)
        self.graph.replay()
        return self.out_u8 if self.uint8_io else self.out


# ---- host <-> device pipeline with cached, graph-captured chunk engines --------------------------------------------
# The chunked host paths (CPU tensor in -> CPU tensor out) run one engine call per chunk.  Enqueueing a chunk eagerly
# costs ~60 launches with idle gaps between the small kernels (4 x 8 images: 9.4 ms against 8.0 ms for one call of 32,
# tools/e2e_probe.py) and allocates staging buffers per call.  Instead every (device, chunk shape, parameters, 8-bit?)
# gets two GraphedPolyblur engines (double buffering) that are kept across calls: the copies go straight into / out of
# their static buffers and a chunk is one graph launch.  A small LRU bounds the device memory this holds.
_ENGINE_CACHE: "dict[tuple, list]" = {}
_PIPE_COMPUTE_STREAMS = int(os.environ.get("PB_PIPE_STREAMS", "2"))   # compute streams the chunks alternate between
_ENGINE_CACHE_MAX_BYTES = 24 << 30          # device memory the cached engines may hold (per process)


def clear_cache() -> None:
    """Drops the cached chunk engines (graphs, staging buffers, workspaces) of the host pipelines."""
    _ENGINE_CACHE.clear()


def _chunk_engines(dev: torch.device, shape, p: "_lib.PbParams", uint8_io: bool):
    key = (dev.index, tuple(shape), bytes(p), bool(uint8_io))
    hit = _ENGINE_CACHE.pop(key, None)
    if hit is None:
        hit = [GraphedPolyblur(shape, device=dev, params=p, uint8_io=uint8_io) for _ in range(2)]
    _ENGINE_CACHE[key] = hit                       # most recently used last

    def held(engs):
        return sum(t.numel() * t.element_size() for g in engs for t in (g.x, g.out, g.ws))

    while len(_ENGINE_CACHE) > 1 and sum(held(v) for v in _ENGINE_CACHE.values()) > _ENGINE_CACHE_MAX_BYTES:
        _ENGINE_CACHE.pop(next(iter(_ENGINE_CACHE)))
    return hit


def _run_host_pipeline(x: torch.Tensor, host: torch.Tensor, sizes, p: "_lib.PbParams", dev: torch.device,
                       uint8_io: bool):
    """x, host: pinned CPU tensors, (B,C,H,W) float32 or (B,H,W,C) uint8; chunks of ``sizes`` images flow through
    three streams: H2D of chunk k+1 | graph of chunk k | D2H of chunk k-1.  One host synchronisation at the end."""
    bounds = [0]
    for n_k in sizes:
        bounds.append(bounds[-1] + n_k)
    with torch.cuda.device(dev):
        # two compute streams, chunks alternate between them: the kernels of a chunk are persistent grids that end in
        # a partial wave (a chunk of 8 images is 3.9 waves of the column pass), and the other chunk's kernels fill it
        s_in, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
        s_runs = [torch.cuda.Stream(dev) for _ in range(_PIPE_COMPUTE_STREAMS)]
        for st in [s_in, s_out] + s_runs:
            st.wait_stream(torch.cuda.current_stream(dev))
        # every chunk size's engines exist (captured on first use) before the first copy is in flight
        engines = {}
        for n in sorted(set(sizes)):
            shape = (n, x.shape[3], x.shape[1], x.shape[2]) if uint8_io else (n,) + tuple(x.shape[1:])
            engines[n] = _chunk_engines(dev, shape, p, uint8_io)
        used = {}                                   # (size, slot) -> [event: input consumed, event: output copied]
        for k, (a, b) in enumerate(zip(bounds, bounds[1:])):
            n = b - a
            slot = sum(1 for kk in range(k) if bounds[kk + 1] - bounds[kk] == n) & 1
            g = engines[n][slot]
            ev = used.setdefault((n, slot), [None, None])
            gin = g.x_u8 if uint8_io else g.x
            gout = g.out_u8 if uint8_io else g.out
            with torch.cuda.stream(s_in):
                if ev[0] is not None:
                    s_in.wait_event(ev[0])
                gin.copy_(x[a:b], non_blocking=True)
                loaded = s_in.record_event()
            s_run = s_runs[k % len(s_runs)]
            with torch.cuda.stream(s_run):
                s_run.wait_event(loaded)
                if ev[0] is not None:
                    s_run.wait_event(ev[0])         # the engine's previous replay (it may have run on the other stream)
                if ev[1] is not None:
                    s_run.wait_event(ev[1])
                g.graph.replay()
                done = s_run.record_event()
                ev[0] = done
            with torch.cuda.stream(s_out):
                s_out.wait_event(done)
                host[a:b].copy_(gout, non_blocking=True)
                ev[1] = s_out.record_event()
        s_out.synchronize()
        for st in s_runs:
            st.synchronize()
    return host


def polyblur_deblurring(img, n_iter=1, c=0.352, b=0.768, alpha=2, beta=3, sigma_r=0.8, sigma_s=2.0,
                        ker_size=25, q=0.0, n_angles=6, n_interpolated_angles=30, remove_halo=False,
                        edgetaping=False, prefiltering=False, discard_saturation=False,
                        multichannel_kernel=False, method='fft', verbose=False, **engine_kw):
    """Blind deblurring by polynomial reblurring; same signature and defaults as the
    reference (polyblur/deblurring.py:23-25).

    ``img``: (H,W) / (H,W,C) ndarray -> float32 ndarray squeezed like ``utils.to_array``;
    or a (B,C,H,W) float32 tensor -> tensor on the same device.  Every ``method`` gives the
    reference's ``'fft'`` result (its other methods are broken for batches).
    Extra keyword arguments (``engine``, ``prefilter``, ``tap_rel_threshold``, ``chunk_images``,
    ``return_estimates``) are engine knobs that the reference does not have.
    """
    if method not in _METHODS:
        raise ValueError(f"unknown method {method!r}; expected one of {_METHODS}")
    if n_angles != 6 or n_interpolated_angles != 30:
        raise ValueError("only n_angles=6 and n_interpolated_angles=30 work in the reference "
                         "(SURVEY.md Appendix B.10); other values are rejected here")
    return_estimates = bool(engine_kw.pop("return_estimates", False))
    estimate_grad = bool(engine_kw.pop("estimate_grad", True))
    p = _make_params(n_iter, c, b, alpha, beta, sigma_r, sigma_s, ker_size, q, remove_halo, edgetaping,
                     prefiltering, discard_saturation, **engine_kw)

    flag_numpy = isinstance(img, np.ndarray)
    if flag_numpy:
        x = utils.to_tensor(img).unsqueeze(0)
    else:
        x = img
        if not isinstance(x, torch.Tensor):
            raise TypeError("img must be a numpy array or a torch tensor")
        if x.ndim != 4:
            raise ValueError("tensor input must be (B,C,H,W)")
        if x.dtype != torch.float32:
            raise TypeError(f"float32 only (got {x.dtype}), like the reference")
    if n_iter == 0:                     # the reference returns its input object (:60,:96)
        return utils.to_array(x) if flag_numpy else img
    if not flag_numpy and x.requires_grad and torch.is_grad_enabled():
        # differentiable path (autograd.py): gradient with respect to the image, blur estimates held constant
        if edgetaping or return_estimates or q > 0 or (p.flags & _lib.FLAG_PREFILTER_RF):
            raise NotImplementedError("gradients are implemented for the default options, remove_halo, prefiltering "
                                      "(bilateral) and discard_saturation (no edgetaping / q / RF prefilter); call "
                                      "under torch.no_grad() or detach the input")
        from . import autograd as _autograd
        if remove_halo or prefiltering:
            return _autograd.polyblur_deblurring_halo_grad(x, n_iter=n_iter, c=c, b=b, alpha=alpha, beta=beta,
                                                           ker_size=ker_size, engine=p.engine,
                                                           estimate_grad=estimate_grad,
                                                           discard_saturation=discard_saturation,
                                                           remove_halo=remove_halo, prefiltering=prefiltering)
        return _autograd.polyblur_deblurring_grad(x, n_iter=n_iter, c=c, b=b, alpha=alpha, beta=beta,
                                                  ker_size=ker_size, engine=p.engine, estimate_grad=estimate_grad,
                                                  discard_saturation=discard_saturation)

    dev = _lib.require_cuda(x)
    src_device = x.device
    xd = x.detach()
    # the chunked host pipeline runs one engine call per chunk; the reference's edgetaper normalises its weights by
    # a batch-global max (edgetaper.py:15,21), which must see the whole batch, so that option takes the one-shot path
    batch_coupled = bool(p.flags & _lib.FLAG_EDGETAPER_BATCHMAX)
    if not xd.is_cuda and not flag_numpy and not return_estimates and xd.shape[0] >= 2 and not batch_coupled:
        return _polyblur_host_pipelined(xd, p, dev)
    if not xd.is_cuda and not flag_numpy:
        xd = xd.pin_memory() if not xd.is_pinned() and xd.numel() > (1 << 20) else xd
    xd = xd.to(dev, non_blocking=True).contiguous()
    res = polyblur_device(xd, p, return_estimates=return_estimates)
    out, est = res if return_estimates else (res, None)
    if flag_numpy:
        out = utils.to_array(out)
    elif src_device != dev:
        # device -> pinned host memory on the same stream, one synchronisation at the end
        host = torch.empty(out.shape, dtype=out.dtype, pin_memory=True)
        host.copy_(out, non_blocking=True)
        torch.cuda.current_stream(dev).synchronize()
        out = host
    if return_estimates:
        return out, est.to(src_device)
    return out


def _polyblur_host_pipelined(x: torch.Tensor, p: "_lib.PbParams", dev: torch.device, max_chunks: int = 16,
                             ramp=(1,)):
    """CPU tensor in -> CPU tensor out with the PCIe transfers hidden behind the kernels.

    Images are independent, so the batch is cut into chunks that flow through three streams:
    host->device copy of chunk k+1, the Polyblur kernels of chunk k (one CUDA-graph launch of a cached
    engine) and the device->host copy of chunk k-1 run concurrently (the link is full duplex).  Results are
    identical to the one-shot path; the only host synchronisation is the final one."""
    B = x.shape[0]
    x = x.contiguous()
    if not x.is_pinned():
        x = x.pin_memory()
    # float32 over PCIe is the bottleneck (1.6 GB per 32 x 1080p step against 7 ms of kernels): many small chunks,
    # with one-image chunks first and last so that only 1/B of the transfer is not overlapped
    sizes = sharding.pipeline_chunks(B, -(-B // max(1, min(max_chunks, B))), ramp)
    host = torch.empty(x.shape, dtype=torch.float32, pin_memory=True)
    return _run_host_pipeline(x, host, sizes, p, dev, uint8_io=False)


def inverse_filtering_rank3(img, kernel, alpha=2, b=4, correlate=False, remove_halo=False,
                            do_edgetaper=False, grad_img=None, method='direct', engine=_lib.ENGINE_AUTO,
                            edgetaper_batch_max=True):
    """pad -> [edgetaper] -> polynomial deconvolution on the torus -> crop -> [halo masking] ->
    clamp (polyblur/deblurring.py:211-239), for explicit kernels (B,1,k,k) or (1,1,k,k).

    Always the 'fft' (circular, replicate-padded) semantics.  ``correlate`` rotates the kernel by
    180 degrees like the reference; ``grad_img`` = (grad_x, grad_y) of the blurry image, computed
    from ``img`` when None (deblurring.py:200-203)."""
    if img.dtype != torch.float32 or img.ndim != 4:
        raise TypeError("img must be a float32 (B,C,H,W) tensor")
    if (img.requires_grad or (isinstance(kernel, torch.Tensor) and kernel.requires_grad)) and torch.is_grad_enabled():
        if do_edgetaper:
            raise NotImplementedError("gradients are not implemented through the edgetaper")
        if kernel.shape[1] != 1 or kernel.shape[-1] != kernel.shape[-2]:
            raise NotImplementedError("one square kernel per image (B,1,k,k) or (1,1,k,k)")
        from . import autograd as _autograd
        kk = torch.rot90(kernel, k=2, dims=(-2, -1)) if correlate else kernel
        if remove_halo:
            return _autograd.inverse_filtering_rank3_halo(img, kk, alpha, b, grad_img, engine)
        return _autograd.DeconvolutionFunction.apply(img, kk, alpha, b, engine, True)
    dev = _lib.require_cuda(img)
    src = img.device
    x = img.detach().to(dev).contiguous()
    B, Cn, H, W = x.shape
    k = kernel.detach().to(dev, torch.float32)
    if correlate:
        k = torch.rot90(k, k=2, dims=(-2, -1))
    ks = k.shape[-1]
    if k.shape[-2] != ks:
        raise ValueError("square kernels only")
    if k.shape[1] != 1:
        raise NotImplementedError("one kernel per image (B,1,k,k); per-channel kernels are not supported")
    k = k.expand(B, 1, ks, ks).contiguous()
    flags = 0
    if remove_halo:
        flags |= _lib.FLAG_REMOVE_HALO
    if do_edgetaper:
        flags |= _lib.FLAG_EDGETAPER | (_lib.FLAG_EDGETAPER_BATCHMAX if edgetaper_batch_max else 0)
    gx = gy = None
    if remove_halo and grad_img is not None:
        gx = grad_img[0].detach().to(dev, torch.float32).contiguous()
        gy = grad_img[1].detach().to(dev, torch.float32).contiguous()
        if gx.shape != x.shape or gy.shape != x.shape:
            raise ValueError("grad_img must hold two tensors of img's shape")
    with torch.cuda.device(dev):
        p = _lib.default_params()
        p.flags = flags
        p.ker_size = ks
        p.engine = int(engine)
        ws = _lib.workspace(B, Cn, H, W, p, dev)
        out = torch.empty_like(x)
        rc = _lib.lib().pb_deconv_ex_f32(x.data_ptr(), out.data_ptr(), B, Cn, H, W, k.data_ptr(), ks,
                                         float(alpha), float(b), int(engine), flags, _lib.ptr(gx), _lib.ptr(gy),
                                         ws.data_ptr(), ws.numel(), _lib.stream_ptr(dev))
        _lib.check(rc, "pb_deconv_ex_f32")
    return out.to(src)


class PolyblurDeblurring(nn.Module):
    """nn.Module wrapper with the reference's constructor and forward signature
    (polyblur/deblurring.py:250-347).  No parameters, no buffers.

    ``patch_decomposition=True`` (spatially varying blur: every patch gets its own estimate,
    Kaiser-window overlap-add, deblurring.py:269-340) raises NameError in the reference
    (``handling_saturation`` is undefined) and its overlap-add indexing is only
    right for an image batch of 1.  Here the intended flow is implemented for any batch: patches
    are cut on the device, deblurred ``batch_size`` patch positions at a time by the CUDA engine,
    and blended on the device.
    """

    def __init__(self, patch_decomposition=False, patch_size=400, patch_overlap=0.25, batch_size=1):
        super().__init__()
        self.batch_size = batch_size
        self.patch_decomposition = patch_decomposition
        self.patch_size = (patch_size, patch_size)
        self.patch_overlap = patch_overlap

    def forward(self, images, n_iter=1, c=0.352, b=0.468, alpha=2, beta=4, sigma_s=2, ker_size=25,
                sigma_r=0.4, q=0.0, n_angles=6, n_interpolated_angles=30, remove_halo=False,
                edgetaping=False, prefiltering=False, discard_saturation=False, multichannel_kernel=False,
                method='fft', device=None):
        kw = dict(n_iter=n_iter, c=c, b=b, alpha=alpha, beta=beta, ker_size=ker_size, sigma_s=sigma_s,
                  sigma_r=sigma_r, remove_halo=remove_halo, edgetaping=edgetaping, prefiltering=prefiltering,
                  discard_saturation=discard_saturation, multichannel_kernel=multichannel_kernel, method=method,
                  q=q, n_angles=n_angles, n_interpolated_angles=n_interpolated_angles)
        if device is not None and isinstance(images, torch.Tensor):
            images = images.to(device)
        if not self.patch_decomposition:
            return polyblur_deblurring(images, **kw)
        if not isinstance(images, torch.Tensor) or images.ndim != 4:
            raise ValueError("patch decomposition expects a (B,C,H,W) tensor")
        if images.requires_grad and torch.is_grad_enabled():
            raise NotImplementedError("gradients are not implemented through the patch decomposition; call under "
                                      "torch.no_grad() or detach the input")
        if images.dtype != torch.float32:
            raise TypeError(f"float32 only (got {images.dtype}), like the reference")
        src_device = images.device
        dev = _lib.require_cuda(images)
        x = images.detach().to(dev).contiguous()
        nimg, Cn, H0, W0 = x.shape
        ph, pw = self.patch_size
        # even dimensions (:273-279: the last row / column is dropped), centre replicate pad to a whole number of
        # steps (:282-287) -- both as index arithmetic inside the extraction kernel, nothing is copied
        h, w = H0 - (H0 % 2), W0 - (W0 % 2)
        step_h = int(ph * (1 - self.patch_overlap))
        step_w = int(pw * (1 - self.patch_overlap))
        if step_h < 1 or step_w < 1:
            raise ValueError("patch_overlap leaves no step between patches")
        new_h = int(np.ceil((h - ph) / step_h) * step_h) + ph
        new_w = int(np.ceil((w - pw) / step_w) * step_w) + pw
        if new_h < h or new_w < w:
            raise ValueError(f"patch_size {self.patch_size} is larger than the ({h}, {w}) image")
        pad_top, pad_left = int(np.floor((new_h - h) / 2)), int(np.floor((new_w - w) / 2))
        ny, nx = (new_h - ph) // step_h + 1, (new_w - pw) // step_w + 1
        npatch = ny * nx
        p = _make_params(n_iter, c, b, alpha, beta, sigma_r, sigma_s, ker_size, q, remove_halo, edgetaping,
                         prefiltering, discard_saturation)
        if method not in _METHODS:
            raise ValueError(f"unknown method {method!r}; expected one of {_METHODS}")
        if n_angles != 6 or n_interpolated_angles != 30:
            raise ValueError("only n_angles=6 and n_interpolated_angles=30 work in the reference "
                             "(SURVEY.md Appendix B.10); other values are rejected here")
        with torch.cuda.device(dev):
            st = _lib.stream_ptr(dev)
            wy, wx = self._window_1d(ph, dev), self._window_1d(pw, dev)
            patches = torch.empty(npatch * nimg, Cn, ph, pw, dtype=torch.float32, device=dev)
            geom = (nimg, Cn, h, w, ph, pw, step_h, step_w, ny, nx, pad_top, pad_left)
            rc = _lib.lib().pb_patch_extract_f32(x.data_ptr(), H0 * W0, W0, patches.data_ptr(), *geom, st)
            _lib.check(rc, "pb_patch_extract_f32")
            # every patch of every image is an independent "image" of the engine (its own blur estimate); the batch
            # goes through in as few calls as ~1 GB of patches per call allows (results do not depend on the split)
            done = torch.empty_like(patches)
            per_call = max(1, min(patches.shape[0], (1 << 30) // (Cn * ph * pw * 4)))
            if p.flags & _lib.FLAG_EDGETAPER_BATCHMAX:
                # the reference's edgetaper normalises by the max over the patches of one call (edgetaper.py:15,21):
                # keep its grouping, batch_size patch positions x the image batch
                per_call = max(1, self.batch_size) * nimg
            for a in range(0, patches.shape[0], per_call):
                if p.n_iter == 0:
                    done[a:a + per_call].copy_(patches[a:a + per_call])
                else:
                    polyblur_device(patches[a:a + per_call], p, out=done[a:a + per_call])
            restored = torch.empty(nimg, Cn, h, w, dtype=torch.float32, device=dev)
            rc = _lib.lib().pb_patch_blend_f32(done.data_ptr(), wy.data_ptr(), wx.data_ptr(), restored.data_ptr(), *geom, st)
            _lib.check(rc, "pb_patch_blend_f32")
        return restored.to(src_device)

    _WINDOWS: dict = {}

    def _window_1d(self, n, dev):
        """torch.kaiser_window(n, beta=5, periodic=True) (deblurring.py:349-366), cached per length and device."""
        key = (int(n), str(dev))
        if key not in self._WINDOWS:
            self._WINDOWS[key] = torch.kaiser_window(int(n), beta=5, periodic=True).to(dev).contiguous()
        return self._WINDOWS[key]

    def build_window(self, image_size, window_type='kaiser'):
        """Separable 2-D window (deblurring.py:349-366)."""
        H, W = image_size
        makers = {'kaiser': lambda n: torch.kaiser_window(n, beta=5, periodic=True),
                  'hann': lambda n: torch.hann_window(n, periodic=True),
                  'hamming': lambda n: torch.hamming_window(n, periodic=True),
                  'bartlett': lambda n: torch.bartlett_window(n, periodic=True)}
        if window_type not in makers:
            raise ValueError(f"window {window_type!r} not implemented")
        return makers[window_type](H).unsqueeze(-1) * makers[window_type](W).unsqueeze(0)

    def pad_with_new_size(self, img, new_size, mode='constant'):
        """Centre pad to new_size (deblurring.py:368-377)."""
        h, w = img.shape[-2:]
        new_h, new_w = new_size
        padding = [int(np.floor((new_w - w) / 2)), int(np.ceil((new_w - w) / 2)),
                   int(np.floor((new_h - h) / 2)), int(np.ceil((new_h - h) / 2))]
        return torch.nn.functional.pad(img, padding, mode=mode)

    def crop_with_old_size(self, img, old_size):
        """Inverse of pad_with_new_size (deblurring.py:379-394)."""
        h, w = img.shape[-2:]
        old_h, old_w = old_size
        left, right = int(np.floor((w - old_w) / 2)), int(np.ceil((w - old_w) / 2))
        top, bottom = int(np.floor((h - old_h) / 2)), int(np.ceil((h - old_h) / 2))
        return img[..., top:h - bottom, left:w - right]
