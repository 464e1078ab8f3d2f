#!/usr/bin/env python
"""Benchmark of the Polyblur hot path (BASELINE.json metric: Mpix/s end-to-end
polyblur_deblurring, n_iter=3, alpha=6, beta=1).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl cuda|reference] [--config C2|C3|C4|C5]

One "step" = one polyblur_deblurring call over one batch of synthetic images.  Mpix/s = pixels of the
batch (counted once: not x n_iter, not x channels) / time.  The named configurations of BASELINE.json:

    C2  32 x 3 x 1080 x 1920 on one GPU, n_iter = 3            (default for --gpus 1: the headline)
    C3  256 x 3 x 2160 x 3840 sharded per image over 8 GPUs = 32 images per GPU, n_iter = 3
                                                               (default for --gpus N > 1, weak scaling)
    C4  one 3 x 9000 x 12000 image, n_iter = 5, one GPU
    C5  64 x 3 x 2160 x 3840 over 8 GPUs = 8 per GPU, n_iter sweep 1..10, with and without the
        domain-transform (RF) prefilter; value = the n_iter = 3 point without the prefilter

With N > 1 (torchrun, one rank per GPU) every rank owns its own images (no data-path collective) and the
time is the max over ranks.  Rank 0 prints one JSON line.

  value        device-resident inputs, CUDA events around K steps (CUDA-graph replay of the enqueue)
  eager_api    the same through polyblur_b200.polyblur_deblurring(CUDA tensor), no graph
  e2e          the public API called with pinned HOST tensors: H2D + kernels + D2H per step, with the
               host link each rank sees while all ranks copy (link_probe)
  roofline     dominant kernel group: algorithmic bytes / CUDA-event duration vs measured HBM peak, the unit
               that binds it according to the committed ncu capture (profiles/roofline_pipes.json)
  cpu_baseline the unmodified reference (baseline/_ref, method='fft') timed on this box's host cores on a
               bounded sample (N = 1 only); falls back to the ATen-CPU port (oracle/) if it is not installed
  secondary    the other synthetic distribution; the reference on this GPU through torch CUDA (library_gpu);
               at N = 1 the C3 per-GPU shape, so that a scaling run has a same-config one-GPU number
  --impl reference   times only the CPU reference (rank 0), same metric / unit / config.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time
import types

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "Mpix/s end-to-end polyblur_deblurring n_iter=3"
BYTES_PER_PX_ITER = {"deconvolution": 24.0, "estimate": 12.0}
STEP_BYTES_PER_PX_ITER = 36.0
RF_BYTES_PER_PX_ITER = 96.0                                      # with the RF prefilter

CONFIGS = {
    "C2": dict(batch=32, height=1080, width=1920, n_iter=3,
               desc="BASELINE configs[1]: batch 32 synthetic 1920x1080 RGB, n_iter=3 alpha=6 beta=1, 1xB200"),
    "C3": dict(batch=32, height=2160, width=3840, n_iter=3,
               desc="BASELINE configs[2]: batch 256 synthetic 3840x2160 RGB, n_iter=3, per-image sharded across "
                    "8xB200 = 32 images per GPU"),
    "C4": dict(batch=1, height=9000, width=12000, n_iter=5,
               desc="BASELINE configs[3]: single 12000x9000 RGB, n_iter=5, 1xB200 (one GPU holds the whole image: "
                    "no tiling or halo exchange needed)"),
    "C5": dict(batch=8, height=2160, width=3840, n_iter=3,
               desc="BASELINE configs[4]: n_iter sweep 1-10, batch 64 at 3840x2160 over 8xB200 = 8 images per GPU, "
                    "with and without the domain-transform prefilter; value = n_iter 3 without it"),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--config", default=None, choices=sorted(CONFIGS),
                    help="named BASELINE configuration (default: C2 on one GPU, C3's per-GPU shape on several)")
    ap.add_argument("--dist", default="mosaic", choices=["mosaic", "white"],
                    help="synthetic distribution of the headline numbers (SURVEY.md 8d)")
    ap.add_argument("--batch", type=int, default=None, help="images per GPU (overrides the configuration)")
    ap.add_argument("--height", type=int, default=None)
    ap.add_argument("--width", type=int, default=None)
    ap.add_argument("--n-iter", type=int, default=None)
    ap.add_argument("--engine", type=int, default=0, help="0 auto, 1 spatial, 2 fft")
    ap.add_argument("--cpu-sample", type=int, default=2, help="images in the CPU baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the other distribution / library legs")
    ap.add_argument("--no-graph", action="store_true", help="time the eager enqueue instead of the CUDA graph replay")
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if a.config is None:
        a.config = "C2" if max(a.gpus, world) == 1 else "C3"
    cfg = CONFIGS[a.config]
    a.custom = any(v is not None for v in (a.batch, a.height, a.width, a.n_iter))
    a.batch = cfg["batch"] if a.batch is None else a.batch
    a.height = cfg["height"] if a.height is None else a.height
    a.width = cfg["width"] if a.width is None else a.width
    a.n_iter = cfg["n_iter"] if a.n_iter is None else a.n_iter
    return a


def workload_name(a, dist, cfg=None, batch=None, h=None, w=None, n_iter=None):
    cfg = cfg or a.config
    batch, h, w, n_iter = batch or a.batch, h or a.height, w or a.width, n_iter or a.n_iter
    tag = cfg + (" (custom shape)" if a.custom and cfg == a.config else "")
    return (f"{tag}: batch {batch} synthetic {w}x{h} RGB float32 per GPU, "
            f"{'blurred-mosaic (M)' if dist == 'mosaic' else 'white-noise (W)'} distribution of SURVEY 8d, "
            f"n_iter={n_iter} alpha=6 beta=1")


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [t.strip() for t in line.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                "samples": len(sm)}


# ---- the reference itself (baseline/_ref: `pip install --target` of /root/reference, see baseline/install_ref.sh) ----
def import_reference():
    """-> the UNMODIFIED reference package, or None.  The only shim is a stub for `skimage`, which the reference
    imports for img_as_float32 (utils.py:3) and which is not in this image; tensors never touch it."""
    ref_dir = os.path.join(ROOT, "baseline", "_ref")
    if not os.path.isdir(os.path.join(ref_dir, "polyblur")):
        return None
    if "skimage" not in sys.modules:
        try:
            import skimage  # noqa: F401
        except Exception:
            sk = types.ModuleType("skimage")
            sk.img_as_float32 = lambda x: x
            sys.modules["skimage"] = sk
    if ref_dir not in sys.path:
        sys.path.insert(0, ref_dir)
    try:
        import polyblur
        return polyblur
    except Exception as exc:       # noqa: BLE001
        print(f"reference import failed: {exc}", file=sys.stderr)
        return None


def cpu_sample_shape(a):
    """Bounded sample of the workload for the CPU legs (about 10-30 s of CPU work)."""
    if a.height * a.width > 40e6:          # C4: a quarter-size crop of the single image
        return 1, a.height // 2, a.width // 2, "one image of half the side lengths (a quarter of the pixels)"
    n = max(1, min(a.cpu_sample, a.batch))
    return n, a.height, a.width, f"{n} image(s) of the batch"


def cpu_reference_throughput(a, dist, steps, warmup):
    """Mpix/s of the reference's own CPU implementation of the path on a bounded sample (rank 0): the unmodified
    reference when baseline/_ref holds it (kind "reference"), else the ATen-CPU port under oracle/ (kind "port")."""
    import torch
    from polyblur_b200 import synthetic
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    n, h, w, what = cpu_sample_shape(a)
    x = synthetic.make(dist, n, 3, h, w, device="cpu")
    ref = import_reference()
    if ref is not None:
        kind = "reference"

        def call():
            with torch.no_grad():
                return ref.polyblur_deblurring(x, n_iter=a.n_iter, alpha=6, beta=1, method="fft")
        impl = "unmodified teboli/polyblur from baseline/_ref, polyblur.polyblur_deblurring(CPU tensor, method='fft')"
    else:
        from oracle import polyblur_oracle_torch as pt
        kind = "port"

        def call():
            return pt.polyblur_deblurring(x, n_iter=a.n_iter, alpha=6, beta=1)
        impl = "ATen-CPU port of the reference path (oracle/polyblur_oracle_torch.py); baseline/_ref not installed"
    for _ in range(warmup):
        call()
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        call()
        times.append(time.perf_counter() - t0)
    dt = sum(times) / len(times)
    mpix = n * h * w / 1e6
    sample = (f"{what}: {n} x 3 x {h} x {w} ({dist}), n_iter={a.n_iter}, {steps} timed call(s) after {warmup} "
              f"warm-up; {impl}; torch {torch.__version__} CPU, {torch.get_num_threads()} threads")
    return mpix / dt, dt, cores, sample, kind


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return None
    val, dt, cores, sample, kind = cpu_reference_throughput(a, a.dist, max(1, a.steps), max(0, a.warmup))
    return {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "Mpix/s", "n_gpus": a.gpus,
        "steps": a.steps, "warmup": a.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32 (c64 FFT)", "data": "synthetic",
        "config": {"workload": workload_name(a, a.dist), "sample": sample},
        "cpu_baseline": {"value": val, "unit": "Mpix/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": val, "unit": "Mpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }


def time_steps(fn, steps, barrier):
    import torch
    barrier()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    barrier()
    return e0.elapsed_time(e1) / steps


def load_json(name):
    try:
        with open(os.path.join(ROOT, "profiles", name)) as f:
            return json.load(f)
    except Exception:
        return {}


def binding_unit(members_ms, key_suffix):
    """The unit that binds the kernel group according to the committed ncu capture of this configuration
    (tools/ncu_summary.py --pipes -> profiles/roofline_pipes.json): time-weighted utilisation of DRAM, the
    L1/shared-memory pipe, the FMA pipe and the issue slots over the group's kernels; bound = the largest."""
    pipes = load_json("roofline_pipes.json")
    acc, tot, src = {}, 0.0, set()
    for name, ms in members_ms.items():
        # the profiler's class names are generation-neutral; the captured kernels of the current build may carry a "2"
        def lookup(sfx):
            for gen in ("3", "2", ""):
                if f"{name}{gen}:{sfx}" in pipes:
                    return pipes[f"{name}{gen}:{sfx}"]
            # the same image size at another batch size: the per-kernel utilisation does not depend on the batch once
            # the grid fills the GPU (the 4K capture is taken at 8 images, the C3 line runs 32)
            dist_, _, shape_ = sfx.partition(":")
            hw = shape_.split("x", 1)[1] if "x" in shape_ else None
            for gen in ("3", "2", ""):
                for k, v in pipes.items():
                    kn, kd, ks = (k.split(":") + ["", ""])[:3]
                    if kn == f"{name}{gen}" and kd == dist_ and hw and ks.split("x", 1)[-1] == hw:
                        v = dict(v)
                        v["source"] = v.get("source", "?") + f" (captured at {ks})"
                        return v
            return None
        rec = lookup(key_suffix)
        if not rec:                  # the estimate kernels do the same work for both synthetic distributions
            alt = key_suffix.replace("white:", "mosaic:") if key_suffix.startswith("white:") else key_suffix.replace("mosaic:", "white:")
            rec = lookup(alt)
        if not rec or ms <= 0:
            continue
        tot += ms
        src.add(rec.get("source", "?"))
        for k in ("dram_pct", "l1tex_pct", "fma_pipe_pct", "issue_pct", "l2_pct", "warps_pct"):
            if rec.get(k) is not None:
                acc[k] = acc.get(k, 0.0) + rec[k] * ms
    if not tot:
        return "unknown (no ncu capture of this configuration is committed)", None
    util = {k: round(v / tot, 1) for k, v in acc.items()}
    label = {"dram_pct": "hbm", "l1tex_pct": "l1tex (shared-memory / L1 pipe)", "fma_pipe_pct": "fp32 fma pipe",
             "issue_pct": "instruction issue"}
    cand = {k: util[k] for k in label if k in util}
    top = max(cand, key=lambda k: cand[k])
    util["source"] = "profiles/roofline_pipes.json <- " + ", ".join(sorted(src))
    return label[top], util


def run_cuda(a):
    import torch
    import torch.distributed as dist
    import polyblur_b200
    from polyblur_b200 import _lib, deblurring, sharding, synthetic

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl cuda needs a GPU (there is no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    # one process per GPU: keep it (and the pinned staging buffers it allocates) on the GPU's NUMA node, so that the
    # host <-> device copies of the N ranks of a box do not cross the socket link (BENCH_NUMA=0 turns it off)
    numa = sharding.bind_to_gpu_numa_node(local) if (world > 1 and os.environ.get("BENCH_NUMA", "1") != "0") else None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()

    def max_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def gather_floats(vals):
        t = torch.tensor(vals, dtype=torch.float64, device=dev)
        if world == 1:
            return [t.tolist()]
        parts = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(parts, t)
        return [p.tolist() for p in parts]

    B, H, W = a.batch, a.height, a.width
    first, _ = sharding.shard_range(B * world, rank, world)     # weak scaling: B images per rank
    peak, peak_src = measured_peaks()

    def params(n_iter, rf=False):
        return deblurring._make_params(n_iter, 0.352, 0.768, 6, 1, 0.8, 2.0, 25, 0.0, False, False, rf, False,
                                       engine=a.engine, prefilter="rf" if rf else "bilateral")

    def measure(kind, with_profile, b=B, h=H, w=W, n_iter=a.n_iter, steps=a.steps, rf=False, eager_too=False):
        """-> dict(ms per step of K timed steps, per-kernel profile, clocks, launch mode, eager ms).

        The headline timing replays the step as one CUDA graph (the enqueue has no host sync) unless
        --no-graph.  The per-kernel CUDA events of the library's profiler cost ~5 % (two event records
        per launch), so the roofline pass is a second region of K eager steps right after."""
        x = synthetic.make(kind, b, 3, h, w, first_index=first if b == B else 0, device=dev)
        out = torch.empty_like(x)
        p = params(n_iter, rf)
        graphed = None
        if not a.no_graph:
            try:
                graphed = deblurring.GraphedPolyblur((b, 3, h, w), device=dev, n_iter=n_iter, alpha=6, beta=1,
                                                     engine=a.engine, prefiltering=rf, prefilter="rf" if rf else "bilateral")
                graphed.x.copy_(x)
            except Exception as exc:          # capture not possible on this stack: time the eager enqueue
                print(f"CUDA graph capture failed ({exc}); timing the eager enqueue", file=sys.stderr)
                graphed = None

        def step_eager():
            deblurring.polyblur_device(x, p, out=out)

        def step_api():
            polyblur_b200.polyblur_deblurring(x, n_iter=n_iter, alpha=6, beta=1, engine=a.engine, prefiltering=rf,
                                              prefilter="rf" if rf else "bilateral")

        step = graphed if graphed is not None else step_eager
        # clocks / throttle reasons are sampled from the warm-up through the timed region to the end of
        # the profiled pass (all under the same load; the timed region alone lasts ~0.1 s = 1 sample)
        clocks = None
        sampler = ClockSampler(local) if (rank == 0 and with_profile) else None
        if sampler:
            sampler.start()
        for _ in range(a.warmup):
            step()
        ms = max_over_ranks(time_steps(step, steps, barrier))
        prof, ms_api = {}, None
        if eager_too:
            for _ in range(2):
                step_api()
            ms_api = max_over_ranks(time_steps(step_api, steps, barrier))
        if with_profile:
            step_eager()
            _lib.profile_begin()
            time_steps(step_eager, steps, barrier)
            prof = _lib.profile_end()
            if sampler:
                # keep the load on for a few more sampling periods
                t_end = time.time() + 0.6
                while time.time() < t_end:
                    step()
                    torch.cuda.synchronize()
        if sampler:
            clocks = sampler.stop()
        mode = "cuda-graph replay" if graphed is not None else "eager enqueue"
        res = dict(x=x, ms=ms, prof=prof, clocks=clocks, mode=mode, ms_api=ms_api)
        del graphed, out
        return res

    m = measure(a.dist, True, eager_too=True)
    x, ms, prof, clocks, mode = m["x"], m["ms"], m["prof"], m["clocks"], m["mode"]
    pix_all = B * world * H * W
    value = pix_all / 1e6 / (ms / 1e3)
    eager_api = {"value": pix_all / 1e6 / (m["ms_api"] / 1e3), "unit": "Mpix/s", "ms_per_step": m["ms_api"],
                 "api": "polyblur_b200.polyblur_deblurring(CUDA tensor) -> CUDA tensor, eager enqueue (no graph)"}

    # ---- roofline of the dominant kernel group (CUDA events recorded by the library around each launch)
    # One "launch" of a group = the kernels one Polyblur iteration runs for it over the whole batch:
    #   estimate      = k_rows + k_cols + k_params (k_rows3 / k_cols3 of csrc/estimate3.cu for the named shapes)   12 B/px/iter algorithmic
    #   deconvolution = the engines (narrow / tiled / FFT passes; every image goes through exactly one,
    #                   chosen on the device, so their times add up to one pass over the batch)  24 B/px/iter
    EST = ("k_cols", "k_rows", "k_params")        # the profiler's class names are generation-neutral
    DEC = ("k_deconv_narrow", "k_deconv_spatial", "k_fft_rows_fwd", "k_fft_cols", "k_fft_rows_inv")
    groups = {"estimate": sum(prof.get(k, (0.0, 0))[0] for k in EST),
              "deconvolution": sum(prof.get(k, (0.0, 0))[0] for k in DEC)}
    dom = max(groups, key=lambda k: groups[k])
    dom_ms = groups[dom]
    members = [k for k in (EST if dom == "estimate" else DEC) if prof.get(k, (0.0, 0))[0] > 0.05 * dom_ms]
    iters = max(1, a.steps * a.n_iter)
    avg_ms = dom_ms / iters
    alg_bytes = BYTES_PER_PX_ITER[dom] * B * H * W
    achieved = alg_bytes / (avg_ms / 1e3) / 1e9 if avg_ms > 0 else 0.0
    key_suffix = f"{a.dist}:{B}x{H}x{W}"
    traffic = load_json("roofline_traffic.json").get(f"{dom}:{key_suffix}")
    bound, util = binding_unit({k: prof[k][0] for k in members}, key_suffix)
    step_bytes = STEP_BYTES_PER_PX_ITER * a.n_iter * B * H * W
    total_prof_ms = sum(v[0] for v in prof.values())
    roofline = {
        "bound": bound, "kernel": dom + " (" + " + ".join(members) + ")", "achieved": achieved, "peak": peak,
        "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
        "pipe_utilisation_pct": util,
        "algorithmic_bytes_per_launch": alg_bytes, "avg_launch_ms": avg_ms,
        "launch_definition": "one Polyblur iteration of this kernel group over the whole batch",
        "timed_with": "per-launch CUDA events of the library's profiler over a second region of K eager steps",
        "kernel_share_of_step": dom_ms / total_prof_ms if total_prof_ms else None,
        "step": {"algorithmic_bytes": step_bytes, "achieved": step_bytes / (ms / 1e3) / 1e9,
                 "frac": step_bytes / (ms / 1e3) / 1e9 / peak},
        "kernels_ms_per_step": {k: v[0] / a.steps for k, v in prof.items()},
        "groups_ms_per_iteration": {k: v / iters for k, v in groups.items()},
    }
    gpu_launches = sum(v[1] for v in prof.values())

    # ---- C5: the n_iter scan, with and without the domain-transform prefilter (kernel-only) ----------------
    sweep = None
    if a.config == "C5" and not a.custom:
        sweep = []
        ks = max(2, min(a.steps, 3))
        for rf in (False, True):
            for n in range(1, 11):
                r = measure(a.dist, False, n_iter=n, steps=ks, rf=rf)
                bpp = (RF_BYTES_PER_PX_ITER if rf else STEP_BYTES_PER_PX_ITER) * n * B * H * W
                sweep.append({"n_iter": n, "rf_prefilter": rf, "ms_per_step": r["ms"],
                              "value": pix_all / 1e6 / (r["ms"] / 1e3),
                              "step_roofline_frac": bpp / (r["ms"] / 1e3) / 1e9 / peak})
                del r

    # ---- end to end through the public API with pinned host tensors -----------------------------
    e2e = None
    if not a.no_e2e:
        xh = torch.empty(x.shape, dtype=torch.float32, pin_memory=True)
        xh.copy_(x)
        torch.cuda.synchronize()

        def step_e2e():
            y = polyblur_b200.polyblur_deblurring(xh, n_iter=a.n_iter, alpha=6, beta=1, engine=a.engine)
            assert y.device.type == "cpu"

        for _ in range(2):
            step_e2e()
        k_e2e = max(2, min(a.steps, 5))
        ms_e2e = max_over_ranks(time_steps(step_e2e, k_e2e, barrier))
        nbytes = x.numel() * 4
        e2e = {"value": pix_all / 1e6 / (ms_e2e / 1e3), "unit": "Mpix/s", "h2d_bytes_per_step": nbytes,
               "d2h_bytes_per_step": nbytes, "ms_per_step": ms_e2e, "steps": k_e2e,
               "api": "polyblur_b200.polyblur_deblurring(pinned CPU tensor) -> CPU tensor"}
        # the host link every rank sees while ALL ranks copy in both directions at once: the ceiling of e2e
        probe_n = min(nbytes, 1 << 30)
        src = xh.view(-1)[: probe_n // 4]
        dstd = torch.empty(probe_n // 4, dtype=torch.float32, device=dev)
        srcd = torch.empty_like(dstd)
        dsth = torch.empty(probe_n // 4, dtype=torch.float32, pin_memory=True)
        s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
        rates = []
        for it in range(3):
            barrier()
            torch.cuda.synchronize()
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
            with torch.cuda.stream(s1):
                ev[0].record()
                dstd.copy_(src, non_blocking=True)
                ev[1].record()
            with torch.cuda.stream(s2):
                ev[2].record()
                dsth.copy_(srcd, non_blocking=True)
                ev[3].record()
            torch.cuda.synchronize()
            rates.append((probe_n / ev[0].elapsed_time(ev[1]) / 1e6, probe_n / ev[2].elapsed_time(ev[3]) / 1e6))
        best = [max(r[0] for r in rates), max(r[1] for r in rates), float(numa["node"]) if numa else -1.0]
        per_rank = gather_floats(best)
        h2d = [round(r[0], 2) for r in per_rank]
        d2h = [round(r[1], 2) for r in per_rank]
        numa_nodes = [int(r[2]) for r in per_rank]
        ceil_ms = max(nbytes / (min(h2d) * 1e6), nbytes / (min(d2h) * 1e6))
        e2e["link_probe"] = {"h2d_gbs_per_rank": h2d, "d2h_gbs_per_rank": d2h, "bytes": probe_n,
                             "how": "pinned host <-> device copies of this size in both directions at once on every "
                                    "rank simultaneously, best of 3, CUDA events",
                             "numa_node_per_rank": numa_nodes if world > 1 else None,   # -1: process not bound
                             "link_bound_ms_per_step": ceil_ms,
                             "link_bound_value": pix_all / 1e6 / (ceil_ms / 1e3),
                             "e2e_over_link_bound": (pix_all / 1e6 / (ms_e2e / 1e3)) / (pix_all / 1e6 / (ceil_ms / 1e3))}
        del xh, src, dstd, srcd, dsth
        # the same images as 8-bit HWC host buffers through polyblur_b200.io.deblur_uint8 (conversions on
        # the device: 1 byte per sample over PCIe instead of 4) -- informational, not the contract's e2e
        from polyblur_b200 import io as pbio
        xu = (x.permute(0, 2, 3, 1) * 255).round().to(torch.uint8).contiguous().cpu().pin_memory()

        def step_u8():
            y = pbio.deblur_uint8(xu, n_iter=a.n_iter, alpha=6, beta=1, engine=a.engine)
            assert y.device.type == "cpu" and y.dtype == torch.uint8

        for _ in range(2):
            step_u8()
        ms_u8 = max_over_ranks(time_steps(step_u8, k_e2e, barrier))
        e2e["uint8_io"] = {"value": pix_all / 1e6 / (ms_u8 / 1e3), "unit": "Mpix/s", "ms_per_step": ms_u8,
                           "h2d_bytes_per_step": xu.numel(), "d2h_bytes_per_step": xu.numel(),
                           "vs_device_resident": ms / ms_u8,
                           "api": "polyblur_b200.io.deblur_uint8(pinned uint8 (B,H,W,C)) -> uint8 CPU tensor"}
        del xu

    # ---- secondary legs (kernel-only) ----------------------------------------------------------------------
    secondary = None
    if not a.no_secondary:
        other = "white" if a.dist == "mosaic" else "mosaic"
        x_keep = x[: min(B, 8)].clone() if H * W <= 2160 * 3840 else None
        del x, m
        torch.cuda.empty_cache()
        r2 = measure(other, False)
        secondary = {"workload": workload_name(a, other), "value": pix_all / 1e6 / (r2["ms"] / 1e3),
                     "unit": "Mpix/s", "ms_per_step": r2["ms"],
                     "step_roofline_frac": step_bytes / (r2["ms"] / 1e3) / 1e9 / peak}
        del r2
        torch.cuda.empty_cache()
        if world == 1 and a.config == "C2" and not a.custom:
            # the per-GPU shape the multi-GPU runs use (C3), on this one GPU: the same-config reference point
            c3 = CONFIGS["C3"]
            r3 = measure(a.dist, False, b=c3["batch"], h=c3["height"], w=c3["width"], n_iter=c3["n_iter"],
                         steps=max(2, min(a.steps, 5)))
            pix3 = c3["batch"] * c3["height"] * c3["width"]
            secondary["c3_shape_one_gpu"] = {
                "workload": workload_name(a, a.dist, "C3", c3["batch"], c3["height"], c3["width"], c3["n_iter"]),
                "value": pix3 / 1e6 / (r3["ms"] / 1e3), "unit": "Mpix/s", "ms_per_step": r3["ms"],
                "step_roofline_frac": STEP_BYTES_PER_PX_ITER * c3["n_iter"] * pix3 / (r3["ms"] / 1e3) / 1e9 / peak}
            del r3
            torch.cuda.empty_cache()
        # "library-call GPU": the unmodified reference on this GPU through torch CUDA (method='fft': cuFFT + ATen)
        if rank == 0 and x_keep is not None:
            ref = import_reference()
            if ref is None:
                secondary["library_gpu"] = {"unavailable": "baseline/_ref is not installed"}
            else:
                try:
                    def step_lib():
                        with torch.no_grad():
                            return ref.polyblur_deblurring(x_keep, n_iter=a.n_iter, alpha=6, beta=1, method="fft")
                    for _ in range(2):
                        y = step_lib()
                    assert y.is_cuda
                    kl = max(2, min(a.steps, 5))
                    torch.cuda.synchronize()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    for _ in range(kl):
                        step_lib()
                    e1.record()
                    torch.cuda.synchronize()
                    ms_lib = e0.elapsed_time(e1) / kl
                    nl = x_keep.shape[0]
                    secondary["library_gpu"] = {
                        "value": nl * H * W / 1e6 / (ms_lib / 1e3), "unit": "Mpix/s", "ms_per_step": ms_lib,
                        "sample": f"{nl} images of the batch ({a.dist}), device resident, {kl} timed calls",
                        "impl": "unmodified teboli/polyblur (baseline/_ref) polyblur_deblurring(CUDA tensor, "
                                "method='fft') = torch.fft / cuFFT + ATen kernels on this B200",
                        "peak_mem_gb": torch.cuda.max_memory_allocated(dev) / 1e9}
                    del y
                except Exception as exc:       # noqa: BLE001
                    secondary["library_gpu"] = {"unavailable": f"{type(exc).__name__}: {exc}"[:300]}
        del x_keep
        torch.cuda.empty_cache()

    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        val, dt, cores, sample, kind = cpu_reference_throughput(a, a.dist, 1, 1)
        cpu = {"value": val, "unit": "Mpix/s", "cores": cores, "kind": kind, "sample": sample}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "Mpix/s", "n_gpus": world, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(a, a.dist), "name": a.config, "baseline_config": CONFIGS[a.config]["desc"],
                       "global_batch": B * world,
                       "parallelism": f"batch-sharded x{world}, no data-path collective",
                       "l2": f"inputs ({B * 3 * H * W * 4 / 1e6:.0f} MB per GPU) and every intermediate are larger than the 126 MB L2; no flush needed",
                       "engine": {0: "auto", 1: "spatial", 2: "fft"}[a.engine], "launch": mode},
            "eager_api": eager_api,
            "e2e": e2e, "gpu_launches": gpu_launches, "clocks": clocks, "roofline": roofline,
            "cpu_baseline": cpu, "secondary": secondary,
        }
        if world > 1 and a.config == "C3":
            # the N = 1 headline is C2 (BASELINE's metric config); the same-workload one-GPU number a weak-scaling
            # efficiency divides by is that line's secondary.c3_shape_one_gpu.value
            line["config"]["weak_scaling_reference"] = ("the N=1 line's secondary.c3_shape_one_gpu.value (32 x 4K on one "
                                                        "GPU); its headline value is C2, another workload")
        if sweep is not None:
            line["sweep"] = sweep
    else:
        line = None
    if world > 1:
        dist.destroy_process_group()
    return line


class StdoutToStderr:
    """Everything written to fd 1 while active (NCCL's version banner, library chatter) goes to
    stderr, so that stdout carries exactly one line: the JSON result."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)
        return self

    def __exit__(self, *exc):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)
        return False


def main():
    a = parse()
    with StdoutToStderr():
        line = run_reference(a) if a.impl == "reference" else run_cuda(a)
    if line is not None:
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
