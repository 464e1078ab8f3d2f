#!/usr/bin/env python
"""Benchmark of the Polyblur hot path (BASELINE.json metric: Mpix/s end-to-end
polyblur_deblurring, n_iter=3, alpha=6, beta=1).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl cuda|reference]

One "step" = one polyblur_deblurring call over one batch of synthetic images (config 2 of
BASELINE.json: 32 x 3 x 1080 x 1920 float32 per GPU).  Mpix/s = pixels of the batch (counted
once: not x n_iter, not x channels) / time.  With N > 1 (torchrun, one rank per GPU) every
rank owns its own 32 images (weak scaling, no data-path collective) and the time is the max
over ranks.  Rank 0 prints one JSON line.

  value        device-resident inputs, CUDA events around K steps
  e2e          the public API called with pinned HOST tensors: H2D + kernels + D2H per step
  roofline     dominant kernel: algorithmic bytes / CUDA-event duration vs measured HBM peak
  cpu_baseline the ATen-CPU port of the reference path (oracle/polyblur_oracle_torch.py)
               timed on this box's host cores on a 2-image sample (N = 1 only)
  --impl reference   times only that CPU port (rank 0), same metric / unit / config.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "Mpix/s end-to-end polyblur_deblurring n_iter=3"
BYTES_PER_PX_ITER = {"deconvolution": 24.0, "estimate": 12.0}
STEP_BYTES_PER_PX_ITER = 36.0


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--dist", default="mosaic", choices=["mosaic", "white"],
                    help="synthetic distribution of the headline numbers (SURVEY.md 8d)")
    ap.add_argument("--batch", type=int, default=32, help="images per GPU")
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--n-iter", type=int, default=3)
    ap.add_argument("--engine", type=int, default=0, help="0 auto, 1 spatial, 2 fft")
    ap.add_argument("--cpu-sample", type=int, default=2, help="images in the CPU baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the other distribution")
    ap.add_argument("--no-graph", action="store_true", help="time the eager enqueue instead of the CUDA graph replay")
    return ap.parse_args()


def workload_name(a, dist):
    return (f"C2: batch {a.batch} synthetic {a.width}x{a.height} RGB float32 per GPU, "
            f"{'blurred-mosaic (M)' if dist == 'mosaic' else 'white-noise (W)'} distribution of SURVEY 8d, "
            f"n_iter={a.n_iter} alpha=6 beta=1")


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


def cpu_port_throughput(a, dist, steps, warmup):
    """Mpix/s of the ATen-CPU port of the reference path on a bounded sample (rank 0)."""
    import torch
    from oracle import polyblur_oracle_torch as pt
    from polyblur_b200 import synthetic
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    n = max(1, a.cpu_sample)
    x = synthetic.make(dist, n, 3, a.height, a.width, device="cpu")
    for _ in range(warmup):
        pt.polyblur_deblurring(x, n_iter=a.n_iter, alpha=6, beta=1)
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        pt.polyblur_deblurring(x, n_iter=a.n_iter, alpha=6, beta=1)
        times.append(time.perf_counter() - t0)
    dt = sum(times) / len(times)
    mpix = n * a.height * a.width / 1e6
    sample = (f"{n} images of the same workload ({a.width}x{a.height} RGB, {dist}), n_iter={a.n_iter}, "
              f"{steps} timed call(s) after {warmup} warm-up, torch {torch.__version__} CPU, "
              f"{torch.get_num_threads()} threads")
    return mpix / dt, dt, cores, sample


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return None
    val, dt, cores, sample = cpu_port_throughput(a, a.dist, max(1, a.steps), max(0, a.warmup))
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "Mpix/s", "n_gpus": a.gpus,
        "steps": a.steps, "warmup": a.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32 (c64 FFT)", "data": "synthetic",
        "config": {"workload": workload_name(a, a.dist), "sample": sample},
        "cpu_baseline": {"value": val, "unit": "Mpix/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "Mpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    return line


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

    B, H, W = a.batch, a.height, a.width
    first, _ = sharding.shard_range(B * world, rank, world)     # weak scaling: B images per rank
    pix_all = B * world * H * W
    peak, peak_src = measured_peaks()

    def params():
        return deblurring._make_params(a.n_iter, 0.352, 0.768, 6, 1, 0.8, 2.0, 25, 0.0, False, False, False,
                                       False, engine=a.engine)

    def measure(kind, with_profile):
        """-> x, out, ms per step (K timed steps, no profiling inside), per-kernel profile, clocks.

        The headline timing replays the step as one CUDA graph (the enqueue has no host sync) unless
        --no-graph.  The per-kernel CUDA events of the library's profiler cost ~5 % (two event records
        per launch), so the roofline pass is a second region of K eager steps right after."""
        x = synthetic.make(kind, B, 3, H, W, first_index=first, device=dev)
        out = torch.empty_like(x)
        p = params()
        graphed = None
        if not a.no_graph:
            try:
                graphed = deblurring.GraphedPolyblur((B, 3, H, W), device=dev, n_iter=a.n_iter, alpha=6, beta=1,
                                                     engine=a.engine)
                graphed.x.copy_(x)
            except Exception as exc:          # capture not possible on this stack: time the eager enqueue
                print(f"CUDA graph capture failed ({exc}); timing the eager enqueue", file=sys.stderr)
                graphed = None

        def step_eager():
            deblurring.polyblur_device(x, p, out=out)

        step = graphed if graphed is not None else step_eager
        # clocks / throttle reasons are sampled from the warm-up through the timed region to the end of
        # the profiled pass (all under the same load; the timed region alone lasts ~0.1 s = 1 sample)
        clocks = None
        sampler = ClockSampler(local) if (rank == 0 and with_profile) else None
        if sampler:
            sampler.start()
        for _ in range(a.warmup):
            step()
        ms = time_steps(step, a.steps, barrier)
        ms = max_over_ranks(ms)
        prof = {}
        if with_profile:
            step_eager()
            _lib.profile_begin()
            time_steps(step_eager, a.steps, barrier)
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
        return x, (graphed.out if graphed is not None else out), ms, prof, clocks, mode

    x, out, ms, prof, clocks, mode = measure(a.dist, True)
    value = pix_all / 1e6 / (ms / 1e3)

    # ---- roofline of the dominant kernel group (CUDA events recorded by the library around each launch)
    # One "launch" of a group = the kernels one Polyblur iteration runs for it over the whole batch:
    #   estimate      = k_rows2 + k_cols2 + k_params                        12 B/px/iter algorithmic
    #   deconvolution = the engines (narrow / tiled / FFT passes; every image goes through exactly one,
    #                   chosen on the device, so their times add up to one pass over the batch)  24 B/px/iter
    EST = ("k_cols2", "k_rows2", "k_params")      # (k_cols / k_rows of estimate.cu for lengths with a prime factor > 13)
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
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as f:
            traffic = json.load(f).get(f"{dom}:{a.dist}:{B}x{H}x{W}")
    except Exception:
        pass
    step_bytes = STEP_BYTES_PER_PX_ITER * a.n_iter * B * H * W
    total_prof_ms = sum(v[0] for v in prof.values())
    roofline = {
        "bound": "hbm", "kernel": dom + " (" + " + ".join(members) + ")", "achieved": achieved, "peak": peak,
        "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
        "algorithmic_bytes_per_launch": alg_bytes, "avg_launch_ms": avg_ms,
        "launch_definition": "one Polyblur iteration of this kernel group over the whole batch",
        "timed_with": "per-launch CUDA events of the library's profiler over a second region of K eager steps",
        "kernel_share_of_step": dom_ms / total_prof_ms if total_prof_ms else None,
        "step": {"algorithmic_bytes": step_bytes, "achieved": step_bytes / (ms / 1e3) / 1e9,
                 "frac": step_bytes / (ms / 1e3) / 1e9 / peak},
        "kernels_ms_per_step": {k: v[0] / a.steps for k, v in prof.items()},
    }
    gpu_launches = sum(v[1] for v in prof.values())

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
        del xh
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
                           "api": "polyblur_b200.io.deblur_uint8(pinned uint8 (B,H,W,C)) -> uint8 CPU tensor"}
        del xu

    # ---- the other synthetic distribution, same measurement (kernel-only) ----------------------
    secondary = None
    if not a.no_secondary:
        other = "white" if a.dist == "mosaic" else "mosaic"
        del x, out
        torch.cuda.empty_cache()
        _, _, ms2, _, _, _ = measure(other, False)
        secondary = {"workload": workload_name(a, other), "value": pix_all / 1e6 / (ms2 / 1e3),
                     "unit": "Mpix/s", "ms_per_step": ms2,
                     "step_roofline_frac": step_bytes / (ms2 / 1e3) / 1e9 / peak}

    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        val, dt, cores, sample = cpu_port_throughput(a, a.dist, 1, 1)
        cpu = {"value": val, "unit": "Mpix/s", "cores": cores, "kind": "port", "sample": sample}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "Mpix/s", "n_gpus": world, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(a, a.dist), "global_batch": B * world,
                       "parallelism": f"batch-sharded x{world}, no data-path collective",
                       "l2": f"inputs ({B * 3 * H * W * 4 / 1e6:.0f} MB per GPU) and every intermediate are larger than the 126 MB L2; no flush needed",
                       "engine": {0: "auto", 1: "spatial", 2: "fft"}[a.engine], "launch": mode},
            "e2e": e2e, "gpu_launches": gpu_launches, "clocks": clocks, "roofline": roofline,
            "cpu_baseline": cpu, "secondary": secondary,
        }
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
