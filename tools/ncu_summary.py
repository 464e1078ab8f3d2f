#!/usr/bin/env python
"""Condenses an Nsight Compute report (.ncu-rep, read with `ncu -i ... --page raw --csv`) into the
per-launch table committed under profiles/ and, optionally, the DRAM traffic per kernel group that
bench.py reports as roofline.traffic.

    python tools/ncu_summary.py gpurun_out/x.ncu-rep profiles/r02_x.md [--traffic KEY profiles/roofline_traffic.json]
                                [--pipes KEY profiles/roofline_pipes.json] [--csv profiles/r02_x.csv]

--pipes records, per kernel, the utilisation of the units that can bind it (DRAM, L1/shared-memory pipe, FMA pipe,
issue slots) -- bench.py derives ``roofline.bound`` from it; --csv keeps the raw metric rows those numbers come from.
"""
import csv
import json
import subprocess
import sys

COLS = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex%"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
        ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma%"),
        ("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "alu%"),
        ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu%"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2%"),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem_wavefronts"),
        ("derived__memory_l1_wavefronts_shared_excessive", "smem_excess"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps%"),
        ("launch__registers_per_thread", "regs"), ("launch__shared_mem_per_block_dynamic", "smem/CTA"),
        ("launch__grid_size", "grid"), ("smsp__inst_executed.sum", "warp_inst")]
GROUPS = {"estimate": ("k_rows", "k_cols", "k_params"),
          "deconvolution": ("k_deconv_narrow", "k_deconv_spatial", "k_fft_rows_fwd", "k_fft_cols", "k_fft_rows_inv")}


def to_bytes(v, unit):
    m = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    return float(v.replace(",", "")) * m.get(unit, 1)


def main():
    rep, out = sys.argv[1], sys.argv[2]
    if rep.endswith(".csv"):         # a raw page exported on the GPU box: ncu -i X.ncu-rep --page raw --csv > X.raw.csv
        raw = open(rep).read()
    else:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, body = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    lines = ["| # | kernel | " + " | ".join(n for _, n in COLS) + " |", "|---|---|" + "---|" * len(COLS)]
    traffic = {}
    for n, r in enumerate(body):
        name = r[idx["Kernel Name"]].split("(")[0].replace("void ", "").replace("pb::", "")
        cells = []
        for key, _ in COLS:
            if key not in idx:
                cells.append("-")
                continue
            v, u = r[idx[key]], units[idx[key]]
            try:
                f = float(v.replace(",", ""))
                cells.append(f"{f:.4g} {u}".strip())
            except ValueError:
                cells.append(v)
        lines.append(f"| {n} | `{name}` | " + " | ".join(cells) + " |")
        by = to_bytes(r[idx["dram__bytes_read.sum"]], units[idx["dram__bytes_read.sum"]]) + \
            to_bytes(r[idx["dram__bytes_write.sum"]], units[idx["dram__bytes_write.sum"]])
        for g, members in GROUPS.items():
            if any(name.startswith(m) for m in members):
                traffic.setdefault(g, []).append((name, by))
    with open(out, "w") as f:
        f.write(f"Source: `{rep}` (`ncu --set full --clock-control none`; per-launch, cold-ish caches, serialised).\n\n")
        f.write("\n".join(lines) + "\n")
    if "--traffic" in sys.argv:
        i = sys.argv.index("--traffic")
        key_suffix, path = sys.argv[i + 1], sys.argv[i + 2]
        try:
            cur = json.load(open(path))
        except Exception:
            cur = {}
        for g, items in traffic.items():
            # one "launch" of a group = one Polyblur iteration: distinct kernels once each
            seen, total = set(), 0.0
            for name, by in items:
                if name in seen:
                    continue
                seen.add(name)
                total += by
            cur[f"{g}:{key_suffix}"] = total
        json.dump(cur, open(path, "w"), indent=1, sort_keys=True)
    def num(r, key):
        try:
            return float(r[idx[key]].replace(",", ""))
        except Exception:
            return None

    if "--pipes" in sys.argv:
        i = sys.argv.index("--pipes")
        key_suffix, path = sys.argv[i + 1], sys.argv[i + 2]
        try:
            cur = json.load(open(path))
        except Exception:
            cur = {}
        seen = set()
        for r in body:
            name = r[idx["Kernel Name"]].split("(")[0].replace("void ", "").replace("pb::", "").split("<")[0]
            if name in seen:
                continue
            seen.add(name)
            cur[f"{name}:{key_suffix}"] = {
                "dram_pct": num(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
                "l1tex_pct": num(r, "l1tex__throughput.avg.pct_of_peak_sustained_elapsed"),
                "l2_pct": num(r, "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
                "fma_pipe_pct": num(r, "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
                "issue_pct": num(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
                "warps_pct": num(r, "sm__warps_active.avg.pct_of_peak_sustained_active"),
                "source": rep.split("/")[-1],
            }
        json.dump(cur, open(path, "w"), indent=1, sort_keys=True)
    if "--csv" in sys.argv:
        path = sys.argv[sys.argv.index("--csv") + 1]
        keep = ["ID", "Kernel Name"] + [k for k, _ in COLS if k in idx] + [
            k for k in hdr if k.startswith(("smsp__pcsamp_warps_issue_stalled", "smsp__average_warps_issue_stalled",
                                            "l1tex__data_bank_conflicts_pipe_lsu_mem_shared", "sm__inst_executed_pipe_",
                                            "launch__occupancy", "smsp__warp_issue_stalled")) and k.endswith(
                                                ("ratio", ".sum", "pct_of_peak_sustained_active", "limit_registers",
                                                 "limit_shared_mem", "limit_warps"))]
        keep = [k for i2, k in enumerate(keep) if k in idx and k not in keep[:i2]]
        with open(path, "w", newline="") as f:
            w = csv.writer(f)
            w.writerow(keep)
            w.writerow([units[idx[k]] for k in keep])
            for r in body:
                w.writerow([r[idx[k]] for k in keep])
    print("wrote", out)


if __name__ == "__main__":
    main()
