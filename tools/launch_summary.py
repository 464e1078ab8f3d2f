#!/usr/bin/env python
"""Condenses an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel totals and
shares (markdown).    python tools/launch_summary.py launches.csv out.md "<how it was produced>" """
import collections
import csv
import sys


def main():
    src, out, how = sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else ""
    rows = [r for r in csv.reader(open(src)) if r and not r[0].startswith("==")]
    hdr = rows[0]
    ci = {h: i for i, h in enumerate(hdr)}
    tot = collections.OrderedDict()
    for r in rows[1:]:
        if len(r) < len(hdr) or r[ci["Metric Name"]] != "gpu__time_duration.sum":
            continue
        name = r[ci["Kernel Name"]].split("(")[0].replace("void ", "").replace("pb::", "")
        v = float(r[ci["Metric Value"]].replace(",", ""))
        unit = r[ci["Metric Unit"]]
        ms = v / 1e6 if unit in ("ns", "nsecond") else v / 1e3 if unit in ("us", "usecond") else v
        t = tot.setdefault(name, [0, 0.0])
        t[0] += 1
        t[1] += ms
    total = sum(t[1] for t in tot.values())
    with open(out, "w") as f:
        f.write(f"Source: {how}\n")
        f.write("Cold-cache, serialised launches: compare SHARES with bench.py's CUDA-event shares "
                "(`roofline.kernels_ms_per_step`), not absolutes.\n\n")
        f.write("| kernel | launches | total ms | share |\n|---|---|---|---|\n")
        for k, (n, ms) in tot.items():
            f.write(f"| `{k}` | {n} | {ms:.3f} | {ms / total:.3f} |\n")
        f.write(f"| total | {sum(t[0] for t in tot.values())} | {total:.3f} | 1.000 |\n")
    print("wrote", out)


if __name__ == "__main__":
    main()
