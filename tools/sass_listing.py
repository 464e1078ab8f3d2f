#!/usr/bin/env python
"""SASS evidence for profiles/: per-kernel mnemonic counts of the shipped library from `cuobjdump -sass`
(FFMA / FADD / LDS / STS / LDG / STG / SHFL, and the Blackwell-specific ones: UBLKCP = cp.async.bulk (TMA 1-D),
SYNCS = mbarrier, LDGSTS = cp.async, UTMALDG / UTMASTG = tensor-map TMA, UTC*MMA / LDTM / STTM = tcgen05).

    python tools/sass_listing.py polyblur_b200/libpolyblur_sm100.so profiles/r02_sass_summary.md
"""
import collections
import re
import subprocess
import sys

WATCH = ["FFMA", "FADD", "FMUL", "LDS", "STS", "LDG", "STG", "LDL", "STL", "SHFL", "BAR", "UBLKCP", "SYNCS", "LDGSTS",
         "UTMALDG", "UTMASTG", "UTCHMMA", "UTCQMMA", "LDTM", "STTM", "HMMA", "CCTL"]
HOT = ("k_fft_rows_fwd2", "k_fft_rows_inv2", "k_fft_cols2", "k_fft_cols", "k_rows3", "k_cols3", "k_rows2", "k_cols2", "k_deconv_narrow", "k_rf_rows",
       "k_patch", "k_table_jobs")


def main():
    lib, out = sys.argv[1], sys.argv[2]
    txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    per = collections.OrderedDict()
    cur = None
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            per[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and cur:
            per[cur][m.group(1).split(".")[0]] += 1
    total = collections.Counter()
    for c in per.values():
        total.update(c)
    demangle = subprocess.run(["c++filt"], input="\n".join(per), capture_output=True, text=True).stdout.splitlines()
    names = dict(zip(per, demangle))
    with open(out, "w") as f:
        f.write(f"`cuobjdump -sass {lib}`: {len(per)} kernels, {sum(total.values())} SASS instructions.\n\n")
        f.write("Whole library: " + ", ".join(f"{k} {total[k]}" for k in WATCH) + "\n\n")
        f.write("| kernel | instr | " + " | ".join(WATCH[:15]) + " |\n|---|---|" + "---|" * 15 + "\n")
        for fn, c in per.items():
            name = names.get(fn, fn)
            if not any(h in name for h in HOT):
                continue
            short = re.sub(r"\(.*", "", name).replace("void ", "").replace("pb::", "")[:90]
            f.write(f"| `{short}` | {sum(c.values())} | " + " | ".join(str(c[k]) for k in WATCH[:15]) + " |\n")
    print("wrote", out)


if __name__ == "__main__":
    main()
