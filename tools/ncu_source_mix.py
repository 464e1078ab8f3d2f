#!/usr/bin/env python
"""Instruction mix, shared-memory wavefronts and stall samples per kernel from the source page of an
ncu report:  ncu -i X.ncu-rep --page source --csv --print-source sass | python tools/ncu_source_mix.py
"""
import csv, sys, collections, re

def main():
    rd = csv.reader(sys.stdin)
    kern = None; hdr = None; out = []
    stats = None
    def flush():
        if kern is None or stats is None: return
        tot = sum(v[0] for v in stats['op'].values())
        print(f"\n## {kern[:110]}\nwarp instructions {tot:.4g}; shared wavefronts {stats['wf']:.4g} (ideal {stats['wfi']:.4g}); "
              f"L1 global tag requests {stats['l1g']:.4g}; samples {stats['samples']}")
        print("| opcode | warp inst | % | samples % |\n|---|---|---|---|")
        for op, (n, s) in sorted(stats['op'].items(), key=lambda kv: -kv[1][0])[:28]:
            print(f"| {op} | {n:.4g} | {100*n/max(tot,1):.1f} | {100*s/max(stats['samples'],1):.1f} |")
        st = stats['stall']; ts = sum(st.values())
        print("stalls: " + ", ".join(f"{k[6:]} {100*v/max(ts,1):.0f}%" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:8]))
    for row in rd:
        if not row: continue
        if row[0] == "Kernel Name":
            flush(); kern = row[1]; hdr = None
            stats = dict(op=collections.defaultdict(lambda: [0, 0]), wf=0, wfi=0, l1g=0, samples=0, stall=collections.Counter())
            continue
        if row[0] == "Address":
            hdr = {h: i for i, h in enumerate(row)}; continue
        if hdr is None: continue
        src = row[hdr["Source"]].strip()
        m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", src)
        if not m: continue
        op = m.group(2)
        parts = op.split('.')
        key = parts[0]
        if key in ("LDS", "STS", "LDG", "STG", "LD", "ST", "LDL", "STL"):
            for p in parts[1:]:
                if p in ("64", "128", "U8", "U16"): key += "." + p
        f = lambda name: float(row[hdr[name]] or 0) if name in hdr else 0.0
        n = f("Instructions Executed"); s = f("# Samples")
        stats['op'][key][0] += n; stats['op'][key][1] += s
        stats['wf'] += f("L1 Wavefronts Shared"); stats['wfi'] += f("L1 Wavefronts Shared Ideal")
        stats['l1g'] += f("L1 Tag Requests Global"); stats['samples'] += s
        for h in hdr:
            if h.startswith("stall_") and "Not Issued" not in h:
                stats['stall'][h] += float(row[hdr[h]] or 0)
    flush()
main()
