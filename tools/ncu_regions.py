#!/usr/bin/env python
"""Summarise an `ncu --page source --csv` dump: samples per SASS region.

    ncu -i X.ncu-rep --page source --csv > src.csv ; python tools/ncu_regions.py src.csv [N]

Prints the warp-stall samples grouped into runs of consecutive instructions
(split where the sample density changes), the top stall reasons of each, and
per-opcode totals.  Used to see where the consumer warps of the staged batch
kernel spend their time."""
import csv
import sys
from collections import defaultdict


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr = rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    data = rows[2:]

    def f(r, k):
        try:
            return float(r[ix[k]])
        except (ValueError, KeyError):
            return 0.0
    stall = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    total = sum(f(r, "# Samples") for r in data)
    print("total samples %.0f, instructions executed %.0f" % (total, sum(f(r, "Instructions Executed") for r in data)))
    # regions: split at branch targets / big jumps in execution count
    regions = []
    cur = []
    prev = None
    for r in data:
        n = f(r, "Instructions Executed")
        if prev is not None and (n > 1.5 * prev + 1000 or prev > 1.5 * n + 1000) and cur:
            regions.append(cur)
            cur = []
        cur.append(r)
        prev = n
    if cur:
        regions.append(cur)
    out = []
    for reg in regions:
        s = sum(f(r, "# Samples") for r in reg)
        if s < total * 0.004:
            continue
        st = {h: sum(f(r, h) for r in reg) for h in stall}
        top = sorted(st.items(), key=lambda kv: -kv[1])[:3]
        out.append((s, reg, top))
    for s, reg, top in sorted(out, key=lambda x: -x[0])[: int(sys.argv[2]) if len(sys.argv) > 2 else 25]:
        a0, a1 = reg[0][ix["Address"]][-5:], reg[-1][ix["Address"]][-5:]
        ex = f(reg[len(reg) // 2], "Instructions Executed")
        thr = f(reg[len(reg) // 2], "Avg. Threads Executed")
        ops = defaultdict(int)
        for r in reg:
            src = r[ix["Source"]].split()
            if src:
                ops[src[1] if src[0].startswith("@") and len(src) > 1 else src[0]] += 1
        opss = ",".join("%s x%d" % kv for kv in sorted(ops.items(), key=lambda kv: -kv[1])[:4])
        print("%5.1f%%  %s-%s  %4d instr, executed %10.0f x, %4.1f thr  | %s | %s" % (
            100 * s / total, a0, a1, len(reg), ex, thr, " ".join("%s %.0f" % (k[6:], v) for k, v in top), opss))


if __name__ == "__main__":
    main()
