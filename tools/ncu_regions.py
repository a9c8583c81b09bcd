#!/usr/bin/env python
"""Where a kernel's time goes, by code region: groups the SASS lines of `ncu --page source --csv` output by how often they
execute per row (a loop body, a called routine, straight-line code) and prints each region's share of the stall samples,
its relative cost per executed instruction and its top stall reasons.
usage: ncu -i X.ncu-rep --page source --csv --kernel-name regex:NAME > src.csv; python tools/ncu_regions.py src.csv [rows]"""
import csv
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    nrows = float(sys.argv[2]) if len(sys.argv) > 2 else float(1 << 20)
    # a report may hold several launches of the kernel: keep the first table
    starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
    end = starts[1] if len(starts) > 1 else len(rows)
    name, hdr, data = rows[starts[0]][1], rows[starts[0] + 1], rows[starts[0] + 2:end]
    isrc, isamp, iex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
    st = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    warps = nrows / 32
    tot = sum(int(r[isamp]) for r in data if len(r) > isamp)
    groups, cur = [], None
    for n, r in enumerate(data):
        if len(r) <= max(isamp, iex):
            continue
        m = round(int(r[iex]) / warps, 2)
        if cur is None or abs(cur["m"] - m) > 0.01 * max(1, m):
            cur = {"m": m, "start": n, "n": 0, "samp": 0, "st": [0] * len(st), "first": r[isrc].strip()}
            groups.append(cur)
        cur["n"] += 1
        cur["samp"] += int(r[isamp])
        for j, i in enumerate(st):
            cur["st"][j] += int(r[i] or 0)
    dyn_all = sum(g["n"] * g["m"] for g in groups)
    print(name)
    print("static instructions %d, executed per row %.0f, stall samples %d" % (len(data), dyn_all, tot))
    print("%8s %6s %8s %10s %8s %9s  %s" % ("first", "instrs", "x/row", "exec/row", "time %", "cost/inst", "top stall reasons"))
    for g in groups:
        if g["samp"] < tot * 0.008:
            continue
        dyn = g["n"] * g["m"]
        s = sum(g["st"]) or 1
        top = sorted(zip(g["st"], [hdr[i][6:] for i in st]), reverse=True)[:5]
        print("%8d %6d %8.0f %10.0f %8.1f %9.2f  %s" % (g["start"], g["n"], g["m"], dyn, 100.0 * g["samp"] / tot,
              (g["samp"] / tot) / (dyn / dyn_all), " ".join("%s=%.0f%%" % (k, 100.0 * v / s) for v, k in top)))


if __name__ == "__main__":
    main()
