#!/usr/bin/env python
"""Executed-instruction mix and stall profile of one kernel from `ncu --page source --csv` output.
usage: ncu -i X.ncu-rep --page source --csv > src.csv; python tools/ncu_opmix.py src.csv [rows]"""
import collections
import csv
import re
import sys


def main():
    path = sys.argv[1]
    nrows = float(sys.argv[2]) if len(sys.argv) > 2 else float(1 << 20)
    rd = csv.reader(open(path))
    name = next(rd)
    hdr = next(rd)
    ix = {h: i for i, h in enumerate(hdr)}
    execd = collections.Counter(); stalls = collections.Counter(); samples = collections.Counter()
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tot_inst = 0
    for r in rd:
        if r and r[0] == "Kernel Name":          # a report may hold several tables (launches / views) of the kernel: keep the first
            break
        if len(r) < len(hdr):
            continue
        m = re.match(r"\s*(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[ix["Source"]])
        if not m:
            continue
        op = m.group(1)
        n = float(r[ix["Instructions Executed"]] or 0)
        execd[op] += n; tot_inst += n
        samples[op] += float(r[ix["# Samples"]] or 0)
        for c in stall_cols:
            stalls[c] += float(r[ix[c]] or 0)
    warps = nrows / 32
    print(name[1][:100])
    print("warp-instructions: %.4g  per row: %.0f" % (tot_inst, tot_inst / warps))
    grp = lambda pred: sum(v for k, v in execd.items() if pred(k))
    wide = grp(lambda k: k.startswith(("IMAD.WIDE", "IMAD.HI")))
    fma_o = grp(lambda k: k.startswith(("IMAD", "FFMA", "HFMA2", "FMUL", "FADD"))) - wide
    alu = grp(lambda k: k.startswith(("IADD3", "LOP3", "SHF", "LEA", "PRMT", "SEL", "ISETP", "VIADD", "IADD", "MOV", "PLOP3", "SGXT", "BMSK", "VIMNMX", "P2R", "R2P")))
    mem = grp(lambda k: k.startswith(("LD", "ST", "ATOM", "RED")))
    print("per row: wide-mul %.0f | other fma-pipe %.0f | alu-pipe %.0f | mem %.0f | other %.0f" % (
        wide / warps, fma_o / warps, alu / warps, mem / warps, (tot_inst - wide - fma_o - alu - mem) / warps))
    print("pipe cycles per row (4/wide, 2/other): fma %.0f  alu %.0f  issue %.0f" % ((4 * wide + 2 * fma_o) / warps, 2 * alu / warps, tot_inst / warps))
    print("top opcodes per row: " + "  ".join("%s=%.0f" % (k, v / warps) for k, v in execd.most_common(16)))
    ts = sum(stalls.values())
    print("stall samples: " + "  ".join("%s=%.1f%%" % (k[6:], 100 * v / ts) for k, v in stalls.most_common(8)))
    print("samples by opcode: " + "  ".join("%s=%.1f%%" % (k, 100 * v / sum(samples.values())) for k, v in samples.most_common(8)))


if __name__ == "__main__":
    main()
