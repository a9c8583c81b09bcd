#!/usr/bin/env python
"""Per-primitive SASS instruction counts (offline proxy for kernel cost).  usage: python tools/opcount.py"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CUBIN = "/tmp/fq_opcount.cubin"


def hist(text):
    fn, out = None, {}
    for line in text.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            fn = m.group(1); out[fn] = collections.Counter(); continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and fn:
            out[fn][m.group(1)] += 1
    return out


def main():
    subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-cubin", "-o", CUBIN,
                           os.path.join(ROOT, "tools", "opcount.cu")] + sys.argv[1:])
    h = hist(subprocess.run(["cuobjdump", "-sass", CUBIN], capture_output=True, text=True).stdout)
    names = sorted({re.sub(r"ILi\dE", "", k) for k in h})
    print("%-28s %6s %6s %6s %6s %6s   (cycles/warp/SMSP: FMA-pipe = 4*wide+2*other_fma, ALU-pipe = 2*alu)" % ("primitive", "total", "wide", "fma_o", "alu", "other"))
    for n in names:
        k2 = [k for k in h if re.sub(r"ILi\dE", "", k) == n and "ILi2E" in k][0]
        k3 = [k for k in h if re.sub(r"ILi\dE", "", k) == n and "ILi3E" in k][0]
        d = collections.Counter(h[k3]); d.subtract(h[k2])
        tot = sum(d.values())
        wide = sum(v for k, v in d.items() if k.startswith(("IMAD.WIDE", "IMAD.HI")))
        fma_o = sum(v for k, v in d.items() if k.startswith(("IMAD", "FFMA", "HFMA2", "FMUL", "FADD"))) - wide
        alu = sum(v for k, v in d.items() if k.startswith(("IADD3", "LOP3", "SHF", "LEA", "PRMT", "SEL", "ISETP", "VIADD", "IADD", "MOV", "PLOP3", "SGXT", "BMSK", "VIMNMX")))
        other = tot - wide - fma_o - alu
        short = re.sub(r"^_Z\d+", "", n).split("EvPj")[0]
        print("%-28s %6d %6d %6d %6d %6d   FMA %5d  ALU %5d  issue %5d   %s" % (short, tot, wide, fma_o, alu, other, 4 * wide + 2 * fma_o, 2 * alu, tot,
              " ".join("%s=%d" % kv for kv in d.most_common(9) if kv[1] > 0)))


if __name__ == "__main__":
    main()
