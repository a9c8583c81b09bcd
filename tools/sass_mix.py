#!/usr/bin/env python
"""Per-kernel SASS opcode histogram from `cuobjdump -sass` (offline proxy for pipe balance).
usage: python tools/sass_mix.py <cubin|so|exe> [function-regex]"""
import collections
import re
import subprocess
import sys

FMA_PIPE = ("IMAD", "FFMA", "FMUL", "FADD", "HFMA2")
ALU_PIPE = ("IADD3", "LOP3", "SHF", "LEA", "PRMT", "SEL", "ISETP", "IABS", "FMNMX", "PLOP3", "VIADD", "IADD", "MOV", "VIMNMX", "SGXT", "BMSK")


def main():
    path = sys.argv[1]
    pat = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    fn, hist = None, collections.OrderedDict()
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            fn = m.group(1)
            hist[fn] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and fn:
            hist[fn][m.group(1)] += 1
    for fn, h in hist.items():
        if pat and not pat.search(fn):
            continue
        tot = sum(h.values())
        fma = sum(v for k, v in h.items() if k.startswith(FMA_PIPE))
        alu = sum(v for k, v in h.items() if k.startswith(ALU_PIPE) and not k.startswith("IMAD"))
        wide = sum(v for k, v in h.items() if k.startswith("IMAD.WIDE") or k.startswith("IMAD.HI"))
        mem = sum(v for k, v in h.items() if k.startswith(("LD", "ST", "ATOM", "RED")))
        print("%s\n  total=%d fma_pipe=%d (mul: %d) alu_pipe=%d mem=%d other=%d" % (fn, tot, fma, wide, alu, mem, tot - fma - alu - mem))
        print("  " + " ".join("%s=%d" % kv for kv in h.most_common(14)))


if __name__ == "__main__":
    main()
