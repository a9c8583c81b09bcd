#!/usr/bin/env python
"""Executed IMAD.WIDE (and all) instructions per row of EVERY kernel in an ncu report that holds the SourceCounters section.
usage: python tools/ncu_kernels_opmix.py X.ncu-rep ROWS > profiles/NAME_opmix.json
ROWS = rows each launch processed (a multiple of 128).  Output: {kernel name: {"wide_per_row", "inst_per_row", "launches"}};
when a kernel was launched several times the first launch is kept."""
import csv
import io
import json
import re
import subprocess
import sys


def main():
    rep, rows = sys.argv[1], float(sys.argv[2])
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    out, cur, hdr = {}, None, None
    for r in csv.reader(io.StringIO(raw)):
        if r and r[0] == "Kernel Name":
            name = r[1]
            if name in out:
                out[name]["launches"] += 1; cur = None
            else:
                cur = out[name] = {"wide": 0.0, "inst": 0.0, "launches": 1}
            hdr = None
            continue
        if cur is None:
            continue
        if hdr is None:
            if "Source" in r and "Instructions Executed" in r:
                hdr = {h: i for i, h in enumerate(r)}
            continue
        if len(r) < len(hdr):
            continue
        m = re.match(r"\s*(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[hdr["Source"]])
        if not m:
            continue
        try:
            n = float(r[hdr["Instructions Executed"]] or 0)
        except ValueError:
            continue
        cur["inst"] += n
        if m.group(1).startswith(("IMAD.WIDE", "IMAD.HI")):
            cur["wide"] += n
    res = {k: {"wide_per_row": v["wide"] * 32 / rows, "inst_per_row": v["inst"] * 32 / rows, "launches": v["launches"]} for k, v in out.items()}
    print(json.dumps(res, indent=1, sort_keys=True))


if __name__ == "__main__":
    main()
