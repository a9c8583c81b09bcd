#!/usr/bin/env python
"""Soak run: N random rows (default 2^24) through every scalar-multiplication path, every row compared with the C oracle.
    python tests/checks/soak.py [log2_rows] > profiles/rNN_soak.json"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import fourq_b200 as fq                     # noqa: E402
from oracle import c_oracle as C            # noqa: E402

lg = int(sys.argv[1]) if len(sys.argv) > 1 else 24
n = 1 << lg
rng = np.random.default_rng(2024)
res = {"rows": n, "host_cores": os.cpu_count()}
t0 = time.time()
kp = rng.integers(0, 256, (n, 32), np.uint8)
pub = fq.MUL_base(kp)
want_pub = C.mul_base(kp)
res["mul_base_comb_mismatches"] = int((pub != want_pub).any(axis=1).sum())
for alg in ("endo", "windowed"):
    res["mul_base_%s_mismatches" % alg] = int((fq.MUL_base(kp, algorithm=alg) != want_pub).any(axis=1).sum())
k = rng.integers(0, 256, (n, 32), np.uint8)
pub[::1009] = rng.integers(0, 256, (len(pub[::1009]), 32), np.uint8)          # ~0.1 % arbitrary strings
want, wst = C.dh(k, pub)
res["status_histogram"] = {int(a): int(b) for a, b in zip(*np.unique(wst, return_counts=True))}
for alg in ("endo", "windowed"):
    for strict in (False, True):
        fq.set_select_mode(strict)
        out, st = fq.DH(k, pub, algorithm=alg)
        res["dh_%s_%s_mismatches" % (alg, "strict" if strict else "masked")] = int(((out != want).any(axis=1) | (st != wst)).sum())
fq.set_select_mode(True)                   # back to the library default
res["seconds"] = time.time() - t0
print(json.dumps(res))
