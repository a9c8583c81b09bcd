#!/usr/bin/env python
"""X25519 soak: 65,536 random (k, u) rows on the GPU against the Python oracle (all host cores).  python tests/checks/x25519_soak.py"""
import sys, os, numpy as np, multiprocessing as mp
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import fourq_b200 as fq
from oracle import fourq_oracle as O
def f(a): return O.x25519(a[0], a[1])
n = 1 << 16
rng = np.random.default_rng(123)
k = rng.integers(0, 256, (n, 32), np.uint8); u = rng.integers(0, 256, (n, 32), np.uint8)
got = fq.x25519(k, u)
with mp.get_context("fork").Pool(os.cpu_count()) as pool:
    want = pool.map(f, [(bytes(k[i]), bytes(u[i])) for i in range(n)], chunksize=256)
bad = sum(1 for i in range(n) if bytes(got[i]) != want[i])
print("x25519 soak rows", n, "mismatches", bad)
