#!/usr/bin/env python
"""End-to-end DH rate with page-locked vs ordinary (pageable) numpy arrays.  python tools/pageable_check.py"""
import os, sys, time, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fourq_b200 as fq
n = 1 << 20
rng = np.random.default_rng(1)
k = rng.integers(0, 256, (n, 32), np.uint8)
pub = fq.MUL_base(rng.integers(0, 256, (n, 32), np.uint8))
pk = fq.pinned_empty((n, 32)); pk[:] = k
pp = fq.pinned_empty((n, 32)); pp[:] = pub
po = fq.pinned_empty((n, 32)); ps = fq.pinned_empty((n,))
for name, args, kw in (("pinned in/out", (pk, pp), dict(out=po, status=ps)), ("pageable in, fresh out", (k, pub), {}), ("pageable in, pinned out", (k, pub), dict(out=po, status=ps))):
    for _ in range(2): fq.DH(*args, **kw)
    t = time.perf_counter()
    for _ in range(5): fq.DH(*args, **kw)
    dt = (time.perf_counter() - t) / 5
    print("%-26s %.2f ms  %.1f M rows/s" % (name, dt * 1e3, n / dt / 1e6))
