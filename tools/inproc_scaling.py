#!/usr/bin/env python
"""In-process multi-GPU scaling through the C ABI's own slice dispatcher (no torchrun, one process, ndev = 1, 2, 4, 8):
BASELINE cfg 4 (2^24 fixed-base keygen rows, strong scaling) and cfg 3 (ndev x 2^20 variable-base DH rows, weak scaling),
page-locked host arrays, wall clock of the whole call, best of 5; every result is compared with the ndev = 1 bytes.
    [FQ_TRACE=1] python tools/inproc_scaling.py > profiles/rNN_inproc_scaling.jsonl"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fourq_b200 as fq   # noqa: E402
from fourq_b200 import _lib   # noqa: E402,F401


def best(fn, reps=5):
    fn()
    t = 1e30
    for _ in range(reps):
        t0 = time.perf_counter(); fn(); t = min(t, time.perf_counter() - t0)
    return t


def main():
    gpus = fq.device_count()
    ns = [g for g in (1, 2, 4, 8) if g <= gpus]
    sliced = os.environ.get("FQ_SLICED", "1") != "0"          # FQ_SLICED=0: ordinary page-locked arrays (for the A/B)
    print(json.dumps({"gpu_numa_nodes": [int(_lib.lib().fq_device_numa_node(i)) for i in range(gpus)], "sliced_arrays": sliced}), flush=True)
    n4 = 1 << 24
    top = max(ns) if sliced else 1
    pk = fq.pinned_empty((n4, 32), ndev=top); pk[:] = np.random.default_rng(5).integers(0, 256, (n4, 32), np.uint8)
    ref = fq.pinned_empty((n4, 32), ndev=top); po = fq.pinned_empty((n4, 32), ndev=top)
    t1 = None
    for g in ns:
        o = ref if g == 1 else po
        t = best(lambda: fq.MUL_base(pk, out=o, ndev=g))
        t1 = t1 or t
        print(json.dumps({"config": "cfg4 fixed-base keygen, 2^24 rows, strong scaling", "ndev": g, "ms": t * 1e3, "rows_per_s": n4 / t, "speedup_vs_ndev1": t1 / t,
                          "device_span_ms": fq.last_kernel_ms(), "rows_per_device": fq.device.last_rows_per_device(g), "parity_vs_ndev1": bool((o == ref).all())}), flush=True)
    # the same 64 B of host traffic per row with next to no arithmetic: GFp2.neg on 2^24 rows -- what the box can move
    pn = fq.pinned_empty((n4, 32), ndev=top)
    for g in ns:
        t = best(lambda: _lib.check(_lib.lib().fq_fp2_neg(_lib.ptr(pk), _lib.ptr(pn), n4, g)))
        print(json.dumps({"config": "copy-bound probe: GFp2.neg, 2^24 rows, 32 B in + 32 B out per row", "ndev": g, "ms": t * 1e3, "rows_per_s": n4 / t,
                          "host_traffic_gbs": n4 * 64 / t / 1e9}), flush=True)
    rows = 1 << 20
    e1 = None
    for g in ns:
        n3 = rows * g
        gg = g if sliced else 1
        k = fq.pinned_empty((n3, 32), ndev=gg); k[:] = np.random.default_rng(3).integers(0, 256, (n3, 32), np.uint8)
        pub = fq.pinned_empty((n3, 32), ndev=gg); pub[:] = ref[:n3]
        o = fq.pinned_empty((n3, 32), ndev=gg); s = fq.pinned_empty((n3,), ndev=gg)
        t = best(lambda: fq.DH(k, pub, out=o, status=s, ndev=g))
        span = fq.last_kernel_ms(); rpd = fq.device.last_rows_per_device(g)
        e1 = e1 or rows / t
        o1, s1 = fq.DH(k[rows - 65536:rows + 65536] if g > 1 else k[:131072], ref[rows - 65536:rows + 65536] if g > 1 else ref[:131072], ndev=1)
        got = o[rows - 65536:rows + 65536] if g > 1 else o[:131072]
        print(json.dumps({"config": "cfg3 variable-base DH, 2^20 rows per GPU, weak scaling", "ndev": g, "ms": t * 1e3, "rows_per_s": n3 / t, "scaling_vs_ndev1": n3 / t / e1,
                          "device_span_ms": span, "rows_per_device": rpd, "parity_vs_ndev1": bool((o1 == got).all()) and not s.any()}), flush=True)


if __name__ == "__main__":
    main()
