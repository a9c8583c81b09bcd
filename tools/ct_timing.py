#!/usr/bin/env python
"""Timing of the DH kernels under scalar distributions that drive the masked table loads differently:
random scalars, one scalar for all rows (every lane of a warp selects the same entry in every step), and one scalar per lane
position (lanes of a warp differ, warps are identical).  Prints the CUDA-event times of k_dh_prep / k_dh_ladder / k_dh_finish.
    python tools/ct_timing.py > profiles/rNN_ct_timing.jsonl"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fourq_b200 as fq                       # noqa: E402
from fourq_b200 import device as fqdev        # noqa: E402

n = 1 << 20
rng = np.random.default_rng(51)
pub = fq.MUL_base(rng.integers(0, 256, (n, 32), np.uint8))
cases = {
    "random scalars": rng.integers(0, 256, (n, 32), np.uint8),
    "one scalar for all rows": np.tile(rng.integers(0, 256, (1, 32), np.uint8), (n, 1)),
    "32 scalars, one per lane position": np.tile(rng.integers(0, 256, (32, 32), np.uint8), (n // 32, 1)),
    "scalar 1 for all rows": np.tile(np.frombuffer((1).to_bytes(32, "little"), np.uint8), (n, 1)),
}
dp = fqdev.DeviceBuffer.from_host(0, pub)
do = fqdev.DeviceBuffer(0, n * 32); ds = fqdev.DeviceBuffer(0, n)
for alg, op in (("endo", "dh_endo"), ("windowed", "dh")):
    for name, k in cases.items():
        dk = fqdev.DeviceBuffer.from_host(0, k)
        for _ in range(2):
            fqdev.dev_run(op, 0, dk, dp, do, ds, n)
        ph = []
        for _ in range(5):
            fqdev.flush_l2(0)
            fqdev.dev_run(op, 0, dk, dp, do, ds, n)
            ph.append(fqdev.last_phase_ms())
        med = [float(np.median([p[i] for p in ph])) for i in range(3)]
        print(json.dumps({"select_mode": "strict scan" if fq.get_select_mode() else "masked loads", "algorithm": alg, "scalars": name, "rows": n, "k_dh_prep_ms": med[0], "k_dh_ladder_ms": med[1], "k_dh_finish_ms": med[2]}), flush=True)
