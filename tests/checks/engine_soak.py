#!/usr/bin/env python
"""Soak of the host engine on real GPUs: random batch sizes, GPU counts, operations and buffer kinds (page-locked, pageable, sliced),
several caller threads at once; every multi-GPU result is compared byte for byte with the single-GPU result of the same call, and a
sample of rows with the C oracle.     python tests/checks/engine_soak.py [seconds]"""
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import fourq_b200 as fq                     # noqa: E402
from fourq_b200 import device               # noqa: E402
from oracle import c_oracle as C            # noqa: E402

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
G = fq.device_count()
stats = {"gpus": G, "calls": 0, "rows": 0, "mismatches": 0, "uneven_slices_seen": 0, "by_op": {}}
lock = threading.Lock()


def buf(shape, kind, ndev, rng):
    if kind == "pageable":
        return np.empty(shape, np.uint8)
    return fq.pinned_empty(shape, ndev=ndev if kind == "sliced" else 1)


def worker(seed, t_end):
    rng = np.random.default_rng(seed)
    while time.time() < t_end:
        op = rng.choice(["dh", "keygen", "fp2_mul", "x25519"])
        n = int(rng.choice([1, 7, 129, 4097, 70001, 300007, 1200011, 2500003])) if op != "x25519" else int(rng.choice([1, 129, 70001, 300007]))
        ndev = int(rng.integers(1, G + 1))
        kinds = [str(rng.choice(["pinned", "pageable", "sliced"])) for _ in range(3)]
        a = buf((n, 32), kinds[0], ndev, rng); a[:] = rng.integers(0, 256, (n, 32), np.uint8)
        if op == "keygen":
            ref = fq.MUL_base(a, ndev=1)
            out = fq.MUL_base(a, ndev=ndev, out=buf((n, 32), kinds[2], ndev, rng))
            bad = int((out != ref).any(axis=1).sum())
            m = min(n, 512); bad += int((C.mul_base(np.array(a[:m])) != ref[:m]).any(axis=1).sum())
        elif op == "fp2_mul":
            b = buf((n, 32), kinds[1], ndev, rng); b[:] = rng.integers(0, 256, (n, 32), np.uint8)
            ref = fq.GFp2.mul(a, b, ndev=1); out = fq.GFp2.mul(a, b, ndev=ndev)
            bad = int((out != ref).any(axis=1).sum())
        elif op == "x25519":
            b = buf((n, 32), kinds[1], ndev, rng); b[:] = rng.integers(0, 256, (n, 32), np.uint8)
            ref = fq.x25519(a, b, ndev=1); out = fq.x25519(a, b, ndev=ndev)
            bad = int((out != ref).any(axis=1).sum())
        else:
            pub = buf((n, 32), kinds[1], ndev, rng); pub[:] = fq.MUL_base(rng.integers(0, 256, (n, 32), np.uint8))
            pub[::17] = rng.integers(0, 256, (len(pub[::17]), 32), np.uint8)           # some undecodable strings
            alg = str(rng.choice(["endo", "windowed"]))
            ref, rst = fq.DH(a, pub, ndev=1, algorithm=alg)
            out, st = fq.DH(a, pub, ndev=ndev, algorithm=alg, out=buf((n, 32), kinds[2], ndev, rng), status=buf((n,), kinds[2], ndev, rng))
            bad = int(((out != ref).any(axis=1) | (st != rst)).sum())
            m = min(n, 512); w, ws = C.dh(np.array(a[:m]), np.array(pub[:m])); bad += int(((w != ref[:m]).any(axis=1) | (ws != rst[:m])).sum())
        rows = device.last_rows_per_device(ndev)
        with lock:
            stats["calls"] += 1; stats["rows"] += n; stats["mismatches"] += bad
            stats["by_op"][op] = stats["by_op"].get(op, 0) + 1
            if ndev > 1 and n >= 1000 and max(rows) - min(rows) > 256:
                stats["uneven_slices_seen"] += 1
        if bad:
            print("MISMATCH", op, n, ndev, kinds, bad, flush=True)


t_end = time.time() + budget
ts = [threading.Thread(target=worker, args=(100 + i, t_end)) for i in range(3)]
[t.start() for t in ts]; [t.join() for t in ts]
fq.trim()
print(json.dumps(stats))
sys.exit(1 if stats["mismatches"] else 0)
