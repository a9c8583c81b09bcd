#!/usr/bin/env python
"""Secondary measurements for BASELINE.json configs 2, 4 and 5 (bench.py is config 3, the headline).

    python tests/checks/bench_configs.py [--gpus N] [--quick] > profiles/rNN_configs.jsonl

One JSON line per measurement: kernel-only time with device-resident inputs (CUDA events, best of 5 after 2 warm-ups, L2
flushed between launches) and, where it makes sense, the end-to-end rate through the public API from pinned host arrays.
Every measured batch is spot-checked bit-for-bit against the oracle (first rows) before it is reported.

cfg 2  GF(p^2) field microbench: 2^26 mul / sqr (HBM-bound: 96 / 64 B per element), 2^20 inv (multiplier-bound)
cfg 4  fixed-base keygen: 2^24 scalars x G, per-digit tables ("comb") vs MUL_windowed / MUL_endo with a table, N GPUs
       (comb: 62 mixed additions x 336 multiply-adds + one inversion per 4 rows (1,504 / 4) + 5 GF(p^2) multiplications)
cfg 5  compare.py analogue: X25519 ladder vs Curve4Q DH, 2^20 rows each
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import fourq_b200 as fq                       # noqa: E402
from fourq_b200 import device as fqdev        # noqa: E402
from oracle import fourq_oracle as O          # noqa: E402


def kernel_ms(op, dev, a, b, out, st, n, reps=5):
    for _ in range(2):
        fqdev.dev_run(op, dev, a, b, out, st, n)
    best = 1e30
    for _ in range(reps):
        fqdev.flush_l2(dev)
        best = min(best, fqdev.dev_run(op, dev, a, b, out, st, n))
    return best


def emit(**kw):
    print(json.dumps(kw), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--quick", action="store_true", help="2^22 / 2^20 / 2^18 rows instead of 2^26 / 2^24 / 2^20")
    args = ap.parse_args()
    dev = 0
    hbm = 6547.5
    try:
        hbm = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except (OSError, KeyError):
        pass
    wide_peak, _ = fqdev.imad_peak(dev)

    # ---------------------------------------------------------------- cfg 2
    n = 1 << (22 if args.quick else 26)
    rng = np.random.default_rng(2)
    a = rng.integers(0, 256, (n, 32), np.uint8); b = rng.integers(0, 256, (n, 32), np.uint8)
    for x in (a, b):
        x[:, 15] &= 0x7F; x[:, 31] &= 0x7F
    da = fqdev.DeviceBuffer.from_host(dev, a); db = fqdev.DeviceBuffer.from_host(dev, b); do = fqdev.DeviceBuffer(dev, n * 32)
    for op, nbytes, imads in (("fp2_mul", 96, 48), ("fp2_sqr", 64, 32)):
        ms = kernel_ms(op, dev, da, db if op == "fp2_mul" else None, do, None, n)
        got = do.to_host((256, 32))
        for j in range(256):
            assert bytes(got[j]) == O.row_fp2(op[4:], bytes(a[j]), bytes(b[j]) if op == "fp2_mul" else None), (op, j)
        emit(config="cfg2", op=op, rows=n, kernel_ms=ms, rows_per_s=n / ms * 1e3, gbs=n * nbytes / ms / 1e6, hbm_peak_gbs=hbm,
             frac_of_hbm=n * nbytes / ms / 1e6 / hbm, bound="hbm", note="%d B and %d multiply-adds per element" % (nbytes, imads))
    ni = 1 << (18 if args.quick else 20)
    ms = kernel_ms("fp2_inv", dev, da, None, do, None, ni)
    got = do.to_host((64, 32))
    for j in range(64):
        assert bytes(got[j]) == O.row_fp2("inv", bytes(a[j]), None), j
    inv_imads = 1504 // 4 + 3 * 48                 # one x^(p-2) chain per 4 rows (Montgomery's trick) + 3 GF(p^2) multiplications per row
    emit(config="cfg2", op="fp2_inv", rows=ni, kernel_ms=ms, rows_per_s=ni / ms * 1e3, imad_wide_per_s=ni * inv_imads / ms * 1e3,
         frac_of_imad_peak=ni * inv_imads / ms * 1e3 / wide_peak, bound="imad",
         note="%d multiply-adds per row: 1,504 per inversion chain (SURVEY 8d) shared by 4 rows + 3 multiplications" % inv_imads)
    del da, db, do

    # ---------------------------------------------------------------- cfg 4
    n = 1 << (20 if args.quick else 24)
    k = np.random.default_rng(5).integers(0, 256, (n, 32), np.uint8)
    dk = fqdev.DeviceBuffer.from_host(dev, k); do = fqdev.DeviceBuffer(dev, n * 32)
    ref = None
    for alg, op, imads in (("comb", "mul_base_comb", 20832 + 1504 // 4 + 5 * 48), ("endo", "mul_endo_base", 64 * 656 + 1504 // 4 + 5 * 48), ("windowed", "mul_base", 62 * (4 * 272 + 384) + 1504 // 4 + 5 * 48)):
        ms = kernel_ms(op, dev, dk, None, do, None, n, reps=3)
        got = do.to_host((n, 32))
        if ref is None:
            ref = got
            for j in range(0, 4096, 64):
                assert bytes(got[j]) == O.row_mul_base(bytes(k[j])), j
        assert (got == ref).all(), alg
        emit(config="cfg4", op="fixed-base keygen [k]G", algorithm=alg, rows=n, n_gpus=1, kernel_ms=ms, rows_per_s=n / ms * 1e3,
             imads_per_row=imads, frac_of_imad_peak=n * imads / ms * 1e3 / wide_peak)
    del dk, do
    pk = fq.pinned_empty((n, 32)); pk[:] = k
    po = fq.pinned_empty((n, 32))
    for g in [x for x in (1, 2, 4, 8, 16) if x <= args.gpus]:
        fq.MUL_base(pk, ndev=g, out=po)
        t0 = time.perf_counter()
        fq.MUL_base(pk, ndev=g, out=po)
        dt = time.perf_counter() - t0
        assert (po == ref).all()
        emit(config="cfg4", op="fixed-base keygen [k]G end to end (pinned host in/out)", algorithm="comb", rows=n, n_gpus=g, wall_ms=dt * 1e3,
             rows_per_s=n / dt, pcie_gbs=n * 64 / dt / 1e9)

    # ---------------------------------------------------------------- cfg 5
    n = 1 << (18 if args.quick else 20)
    kk = np.random.default_rng(6).integers(0, 256, (n, 32), np.uint8)
    uu = np.random.default_rng(7).integers(0, 256, (n, 32), np.uint8)
    dk = fqdev.DeviceBuffer.from_host(dev, kk); du = fqdev.DeviceBuffer.from_host(dev, uu); do = fqdev.DeviceBuffer(dev, n * 32); ds = fqdev.DeviceBuffer(dev, n)
    ms_x = kernel_ms("x25519", dev, dk, du, do, None, n)
    got = do.to_host((64, 32))
    for j in range(64):
        assert bytes(got[j]) == O.x25519(bytes(kk[j]), bytes(uu[j])), j
    pub = fq.MUL_base(np.random.default_rng(4).integers(0, 256, (n, 32), np.uint8))
    dp = fqdev.DeviceBuffer.from_host(dev, pub)
    res = {"x25519": ms_x}
    for alg, op in (("endo", "dh_endo"), ("windowed", "dh")):
        res[alg] = kernel_ms(op, dev, dk, dp, do, ds, n)
    for name, ms in res.items():
        emit(config="cfg5", op="x25519 ladder" if name == "x25519" else "Curve4Q DH (%s)" % name, rows=n, kernel_ms=ms, rows_per_s=n / ms * 1e3)
    emit(config="cfg5", op="ratio Curve4Q DH / X25519 throughput", endo=res["x25519"] / res["endo"], windowed=res["x25519"] / res["windowed"],
         note="the draft claims >2x with endomorphisms, 1.2-1.6x without (draft-ladd-cfrg-4q.md:170-171)")
    if args.gpus > 1:
        g = args.gpus
        pk = fq.pinned_empty((n * g, 32)); pu = fq.pinned_empty((n * g, 32)); po = fq.pinned_empty((n * g, 32)); ps = fq.pinned_empty((n * g,))
        pk[:] = np.tile(kk, (g, 1)); pu[:] = np.tile(uu, (g, 1))
        for name, call in (("x25519", lambda: fq.x25519(pk, pu, ndev=g, out=po)),):
            call(); t0 = time.perf_counter(); call(); dt = time.perf_counter() - t0
            emit(config="cfg5", op=name + " end to end", rows=n * g, n_gpus=g, wall_ms=dt * 1e3, rows_per_s=n * g / dt)
        pp = fq.pinned_empty((n * g, 32)); pp[:] = np.tile(pub, (g, 1))
        fq.DH(pk, pp, ndev=g, out=po, status=ps); t0 = time.perf_counter(); fq.DH(pk, pp, ndev=g, out=po, status=ps); dt = time.perf_counter() - t0
        emit(config="cfg5", op="Curve4Q DH (endo) end to end", rows=n * g, n_gpus=g, wall_ms=dt * 1e3, rows_per_s=n * g / dt)


if __name__ == "__main__":
    main()
