#!/usr/bin/env python
"""The batched counterpart of the reference's impl/compare.py, on one B200.

    python tools/compare_ops.py [--rows N] [--opmix profiles/rNN_compare_opmix.json] > profiles/rNN_compare.txt
    ncu --section SourceCounters -f -o gpurun_out/compare python tools/compare_ops.py --rows 65536 --launch-only   (then
    python tools/ncu_kernels_opmix.py gpurun_out/compare.ncu-rep 65536 > profiles/rNN_compare_opmix.json)

Three tables, as compare.py prints them (compare.py:14-49, :51-169, :171-219):
  1. time for N field operations (add, mul, sqr, inv) in GF(p^2) and in GF(2^255-19): kernel time with device-resident rows,
     L2 flushed;
  2. field-operation counts M / S / A / I per operation -- the reference's own counters, read from tests/golden/opcounts.json
     (written by tests/golden/gen_opcounts.py from the reference's code) -- next to the 32x32->64 multiply-adds that count
     implies under the limb model of SURVEY 8d (GF(p^2) M = 48, S = 32; GF(p) M = 16, S = 10) and, when an ncu capture is
     given, the IMAD.WIDE instructions the CUDA kernels of that variant actually executed per row;
  3. time for N Diffie-Hellman operations for the reference's five variants (windowed, windowed fixed base, endomorphisms,
     endomorphisms fixed base, X25519) plus this engine's per-digit-table fixed base: kernel time and end-to-end time
     through the public API from page-locked numpy arrays.
Every batch is produced by the product library (no CPU implementation is involved); parity of these entry points is the job
of tests/."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fourq_b200 as fq                       # noqa: E402
from fourq_b200 import device as fqdev        # noqa: E402

# variant -> (device op, kernels it launches, facade call for the end-to-end time)
VARIANTS = [
    ("DH_windowed", "Curve4Q (windowed)", "dh_affine", ["k_dh_prep<1, 0>", "k_dh_ladder<0,", "k_dh_finish<1, 1>"]),
    ("DH_windowed_fixed", "Curve4Q (win fixed base)", "dh_base", ["k_fixed_base<1, 0,", "k_dh_finish<0, 1>"]),
    ("DH_endo", "Curve4Q (endomorphisms)", "dh_endo_affine", ["k_dh_prep<1, 1>", "k_dh_ladder<1,", "k_dh_finish<1, 1>"]),
    ("DH_endo_fixed", "Curve4Q (endo fixed base)", "dh_endo_base", ["k_fixed_base<1, 1,", "k_dh_finish<0, 1>"]),
    ("DH_comb_fixed", "Curve4Q (per-digit tables, fixed base; not in the reference)", "dh_base_comb", ["k_comb<1,", "k_dh_finish<0, 1>"]),
    ("x25519", "Curve25519", "x25519", ["k_x25519(", "k_x25519_finish"]),
]


def best_ms(op, a, b, out, st, n, reps=3):
    fqdev.dev_run(op, 0, a, b, out, st, n)
    t = 1e30
    for _ in range(reps):
        fqdev.flush_l2(0)
        t = min(t, fqdev.dev_run(op, 0, a, b, out, st, n))
    return t


def model_imads(r):
    g = r.get("GFp") or {"M": 0, "S": 0}
    if r["name"] == "x25519":
        return 64 * r["M"] + 36 * r["S"] + 0 * r["I"]          # 8-limb GF(2^255-19): mul 64, sqr 36 (inversion chain already counted in M, S)
    inv = 0 if g["M"] or g["S"] else 1504 * r["I"]            # the GF(p) counters already hold the inversion chain when they were on
    return int(48 * r["M"] + 32 * r["S"] + 16 * g["M"] + 10 * g["S"] + inv)


def print_counts(opmix):
    # ---- 2. operation counts (compare.py:51-169)
    counts = json.load(open(os.path.join(ROOT, "tests", "golden", "opcounts.json")))
    print("===== Field operation count (the reference's counters; tests/golden/opcounts.json) =====")
    print()
    print("%-28s %7s %7s %7s %5s   %7s %7s   %12s %14s" % ("", "M", "S", "A", "I", "GFp.M", "GFp.S", "model IMADs", "executed WIDE"))
    executed = {}
    if opmix:
        for key, _, _, kernels in VARIANTS:
            tot = 0.0
            for pat in kernels:
                hits = [v for name, v in opmix.items() if pat in name.replace("(bool)", "")]
                if not hits:
                    tot = None; break
                tot += hits[0]["wide_per_row"]
            executed[key] = tot
    alias = {"DH_windowed": "DH_windowed", "DH_windowed_fixed": "DH_windowed_fixed", "DH_endo": "DH_endo", "DH_endo_fixed": "DH_endo_fixed", "x25519": "x25519"}
    for r in counts["rows"]:
        g = r.get("GFp") or {"M": 0, "S": 0}
        ex = executed.get(alias.get(r["name"], ""))
        print("%-28s %7.1f %7.1f %7.1f %5.1f   %7d %7d   %12d %14s" % (r["name"], r["M"], r["S"], r["A"], r["I"], g["M"], g["S"], model_imads(r),
                                                                      "%.0f" % ex if ex else "-"))
    if opmix and executed.get("DH_comb_fixed"):
        print("%-28s %7s %7s %7s %5s   %7s %7s   %12s %14.0f" % ("DH_comb_fixed (not in ref.)", "-", "-", "-", "-", "-", "-", "-", executed["DH_comb_fixed"]))
    print("model IMADs = 48 M + 32 S (GF(p^2)) + 16 M + 10 S (GF(p)); constants 2 and 1/2 counted as the reference counts them.")
    print("executed WIDE = IMAD.WIDE.U32[.X] instructions per row summed over the variant's kernels (ncu SourceCounters); it is lower")
    print("than the model where the kernels skip multiplications by 2 and 1/2, use a dedicated GF(p) squaring, and share one inversion per 16 rows.")
    print()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=1 << 20)
    ap.add_argument("--opmix", default=None, help="JSON written by tools/ncu_kernels_opmix.py: executed IMAD.WIDE per row of every kernel")
    ap.add_argument("--launch-only", action="store_true", help="launch every variant once (for an ncu capture) and print nothing")
    ap.add_argument("--counts-only", action="store_true", help="print the operation-count table only (needs no GPU)")
    args = ap.parse_args()
    n = args.rows
    opmix = json.load(open(args.opmix)) if args.opmix else None
    if args.counts_only:
        print_counts(opmix)
        return
    rng = np.random.default_rng(17)
    k = rng.integers(0, 256, (n, 32), np.uint8)
    u = rng.integers(0, 256, (n, 32), np.uint8)
    # the base point G as an affine row x0 | x1 | y0 | y1 (curve4q.py:19-20; its y half is the encoding 87b2cb2b... of curve4q.py:478)
    G = np.frombuffer(bytes.fromhex("aa33387bad92652805b32f7c2372341af677ac60b39f86969caa78283f551f1e"
                                    "87b2cb2b46a224b95a7820a19bee3f0e5c8b4c8444c3a74942020e63f84a1c6e"), np.uint8)
    xy = np.tile(G, (n, 1))                                   # compare.py multiplies the base point (compare.py:172, :189)
    dk = fqdev.DeviceBuffer.from_host(0, k); du = fqdev.DeviceBuffer.from_host(0, u); dxy = fqdev.DeviceBuffer.from_host(0, xy)
    dout = fqdev.DeviceBuffer(0, n * 64); dst = fqdev.DeviceBuffer(0, n)
    inputs = {"dh_affine": dxy, "dh_endo_affine": dxy, "x25519": du}
    if args.launch_only:
        for _, _, op, _ in VARIANTS:
            fqdev.dev_run(op, 0, dk, inputs.get(op), dout, dst, n)
        return
    print("fourq_b200 compare (batched impl/compare.py), %d rows per batch, table selection: %s" % (n, "strict scan" if fq.get_select_mode() else "masked loads"))
    print()
    # ---- 1. field operations (compare.py:14-49)
    print("===== Time for %d field operations (32-byte rows, kernel time, device-resident) =====" % n)
    print()
    print("%-5s %10s %10s %14s %14s" % ("Op", "GFp2", "GFp25519", "GFp2 ops/s", "GFp25519 ops/s"))
    # compare.py:15-17: one corpus of 256-bit values, read as a GF(p^2) pair and as a GF(2^255-19) element
    a = fqdev.DeviceBuffer.from_host(0, rng.integers(0, 256, (n, 32), np.uint8)); b = fqdev.DeviceBuffer.from_host(0, rng.integers(0, 256, (n, 32), np.uint8))
    for name in ("add", "mul", "sqr", "inv"):
        ms = [best_ms(f + name, a, b if name in ("add", "mul") else None, dout, None, n) for f in ("fp2_", "f25519_")]
        print("%-5s %8.4fms %8.4fms %14.4g %14.4g" % (name, ms[0], ms[1], n / ms[0] * 1e3, n / ms[1] * 1e3))
    print()
    print_counts(opmix)
    # ---- 3. DH timing (compare.py:171-219)
    print("===== Time for %d Diffie-Hellman operations =====" % n)
    print()
    print("%-62s %12s %14s %14s" % ("", "kernel time", "rows/s", "end to end"))
    pk = fq.pinned_empty((n, 32)); pk[:] = k
    pxy = fq.pinned_empty((n, 64)); pxy[:] = xy
    pu = fq.pinned_empty((n, 32)); pu[:] = u
    po64 = fq.pinned_empty((n, 64)); po = fq.pinned_empty((n, 32)); ps = fq.pinned_empty((n,))
    calls = {
        "dh_affine": lambda: fq.DH_windowed(pk, pxy, out=po64, status=ps), "dh_endo_affine": lambda: fq.DH_endo(pk, pxy, out=po64, status=ps),
        "dh_base": lambda: fq.DH_base(pk, out=po, status=ps, algorithm="windowed"), "dh_endo_base": lambda: fq.DH_base(pk, out=po, status=ps, algorithm="endo"),
        "dh_base_comb": lambda: fq.DH_base(pk, out=po, status=ps, algorithm="comb"), "x25519": lambda: fq.x25519(pk, pu, out=po),
    }
    res = {}
    for key, label, op, _ in VARIANTS:
        ms = best_ms(op, dk, inputs.get(op), dout, dst, n)
        calls[op]()
        t0 = time.perf_counter(); calls[op](); e2e = (time.perf_counter() - t0) * 1e3
        res[key] = ms
        print("%-62s %10.3fms %14.4g %12.3fms" % (label, ms, n / ms * 1e3, e2e))
    print()
    print("Curve4Q / Curve25519 throughput: windowed %.2fx, endomorphisms %.2fx   (draft-ladd-cfrg-4q.md:170-171 claims 1.2-1.6x and >2x)" % (
        res["x25519"] / res["DH_windowed"], res["x25519"] / res["DH_endo"]))


if __name__ == "__main__":
    main()
