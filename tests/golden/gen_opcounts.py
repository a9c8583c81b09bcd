#!/usr/bin/env python
"""The reference's field-operation counts (compare.py:51-169 `compare_ops`) produced BY THE REFERENCE'S OWN COUNTERS
(fields.py:10-27, 135-154, 241-256), loaded through ref_loader.py.  Writes tests/golden/opcounts.json, which
tools/compare_ops.py prints next to the executed multiply-adds of the CUDA kernels.

    python tests/golden/gen_opcounts.py          (build container only: needs /root/reference)

Rows and order are compare.py's; one more block counts the GF(p) operations (the reference's GFp counters, which compare.py
does not print: GFp.ctr() is broken at fields.py:27, so the class attributes are read directly) and the byte-level pipeline
decode + DH + encode that BASELINE config 3 runs.  The scalar is fixed (compare.py draws it with getrandbits; the counts do not
depend on it except through the parity fix-up of MUL_windowed, which costs no field operation)."""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_loader  # noqa: E402

fields, c4q, c25519, _ = ref_loader.load_reference()
GFp, GFp2, GFp25519 = fields.GFp, fields.GFp2, fields.GFp25519


def counted(fn, field=None):
    field = field or GFp2
    GFp.ctr_enabled = True
    field.ctr_reset()
    GFp.A = GFp.S = GFp.M = GFp.I = 0
    fn()
    a, s, m, i = field.ctr()
    out = {"M": m, "S": s, "A": a, "I": i}
    if field is GFp2:
        out["GFp"] = {"M": GFp.M, "S": GFp.S, "A": GFp.A, "I": GFp.I}
    GFp.ctr_enabled = False
    return out


def main():
    m = 0x1A2B3C4D5E6F708192A3B4C5D6E7F8091A2B3C4D5E6F708192A3B4C5D6E7F809 % (1 << 256)
    G = c4q.AffineToR1(c4q.Gx, c4q.Gy)
    k = bytes.fromhex("77076d0a7318a57d3c16c17251b26645df4c2f87ebc0992ab177fba51db92c2a")
    u = bytes.fromhex("09" + "00" * 31)
    GFp.ctr_enabled = False
    G392 = c4q.MUL_endo(392, G)
    T_w, T_e = c4q.table_windowed(G), c4q.table_endo(G)
    T392_w, T392_e = c4q.table_windowed(G392), c4q.table_endo(G392)
    G2, G3 = c4q.R1toR2(G), c4q.R1toR3(G)
    Gaff = G[:2]
    enc = bytearray(c4q.encode(*Gaff))
    rows = [
        ("R1toR2", lambda: c4q.R1toR2(G)), ("R1toR3", lambda: c4q.R1toR3(G)), ("R2toR4", lambda: c4q.R2toR4(G2)),
        ("ADD_core", lambda: c4q.ADD_core(G3, G2)), ("ADD", lambda: c4q.ADD(G, G2)), ("DBL", lambda: c4q.DBL(G)),
        ("phi", lambda: c4q.phi(G)), ("psi", lambda: c4q.psi(G)),
        ("MUL_windowed", lambda: c4q.MUL_windowed(m, G)), ("MUL_windowed_fixed", lambda: c4q.MUL_windowed(m, G, table=T_w)),
        ("MUL_endo", lambda: c4q.MUL_endo(m, G)), ("MUL_endo_fixed", lambda: c4q.MUL_endo(m, G, table=T_e)),
        ("DH_windowed", lambda: c4q.DH_windowed(m, Gaff)), ("DH_windowed_fixed", lambda: c4q.DH_windowed(m, Gaff, table=T392_w)),
        ("DH_endo", lambda: c4q.DH_endo(m, Gaff)), ("DH_endo_fixed", lambda: c4q.DH_endo(m, Gaff, table=T392_e)),
        # not in compare.py: the pieces of BASELINE config 3 that it leaves out
        ("decode", lambda: c4q.decode(bytearray(enc))), ("encode", lambda: c4q.encode(*Gaff)), ("GFp2.inv", lambda: GFp2.inv(Gaff[0])),
        ("table_windowed", lambda: c4q.table_windowed(G)), ("table_endo", lambda: c4q.table_endo(G)),
        ("decode+DH_windowed+encode", lambda: c4q.encode(*c4q.DH_windowed(m, c4q.decode(bytearray(enc))))),
        ("decode+DH_endo+encode", lambda: c4q.encode(*c4q.DH_endo(m, c4q.decode(bytearray(enc))))),
    ]
    out = {"scalar": hex(m), "source": "impl/compare.py:51-169 run under tests/golden/ref_loader.py; counters of impl/fields.py", "rows": []}
    for name, fn in rows:
        out["rows"].append(dict(name=name, **counted(fn)))
    out["rows"].append(dict(name="x25519", **counted(lambda: c25519.x25519(k, u), GFp25519)))
    path = os.path.join(HERE, "opcounts.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
        f.write("\n")
    print("wrote %s" % path)
    for r in out["rows"]:
        print("%-28s M %7.1f  S %7.1f  A %7.1f  I %4.1f   GF(p): %s" % (r["name"], r["M"], r["S"], r["A"], r["I"], r.get("GFp")))


if __name__ == "__main__":
    main()
