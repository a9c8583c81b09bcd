#!/usr/bin/env python
"""Generate the committed golden vectors in tests/golden/*.json FROM THE REFERENCE ITSELF.

Run in the build container (needs /root/reference):   python tests/golden/gen_golden.py

Every output value below is produced by calling the reference's own functions
(impl/fields.py, impl/curve4q.py, impl/curve25519.py -- loaded by ref_loader.py, which only applies
py2->py3 syntax edits in memory).  Before generating, the reference's 64 self-checks are run and
must all print [PASS]; their output is stored in reference_selftest.txt.

Inputs are seeded (random.Random), so the files are reproducible.  Conventions of the byte rows
(shared with include/fourq_b200.h and oracle/fourq_oracle.py):
  * GF(p^2) element  = 32 bytes: LE128(re) | LE128(im)
  * affine point      = 64 bytes: x0|x1|y0|y1, each LE128
  * scalar            = 32 bytes little-endian unsigned (curve4q.py:558-559 limb order)
  * status            = 0 ok, 1 reserved bit (curve4q.py:53), 2 y>=p (:62), 3 AttributeError quirk (:77),
                        4 not on curve (:94/:448), 5 neutral result (:460); failed rows are all-zero.
"""
import json
import os
import random
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_loader  # noqa: E402

fields, c4q, c25519, reftest = ref_loader.load_reference()
GFp, GFp2, p = fields.GFp, fields.GFp2, fields.p1271
fields.GFp.ctr_enabled = False


def le16(x):
    return int(x).to_bytes(16, "little")


def f2b(a):
    return (le16(a[0]) + le16(a[1])).hex()


def xyb(P):
    return (le16(P[0][0]) + le16(P[0][1]) + le16(P[1][0]) + le16(P[1][1])).hex()


def kbytes(m):
    return int(m).to_bytes(32, "little").hex()


def ref_decode_status(enc):
    """Calls the reference decode(); maps its exceptions to status codes (see module docstring)."""
    B = bytearray(enc)
    reserved = bool(B[15] & 0x80)          # the condition of curve4q.py:52, evaluated before the call
    try:
        P = c4q.decode(B)
        return 0, P
    except AttributeError:
        return 3, None
    except Exception as e:                  # noqa: BLE001 -- the reference raises bare Exception
        msg = str(e)
        if "reserved bit" in msg:
            return (1 if reserved else 2), None
        if "not on curve" in msg:
            return 4, None
        raise


def ref_dh_status(fn, m, P, table=None):
    try:
        return 0, fn(m, P, table=table)
    except Exception as e:                  # noqa: BLE001
        msg = str(e)
        if "not on curve" in msg:
            return 4, None
        if "neutral" in msg:
            return 5, None
        raise


def dump(name, obj):
    path = os.path.join(HERE, name)
    with open(path, "w") as f:
        json.dump(obj, f, indent=0, sort_keys=True)
        f.write("\n")
    print("wrote %s (%d bytes)" % (name, os.path.getsize(path)))


def main():
    lines = ref_loader.run_reference_selftests((fields, c4q, c25519, reftest))
    bad = [l for l in lines if not l.startswith("[PASS]")]
    with open(os.path.join(HERE, "reference_selftest.txt"), "w") as f:
        f.write("\n".join(lines) + "\n")
    if bad or len(lines) != 64:
        raise SystemExit("reference self-tests did not all pass: %r" % bad)
    print("reference self-tests: %d/64 PASS" % len(lines))

    N = c4q.N
    G = (c4q.Gx, c4q.Gy)
    G_R1 = c4q.AffineToR1(*G)
    T_G = c4q.table_windowed(G_R1)
    G392 = c4q.MUL_endo(392, G_R1)                       # as curve4q.py:758
    T392w = c4q.table_windowed(G392)
    T392e = c4q.table_endo(G392)

    # ------------------------------------------------------------------ fields
    rng = random.Random(0x4F1E1D5)
    edge = [0, 1, 2, p - 1, p, p + 1, 1 << 126, 1 << 127, (1 << 127) + 1, (1 << 128) - 1, (1 << 128) - 2,
            (1 << 64) - 1, 1 << 64, 0xFFFFFFFF, 1 << 32, (1 << 96) - 1]
    f2_inputs = [(a, b) for a in edge[:11] for b in edge[:11]]
    f2_inputs += [(rng.getrandbits(128), rng.getrandbits(128)) for _ in range(192)]
    f2_inputs += [(rng.getrandbits(127), rng.getrandbits(127)) for _ in range(64)]
    fld = {"mul": [], "sqr": [], "add": [], "sub": [], "neg": [], "conj": [], "inv": [], "fp_inv": [], "fp_invsqrt": []}
    for i, a in enumerate(f2_inputs):
        b = f2_inputs[(i * 7 + 3) % len(f2_inputs)]
        fld["mul"].append([f2b(a), f2b(b), f2b(GFp2.mul(a, b))])
        fld["add"].append([f2b(a), f2b(b), f2b(GFp2.add(a, b))])
        fld["sub"].append([f2b(a), f2b(b), f2b(GFp2.sub(a, b))])
        fld["sqr"].append([f2b(a), f2b(GFp2.sqr(a))])
        ar = (a[0] % p, a[1] % p)   # neg/conj/inv are only meaningful on reduced input in the reference
        fld["neg"].append([f2b(ar), f2b(GFp2.neg(ar))])
        fld["conj"].append([f2b(ar), f2b(GFp2.conj(ar))])
        if i % 4 == 0 or i < 121:
            fld["inv"].append([f2b(ar), f2b(GFp2.inv(ar))])
    for x in edge + [rng.getrandbits(127) for _ in range(48)]:
        xr = x % p
        fld["fp_inv"].append([le16(xr).hex(), le16(GFp.inv(xr)).hex()])
        fld["fp_invsqrt"].append([le16(xr).hex(), le16(GFp.invsqrt(xr)).hex()])
    dump("fields.json", fld)

    # ------------------------------------------------------------------ GF(p) ops on 16-byte rows (fields.py:29-122)
    rng = random.Random(0xF9127)
    fp_inputs = [(a, b) for a in edge for b in edge[:8]]
    fp_inputs += [(rng.getrandbits(128), rng.getrandbits(128)) for _ in range(96)]
    fp_inputs += [(rng.getrandbits(127), rng.getrandbits(127)) for _ in range(64)]
    fpv = {"add": [], "sub": [], "mul": [], "sqr": [], "neg": [], "inv": [], "invsqrt": []}
    for i, (a, b) in enumerate(fp_inputs):
        fpv["add"].append([le16(a).hex(), le16(b).hex(), le16(GFp.add(a, b)).hex()])
        fpv["sub"].append([le16(a).hex(), le16(b).hex(), le16(GFp.sub(a, b)).hex()])
        fpv["mul"].append([le16(a).hex(), le16(b).hex(), le16(GFp.mul(a, b)).hex()])
        fpv["sqr"].append([le16(a).hex(), le16(GFp.sqr(a)).hex()])
        ar = a % p                              # neg / inv / invsqrt: reduced input, as the reference's callers pass
        fpv["neg"].append([le16(ar).hex(), le16(GFp.neg(ar) % p).hex()])
        if i % 3 == 0:
            fpv["inv"].append([le16(ar).hex(), le16(GFp.inv(ar)).hex()])
            fpv["invsqrt"].append([le16(ar).hex(), le16(GFp.invsqrt(ar)).hex()])
    dump("fp.json", fpv)

    # ------------------------------------------------------------------ points used below
    def mulG(m):
        return c4q.R1toAffine(c4q.MUL_windowed(m, G_R1, table=T_G))

    rng = random.Random(0xC4)
    pts = [mulG(rng.getrandbits(256)) for _ in range(96)]

    # ------------------------------------------------------------------ encode / decode
    codec = {"encode": [], "decode": []}
    codec["Genc"] = bytes(c4q.encode(*G)).hex()
    for P in [G] + pts:
        codec["encode"].append([xyb(P), bytes(c4q.encode(*P)).hex()])
    # low-order / special points: encode works on any affine pair
    special = [((0, 0), (1, 0)), ((0, 0), (p - 1, 0)), ((0, 1), (0, 0)), ((0, p - 1), (0, 0)),
               ((1, 0), (0, 0)), ((p - 1, 0), (0, 0)), ((0, 1 << 126), (5, 7)), ((0, (1 << 126) - 1), (5, 7))]
    for P in special:
        codec["encode"].append([xyb(P), bytes(c4q.encode(*P)).hex()])

    dec_inputs = []
    for P in [G] + pts:
        e = bytes(c4q.encode(*P))
        dec_inputs.append(e)
        flip = bytearray(e); flip[31] ^= 0x80          # other sign: decodes to -x (still valid)
        dec_inputs.append(bytes(flip))
    for P in special:
        dec_inputs.append(bytes(c4q.encode(*P)))
    e = bytearray(c4q.encode(*G)); e[15] |= 0x80; dec_inputs.append(bytes(e))                 # class 1
    e = bytearray(os.urandom(0)) + bytearray(rng.getrandbits(8) for _ in range(32)); e[15] |= 0x80
    dec_inputs.append(bytes(e))
    dec_inputs.append(le16(p) + le16(5))                                                       # class 2
    dec_inputs.append(le16(5) + le16(p))
    dec_inputs.append(le16(5) + le16(p | (1 << 127)))                                          # y1 == p with sign bit
    dec_inputs.append(le16(p) + le16(p))
    dec_inputs.append(bytes(32))                                                               # y = 0
    dec_inputs.append(le16(1) + le16(0))
    dec_inputs.append(le16(p - 1) + le16(0))
    dec_inputs.append(le16(0) + le16(1))
    dec_inputs.append(le16(0) + le16(p - 1))
    dec_inputs.append(le16(1) + bytes(15) + b"\x80")
    for _ in range(160):                                                                       # class 4 / valid mix
        e = bytearray(rng.getrandbits(8) for _ in range(32)); e[15] &= 0x7F
        dec_inputs.append(bytes(e))
    for e in dec_inputs:
        st, P = ref_decode_status(e)
        codec["decode"].append([e.hex(), st, xyb(P) if st == 0 else "00" * 64])
    dump("codec.json", codec)

    # ------------------------------------------------------------------ scalar multiplication / DH
    edge_k = [0, 1, 2, 3, 15, 16, 17, 31, 32, 33, 392, N - 2, N - 1, N, N + 1, 2 * N, 2 * N + 1, 3 * N - 1,
              (1 << 255) + 5, (1 << 256) - 1, (1 << 256) - 2, 1 << 255, (1 << 246), (1 << 246) - 1,
              1567 * N, 1567 * N + 1, ((1 << 256) // N) * N, ((1 << 256) // N) * N - 1, ((1 << 256) // N) * N + 1,
              0x0028FD6CBDA458F07E38F7C9CFBB91663A8B3C2C6FD86E0C3AD457AB55456230]
    edge_k = [k for k in edge_k if 0 <= k < (1 << 256)]
    rng = random.Random(0xD4)
    ks = edge_k + [rng.getrandbits(256) for _ in range(96)]
    mul = {"mul_base": [], "dh_base": [], "dh": [], "dh_affine": []}
    for k in ks:
        mul["mul_base"].append([kbytes(k), bytes(c4q.encode(*mulG(k))).hex()])
        st, Q = ref_dh_status(c4q.DH_windowed, k, G, T392w)
        mul["dh_base"].append([kbytes(k), st, bytes(c4q.encode(*Q)).hex() if st == 0 else "00" * 32])
        # the endomorphism algorithm must agree wherever it is defined (curve4q.py:706-762)
        st2, Q2 = ref_dh_status(c4q.DH_endo, k, G, T392e)
        assert (st, Q) == (st2, Q2), "DH_windowed != DH_endo for k=%x" % k
    # variable base: k x decode(enc) -> encode
    for i, k in enumerate(ks):
        P = pts[i % len(pts)] if i % 9 else G
        enc = bytes(c4q.encode(*P))
        st, Q = ref_dh_status(c4q.DH_windowed, k, P)
        st2, Q2 = ref_dh_status(c4q.DH_endo, k, P)
        assert (st, Q) == (st2, Q2)
        mul["dh"].append([kbytes(k), enc.hex(), st, bytes(c4q.encode(*Q)).hex() if st == 0 else "00" * 32])
        mul["dh_affine"].append([kbytes(k), xyb(P), st, xyb(Q) if st == 0 else "00" * 64])
    # BASELINE.json configs[0]: 1,024 random 32-byte scalars x the base point through the reference's DH_windowed with the
    # table of [392]G (curve4q.py:743-762), encoded.  Scalars = numpy default_rng(1).integers(0, 256, (1024, 32), uint8)
    # (SURVEY 8d cfg 1); stored as the SHA-256 of the scalars plus the 1,024 outputs.
    import hashlib
    import numpy as np
    k1 = np.random.default_rng(1).integers(0, 256, (1024, 32), np.uint8)
    outs = []
    for row in k1:
        st, Q = ref_dh_status(c4q.DH_windowed, int.from_bytes(bytes(row), "little"), G, T392w)
        assert st == 0
        outs.append(bytes(c4q.encode(*Q)).hex())
    dump("cfg1.json", {"scalars": "numpy.random.default_rng(1).integers(0, 256, (1024, 32), numpy.uint8)",
                       "scalars_sha256": hashlib.sha256(k1.tobytes()).hexdigest(), "out": outs})

    # failure paths of DH_core
    P392 = ((0x1318020702de23bc3c9b73c751b4b192, 0x77ab39a7d8990c0a18e3c409fbd81a95),
            (0x515854b6d19cc2da1ea2b43b5121a22e, 0x763f89e129497361d74dff5063e66682))     # curve4q.py:772-773
    for k, P in [(1, ((0, 0), (0, 0))), (1, P392), (12345, P392), (5, ((0, 0), (1, 0))), (7, ((0, 0), (p - 1, 0))),
                 (9, ((0, 1), (0, 0))), (3, ((1, 2), (3, 4)))]:
        st, Q = ref_dh_status(c4q.DH_windowed, k, P)
        mul["dh_affine"].append([kbytes(k), xyb(P), st, xyb(Q) if st == 0 else "00" * 64])
    # DH on encodings that fail to decode keeps the decode status
    for e, st, _ in codec["decode"]:
        if st != 0 and len(mul["dh"]) < len(ks) + 24:
            mul["dh"].append([kbytes(rng.getrandbits(256)), e, st, "00" * 32])
    # the reference's own KAT chain (curve4q.py:549-567): 1000 chained MUL_windowed from G end at mulP
    mul["mulP_affine"] = xyb(((0x257C122BBFC94A1BDFD2B477BD494BEF, 0x469BF80CB5B11F01769593547237C459),
                              (0x0901B3817C0E936C281C5067996F3344, 0x570B948EACACE2104FE8C429915F1245)))
    sc = [0x3AD457AB55456230, 0x3A8B3C2C6FD86E0C, 0x7E38F7C9CFBB9166, 0x0028FD6CBDA458F0]
    chain = []
    for _ in range(1000):
        sc[1] = sc[2]
        sc[2] = (sc[2] + sc[0]) & 0xFFFFFFFFFFFFFFFF
        chain.append(kbytes(sc[0] + (sc[1] << 64) + (sc[2] << 128) + (sc[3] << 192)))
    mul["mulP_chain_scalars"] = chain
    mul["doubleP_affine"] = xyb(((0x2C3FD8822C82270FC9099C54855859D6, 0x4DA5B9E83AA7A1B2A7B3F6E2043E8E68),
                                 (0x2001EB3A576883963EE089F0EB49AA14, 0x0FFDB0D761421F501FEE5617A7E954CD)))
    mul["P1000_affine"] = xyb(((0x3E243958590C4D906480B1EF0A151DB0, 0x5327AF7D84238CD0AA270F644A65D473),
                               (0x3EF69A49CB7E02375E06003D73C43EB1, 0x293EB1E26DD23B4E4E752648AC2EF0AB)))
    dump("mul.json", mul)

    # ------------------------------------------------------------------ endomorphism pieces (section 8f "next")
    rng = random.Random(0xE4D0)
    endo = {"phi": [], "psi": [], "decompose": [], "recode": [], "recode_windowed": []}
    for P in [G] + pts[:15]:
        R1 = c4q.AffineToR1(*P)
        endo["phi"].append([xyb(P), xyb(c4q.R1toAffine(c4q.phi(R1)))])
        endo["psi"].append([xyb(P), xyb(c4q.R1toAffine(c4q.psi(R1)))])
    for k in edge_k + [rng.getrandbits(256) for _ in range(64)]:
        v = c4q.decompose(k)
        endo["decompose"].append([kbytes(k), [int(x) for x in v]])
        s, d = c4q.recode(v)
        endo["recode"].append([[int(x) for x in v], "".join(str(int(x)) for x in s), "".join(str(int(x)) for x in d)])
    dump("endo.json", endo)

    # ------------------------------------------------------------------ X25519
    rng = random.Random(0x25519)
    xv = []
    nine = bytes([9] + [0] * 31)
    for _ in range(48):
        k = bytes(rng.getrandbits(8) for _ in range(32))
        u = bytes(rng.getrandbits(8) for _ in range(32)) if rng.random() < 0.8 else nine
        xv.append([k.hex(), u.hex(), bytes(c25519.x25519(k, u)).hex()])
    for u in [bytes(32), bytes([1] + [0] * 31), (fields.p25519 - 1).to_bytes(32, "little"),
              fields.p25519.to_bytes(32, "little"), (fields.p25519 + 1).to_bytes(32, "little"), b"\xff" * 32]:
        k = bytes(rng.getrandbits(8) for _ in range(32))
        xv.append([k.hex(), u.hex(), bytes(c25519.x25519(k, u)).hex()])
    dump("x25519.json", {"x25519": xv})

    # ------------------------------------------------------------------ select and GFp2.invsqrt (fields.py:59-64, :236-238, :201-230)
    # A separate file and RNG stream, so that the vectors above stay byte-identical to the round-1 files.
    rng = random.Random(0x5E1EC7)
    sel = {"fp_select": [], "fp2_select": [], "fp2_invsqrt": []}
    vals = edge + [rng.getrandbits(128) for _ in range(40)]
    for i, x in enumerate(vals):
        y = vals[(i * 5 + 2) % len(vals)]
        for c in (0, 1):                                     # the reference's contract; unreduced values pass through unchanged
            sel["fp_select"].append([c, le16(x).hex(), le16(y).hex(), le16(GFp.select(c, x, y)).hex()])
    for c in (2, 3, 128, 255):                               # what the reference's expression gives for other c (mask * c)
        x, y = rng.getrandbits(128), rng.getrandbits(128)
        sel["fp_select"].append([c, le16(x).hex(), le16(y).hex(), le16(GFp.select(c, x, y) & ((1 << 128) - 1)).hex()])
    for i in range(48):
        a = (rng.getrandbits(128), rng.getrandbits(128)); b = (rng.getrandbits(128), rng.getrandbits(128))
        for c in (0, 1):
            sel["fp2_select"].append([c, f2b(a), f2b(b), f2b(GFp2.select(c, a, b))])
    inv_in = [(a, b) for a in (0, 1, 2, 3, 4, p - 1, p - 2, 1 << 126) for b in (0, 1, p - 1, 5)]
    inv_in += [(rng.getrandbits(127) % p, 0) for _ in range(16)]                              # the GF(p) branch, squares and non-squares
    inv_in += [(rng.getrandbits(127) % p, rng.getrandbits(127) % p) for _ in range(96)]
    inv_in += [(5, p), (p, 7), (p + 3, 0), ((1 << 128) - 1, (1 << 128) - 1)]                 # unreduced: a[1] == 0 is tested on the raw value
    for a in inv_in:
        sel["fp2_invsqrt"].append([f2b(a), f2b(GFp2.invsqrt(a))])
    dump("select.json", sel)

    # ------------------------------------------------------------------ GFp25519 field ops (fields.py:267-362), compare_fields' other column
    rng = random.Random(0xF25519)
    GF25 = fields.GFp25519
    q = fields.p25519
    le32 = lambda v: int(v).to_bytes(32, "little").hex()
    edge25 = [0, 1, 2, 19, 38, q - 1, q, q + 1, 2 * q, 2 * q + 1, 2 * q + 37, 1 << 255, (1 << 256) - 1, (1 << 255) - 1, 1 << 128, (1 << 128) - 1]
    vals = edge25 + [rng.getrandbits(256) for _ in range(48)]
    f25 = {"add": [], "sub": [], "mul": [], "sqr": [], "inv": []}
    for i, x in enumerate(vals):
        for y in (vals[(i * 7 + 3) % len(vals)], edge25[i % len(edge25)]):
            for op in ("add", "sub", "mul"):
                f25[op].append([le32(x), le32(y), le32(getattr(GF25, op)(x, y))])
        f25["sqr"].append([le32(x), le32(GF25.sqr(x))])
        f25["inv"].append([le32(x), le32(GF25.inv(x))])
    dump("f25519.json", f25)


if __name__ == "__main__":
    main()
