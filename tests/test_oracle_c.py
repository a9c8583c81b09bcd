"""Pins oracle/fourq_oracle.c (the C restatement used for full-size parity checks) to the golden vectors generated from
the reference's own code and to the Python oracle on random rows.  CPU only."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import c_oracle as C            # noqa: E402
from oracle import fourq_oracle as O        # noqa: E402

H = bytes.fromhex


def R(lst):
    return np.frombuffer(b"".join(lst), np.uint8).reshape(len(lst), -1).copy()


def hexrows(a):
    return [bytes(r).hex() for r in a]


@pytest.mark.parametrize("op", ["mul", "add", "sub", "sqr", "neg", "conj", "inv"])
def test_fp2_golden(golden, op):
    rows = golden["fields"][op]
    a = R([H(r[0]) for r in rows]); b = R([H(r[1]) for r in rows]) if len(rows[0]) == 3 else None
    assert hexrows(C.fp2(op, a, b)) == [r[-1] for r in rows]


@pytest.mark.parametrize("op", ["add", "sub", "mul", "sqr", "neg", "inv", "invsqrt"])
def test_fp_golden(golden, op):
    rows = golden["fp"][op]
    a = R([H(r[0]) for r in rows]); b = R([H(r[1]) for r in rows]) if len(rows[0]) == 3 else None
    assert hexrows(C.fp(op, a, b)) == [r[-1] for r in rows]


def test_codec_golden(golden):
    c = golden["codec"]
    assert hexrows(C.encode(R([H(r[0]) for r in c["encode"]]))) == [r[1] for r in c["encode"]]
    xy, st = C.decode(R([H(r[0]) for r in c["decode"]]))
    assert [int(s) for s in st] == [r[1] for r in c["decode"]]
    assert hexrows(xy) == [r[2] for r in c["decode"]]
    assert set(int(s) for s in st) == {0, 1, 2, 3, 4}


def test_on_curve(golden):
    xy = R([H(r[2]) for r in golden["codec"]["decode"] if r[1] == 0])
    assert C.on_curve(xy).all()
    bad = xy.copy(); bad[:, 3] ^= 1
    assert not C.on_curve(bad).any()
    for j in range(0, len(xy), 16):
        assert O.on_curve(O.xy_from_bytes(bytes(bad[j]))) is False and O.on_curve(O.xy_from_bytes(bytes(xy[j]))) is True


def test_scalar_mult_golden(golden):
    m = golden["mul"]
    rows = m["dh"]
    out, st = C.dh(R([H(r[0]) for r in rows]), R([H(r[1]) for r in rows]))
    assert [(bytes(o).hex(), int(s)) for o, s in zip(out, st)] == [(r[3], r[2]) for r in rows]
    rows = m["mul_base"]
    assert hexrows(C.mul_base(R([H(r[0]) for r in rows]))) == [r[1] for r in rows]
    rows = m["dh_base"]
    out, st = C.dh_base(R([H(r[0]) for r in rows]))
    assert [(bytes(o).hex(), int(s)) for o, s in zip(out, st)] == [(r[2], r[1]) for r in rows]
    rows = m["dh_affine"]
    out, st = C.dh_affine(R([H(r[0]) for r in rows]), R([H(r[1]) for r in rows]))
    assert [(bytes(o).hex(), int(s)) for o, s in zip(out, st)] == [(r[3], r[2]) for r in rows]


def test_random_rows_vs_python_oracle():
    rng = np.random.default_rng(31)
    n = 96
    k = rng.integers(0, 256, (n, 32), np.uint8)
    pub = C.mul_base(rng.integers(0, 256, (n, 32), np.uint8))
    pub[::4] = rng.integers(0, 256, (len(pub[::4]), 32), np.uint8)
    out, st = C.dh(k, pub)
    assert [(bytes(o), int(s)) for o, s in zip(out, st)] == [O.row_dh(bytes(k[j]), bytes(pub[j])) for j in range(n)]
    kb = rng.integers(0, 256, (16, 32), np.uint8)
    assert [bytes(r) for r in C.mul_base(kb)] == [O.row_mul_base(bytes(r)) for r in kb]
    # threads give the same answer as one call
    big_k = np.tile(k, (8, 1)); big_p = np.tile(pub, (8, 1))
    o2, s2 = C.dh(big_k, big_p, threads=8)
    assert (o2 == np.tile(out, (8, 1))).all() and (s2 == np.tile(st, 8)).all()


def _cfg1_scalars(golden):
    import hashlib
    k = np.random.default_rng(1).integers(0, 256, (1024, 32), np.uint8)
    assert hashlib.sha256(k.tobytes()).hexdigest() == golden["cfg1"]["scalars_sha256"]
    return k


def test_baseline_config1_reference_vectors(golden):
    """BASELINE.json configs[0]: 1,024 scalars x G through the reference's DH_windowed (tests/golden/cfg1.json)."""
    k = _cfg1_scalars(golden)
    out, st = C.dh_base(k)
    assert not st.any() and hexrows(out) == golden["cfg1"]["out"]
    for j in range(0, 1024, 64):                                   # the Python oracle on a sample (6 ms per row)
        assert O.row_dh_base(bytes(k[j])) == (H(golden["cfg1"]["out"][j]), 0)
