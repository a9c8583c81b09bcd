"""Batched GF((2^127-1)^2) arithmetic -- the counterpart of the reference's impl/fields.py class GFp2.

An element is a 32-byte row LE128(re) | LE128(im); arrays are (N, 32) uint8.  Like the reference's functions on Python
ints (fields.py:157-199) any 128-bit value per half is accepted and results are canonical (< p)."""
import numpy as np

from . import _lib

p1271 = (1 << 127) - 1      # fields.py:5


def _unary(fn_name, a, ndev):
    a = _lib.rows(a, 32, "a")
    out = np.empty_like(a)
    _lib.check(getattr(_lib.lib(), fn_name)(_lib.ptr(a), _lib.ptr(out), a.shape[0], ndev))
    return out


def _binary(fn_name, a, b, ndev):
    a = _lib.rows(a, 32, "a")
    b = _lib.rows(b, 32, "b")
    if a.shape != b.shape:
        raise ValueError("operands must have the same shape")
    out = np.empty_like(a)
    _lib.check(getattr(_lib.lib(), fn_name)(_lib.ptr(a), _lib.ptr(b), _lib.ptr(out), a.shape[0], ndev))
    return out


def _select(fn_name, width, c, x, y, ndev):
    """out = y ^ ((mask * c) & (x ^ y)) on the device (fields.py:59-64): c == 1 -> x, c == 0 -> y, no reduction."""
    x = _lib.rows(x, width, "x")
    y = _lib.rows(y, width, "y")
    c = np.ascontiguousarray(np.asarray(c).reshape(-1), dtype=np.uint8)
    if x.shape != y.shape or c.shape[0] != x.shape[0]:
        raise ValueError("c must have one entry per row and x, y the same shape")
    out = np.empty_like(x)
    _lib.check(getattr(_lib.lib(), fn_name)(_lib.ptr(c), _lib.ptr(x), _lib.ptr(y), _lib.ptr(out), x.shape[0], ndev))
    return out


class GFp2:
    """Static methods named after fields.py GFp2.*"""

    @staticmethod
    def mul(a, b, ndev=1):       # fields.py:167-173
        return _binary("fq_fp2_mul", a, b, ndev)

    @staticmethod
    def sqr(a, ndev=1):          # fields.py:176-181
        return _unary("fq_fp2_sqr", a, ndev)

    @staticmethod
    def inv(a, ndev=1):          # fields.py:194-199
        return _unary("fq_fp2_inv", a, ndev)

    @staticmethod
    def add(a, b, ndev=1):       # fields.py:157-159
        return _binary("fq_fp2_add", a, b, ndev)

    @staticmethod
    def sub(a, b, ndev=1):       # fields.py:162-164
        return _binary("fq_fp2_sub", a, b, ndev)

    @staticmethod
    def neg(a, ndev=1):          # fields.py:184-186
        return _unary("fq_fp2_neg", a, ndev)

    @staticmethod
    def conj(a, ndev=1):         # fields.py:189-191
        return _unary("fq_fp2_conj", a, ndev)

    @staticmethod
    def invsqrt(a, ndev=1):      # fields.py:201-230 (the reference's own control flow, see include/fourq_b200.h)
        return _unary("fq_fp2_invsqrt", a, ndev)

    @staticmethod
    def select(c, x, y, ndev=1):     # fields.py:236-238; c is a (N,) array of 0/1 (uint8), one condition per row
        return _select("fq_fp2_select", 32, c, x, y, ndev)


def _fp(op, a, b, ndev):
    a = _lib.rows(a, 16, "a")
    if b is not None:
        b = _lib.rows(b, 16, "b")
        if a.shape != b.shape:
            raise ValueError("operands must have the same shape")
    out = np.empty_like(a)
    _lib.check(_lib.lib().fq_fp_op(_lib.FPOP[op], _lib.ptr(a), _lib.ptr(b), _lib.ptr(out), a.shape[0], ndev))
    return out


class GFp:
    """Static methods named after fields.py GFp.*; an element is a 16-byte little-endian row (fields.py:125-126), arrays are
    (N, 16) uint8.  Any 128-bit input is accepted (the reference reduces ints mod p), results are canonical."""

    @staticmethod
    def add(x, y, ndev=1):       # fields.py:30-33
        return _fp("add", x, y, ndev)

    @staticmethod
    def sub(x, y, ndev=1):       # fields.py:36-39
        return _fp("sub", x, y, ndev)

    @staticmethod
    def mul(x, y, ndev=1):       # fields.py:42-45
        return _fp("mul", x, y, ndev)

    @staticmethod
    def sqr(x, ndev=1):          # fields.py:48-51
        return _fp("sqr", x, None, ndev)

    @staticmethod
    def neg(x, ndev=1):          # fields.py:54-57
        return _fp("neg", x, None, ndev)

    @staticmethod
    def inv(x, ndev=1):          # fields.py:67-106
        return _fp("inv", x, None, ndev)

    @staticmethod
    def invsqrt(x, ndev=1):      # fields.py:110-122
        return _fp("invsqrt", x, None, ndev)

    @staticmethod
    def select(c, x, y, ndev=1):     # fields.py:59-64; c is a (N,) array of 0/1 (uint8), one condition per row
        return _select("fq_fp_select", 16, c, x, y, ndev)


def _f25(op, a, b, ndev):
    a = _lib.rows(a, 32, "a")
    if b is not None:
        b = _lib.rows(b, 32, "b")
        if a.shape != b.shape:
            raise ValueError("operands must have the same shape")
    out = np.empty_like(a)
    _lib.check(_lib.lib().fq_fp25519_op(_lib.FPOP[op], _lib.ptr(a), _lib.ptr(b), _lib.ptr(out), a.shape[0], ndev))
    return out


class GFp25519:
    """Static methods named after fields.py GFp25519.* (the comparison field of compare.py:14-49); an element is a 32-byte
    little-endian row, arrays are (N, 32) uint8.  Any 256-bit input is accepted (the reference reduces ints mod p), results
    are canonical.  cswap is X25519's ladder step and lives inside the x25519 kernel."""

    @staticmethod
    def add(x, y, ndev=1):       # fields.py:267-270
        return _f25("add", x, y, ndev)

    @staticmethod
    def sub(x, y, ndev=1):       # fields.py:273-276
        return _f25("sub", x, y, ndev)

    @staticmethod
    def mul(x, y, ndev=1):       # fields.py:279-282
        return _f25("mul", x, y, ndev)

    @staticmethod
    def sqr(x, ndev=1):          # fields.py:285-288
        return _f25("sqr", x, None, ndev)

    @staticmethod
    def inv(x, ndev=1):          # fields.py:293-362
        return _f25("inv", x, None, ndev)


def pack(pairs):
    """[(re, im), ...] Python ints -> (N, 32) uint8 rows (fields.py:125-126 packing of each half)."""
    out = np.empty((len(pairs), 32), np.uint8)
    for i, (re, im) in enumerate(pairs):
        out[i] = np.frombuffer(int(re).to_bytes(16, "little") + int(im).to_bytes(16, "little"), np.uint8)
    return out


def unpack(rows):
    rows = _lib.rows(rows, 32)
    return [(int.from_bytes(bytes(r[:16]), "little"), int.from_bytes(bytes(r[16:]), "little")) for r in rows]
