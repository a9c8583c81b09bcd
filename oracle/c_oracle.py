"""ctypes loader of oracle/_build/libfourq_oracle.so (oracle/fourq_oracle.c) -- TEST INFRASTRUCTURE, NOT PRODUCT.

The C restatement of the reference's Curve4Q path, for checks at sizes the Python oracle cannot reach (2^20 rows in a few
seconds on the host's cores).  Batches are split over threads; ctypes releases the GIL during the calls.
Only tests/, bench.py (as the checker) and __graft_entry__.smoke() may import this module."""
import ctypes
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "libfourq_oracle.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        src = os.path.join(HERE, "fourq_oracle.c")
        if not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
            subprocess.check_call(["make", "-C", HERE, "-s"])
        L = ctypes.CDLL(LIB)
        vp, sz, i = ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int
        L.fqo_dh_batch.argtypes = [vp, vp, vp, vp, sz]; L.fqo_dh_batch.restype = None
        L.fqo_dh_affine_batch.argtypes = [vp, vp, vp, vp, sz]; L.fqo_dh_affine_batch.restype = None
        L.fqo_mul_base_batch.argtypes = [vp, vp, sz]; L.fqo_mul_base_batch.restype = None
        L.fqo_dh_base_batch.argtypes = [vp, vp, vp, sz]; L.fqo_dh_base_batch.restype = None
        L.fqo_decode_batch.argtypes = [vp, vp, vp, sz]; L.fqo_decode_batch.restype = None
        L.fqo_encode_batch.argtypes = [vp, vp, sz]; L.fqo_encode_batch.restype = None
        L.fqo_on_curve_batch.argtypes = [vp, vp, sz]; L.fqo_on_curve_batch.restype = None
        L.fqo_fp2_batch.argtypes = [i, vp, vp, vp, sz]; L.fqo_fp2_batch.restype = i
        L.fqo_fp_batch.argtypes = [i, vp, vp, vp, sz]; L.fqo_fp_batch.restype = i
        _lib = L
    return _lib


def _p(a, off=0):
    return ctypes.c_void_p(a.ctypes.data + off) if a is not None else None


def _arr(a, width):
    a = np.ascontiguousarray(a, dtype=np.uint8)
    assert a.ndim == 2 and a.shape[1] == width, a.shape
    return a


def _parallel(n, fn, threads=None):
    """fn(lo, hi) on contiguous slices of range(n), one slice per thread."""
    threads = threads or min(os.cpu_count() or 1, 64)
    if n < 4 * threads:
        fn(0, n)
        return
    step = (n + threads - 1) // threads
    with ThreadPoolExecutor(threads) as ex:
        list(ex.map(lambda lo: fn(lo, min(n, lo + step)), range(0, n, step)))


def dh(k, enc, threads=None):
    k, enc = _arr(k, 32), _arr(enc, 32)
    n = k.shape[0]
    out = np.zeros((n, 32), np.uint8); st = np.zeros(n, np.uint8)
    L = lib()
    _parallel(n, lambda lo, hi: L.fqo_dh_batch(_p(k, 32 * lo), _p(enc, 32 * lo), _p(out, 32 * lo), _p(st, lo), hi - lo), threads)
    return out, st


def dh_affine(k, xy, threads=None):
    k, xy = _arr(k, 32), _arr(xy, 64)
    n = k.shape[0]
    out = np.zeros((n, 64), np.uint8); st = np.zeros(n, np.uint8)
    L = lib()
    _parallel(n, lambda lo, hi: L.fqo_dh_affine_batch(_p(k, 32 * lo), _p(xy, 64 * lo), _p(out, 64 * lo), _p(st, lo), hi - lo), threads)
    return out, st


def mul_base(k, threads=None):
    k = _arr(k, 32)
    n = k.shape[0]
    out = np.zeros((n, 32), np.uint8)
    L = lib()
    _parallel(n, lambda lo, hi: L.fqo_mul_base_batch(_p(k, 32 * lo), _p(out, 32 * lo), hi - lo), threads)
    return out


def dh_base(k, threads=None):
    k = _arr(k, 32)
    n = k.shape[0]
    out = np.zeros((n, 32), np.uint8); st = np.zeros(n, np.uint8)
    L = lib()
    _parallel(n, lambda lo, hi: L.fqo_dh_base_batch(_p(k, 32 * lo), _p(out, 32 * lo), _p(st, lo), hi - lo), threads)
    return out, st


def decode(enc, threads=None):
    enc = _arr(enc, 32)
    n = enc.shape[0]
    xy = np.zeros((n, 64), np.uint8); st = np.zeros(n, np.uint8)
    L = lib()
    _parallel(n, lambda lo, hi: L.fqo_decode_batch(_p(enc, 32 * lo), _p(xy, 64 * lo), _p(st, lo), hi - lo), threads)
    return xy, st


def encode(xy):
    xy = _arr(xy, 64)
    out = np.zeros((xy.shape[0], 32), np.uint8)
    lib().fqo_encode_batch(_p(xy), _p(out), xy.shape[0])
    return out


def on_curve(xy):
    xy = _arr(xy, 64)
    ok = np.zeros(xy.shape[0], np.uint8)
    lib().fqo_on_curve_batch(_p(xy), _p(ok), xy.shape[0])
    return ok.astype(bool)


FP2_OPS = {"mul": 0, "sqr": 1, "inv": 2, "add": 3, "sub": 4, "neg": 5, "conj": 6}
FP_OPS = {"mul": 0, "sqr": 1, "inv": 2, "add": 3, "sub": 4, "neg": 5, "invsqrt": 6}


def fp2(op, a, b=None):
    a = _arr(a, 32); b = _arr(b, 32) if b is not None else None
    out = np.zeros_like(a)
    assert lib().fqo_fp2_batch(FP2_OPS[op], _p(a), _p(b), _p(out), a.shape[0]) == 0
    return out


def fp(op, a, b=None):
    a = _arr(a, 16); b = _arr(b, 16) if b is not None else None
    out = np.zeros_like(a)
    assert lib().fqo_fp_batch(FP_OPS[op], _p(a), _p(b), _p(out), a.shape[0]) == 0
    return out
