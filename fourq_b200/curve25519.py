"""Batched X25519 (RFC 7748) -- counterpart of the reference's impl/curve25519.py:88-91 x25519(k, u)."""
import numpy as np

from . import _lib


def x25519(k, u, ndev=1, out=None):
    k = _lib.rows(k, 32, "k")
    u = _lib.rows(u, 32, "u")
    if k.shape[0] != u.shape[0]:
        raise ValueError("k and u must have the same number of rows")
    if out is None:
        out = np.empty((k.shape[0], 32), np.uint8)
    elif not isinstance(out, np.ndarray) or out.dtype != np.uint8 or out.shape != (k.shape[0], 32) or not out.flags["C_CONTIGUOUS"]:
        raise ValueError("out must be a C-contiguous uint8 array of shape (N, 32)")
    _lib.check(_lib.lib().fq_x25519(_lib.ptr(k), _lib.ptr(u), _lib.ptr(out), k.shape[0], ndev))
    return out
