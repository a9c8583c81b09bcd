"""Batched X25519 (RFC 7748) -- counterpart of the reference's impl/curve25519.py:88-91 x25519(k, u)."""
import numpy as np

from . import _lib


def x25519(k, u, ndev=1):
    k = _lib.rows(k, 32, "k")
    u = _lib.rows(u, 32, "u")
    if k.shape[0] != u.shape[0]:
        raise ValueError("k and u must have the same number of rows")
    out = np.empty((k.shape[0], 32), np.uint8)
    _lib.check(_lib.lib().fq_x25519(_lib.ptr(k), _lib.ptr(u), _lib.ptr(out), k.shape[0], ndev))
    return out
