"""Adversarial inputs shared by the CPU simulation test and the GPU parity test: y coordinates and scalars built from limb
patterns that stress the carry chains, the Mersenne folds and the recoding (all-ones limbs, values around p and around
multiples of N, single bits at limb boundaries), plus valid points so that whole scalar multiplications run on them."""
import numpy as np

from oracle import c_oracle as C
from oracle import fourq_oracle as O

P = O.P127


def special_fp_values():
    v = {0, 1, 2, P - 1, P - 2, P, (1 << 126), (1 << 126) - 1, (1 << 126) + 1, (1 << 127) - 2 ** 32, (1 << 96) - 1, 1 << 96,
         (1 << 64) - 1, 1 << 64, (1 << 32) - 1, 1 << 32, 0xFFFFFFFF00000000FFFFFFFF00000000 & P, 0x00000000FFFFFFFF00000000FFFFFFFF,
         0x7FFFFFFF00000000000000000000000 * 16 & P, 0x5555555555555555555555555555555 * 16 & P, int("a" * 31, 16)}
    for i in (31, 32, 33, 63, 64, 65, 95, 96, 97, 125, 126):
        v.add(1 << i); v.add((1 << i) - 1); v.add(P - (1 << i))
    return sorted(x for x in v if 0 <= x <= P)


def special_scalars():
    N = O.N
    s = {0, 1, 2, 3, 15, 16, 17, 31, 32, 33, 392, N - 2, N - 1, N, N + 1, 2 * N, 2 * N + 1, 3 * N, (1 << 256) - 1, (1 << 256) - 2, 1 << 255,
         (1 << 255) - 1, (1 << 246), (1 << 246) - 1, ((1 << 256) // N) * N, ((1 << 256) // N) * N - 1, ((1 << 256) // N) * N + 1,
         int("f0" * 32, 16), int("0f" * 32, 16), int("ff00" * 16, 16), int("00000000ffffffff" * 4, 16), int("ffffffff00000000" * 4, 16),
         int("8" + "0" * 63, 16), int("7" + "f" * 63, 16), int("1" * 64, 16), int("8" * 64, 16)}
    for i in (32, 64, 96, 128, 160, 192, 224, 245, 246, 247):
        s.add(1 << i); s.add((1 << i) - 1); s.add((1 << i) + 1)
    return sorted(x for x in s if 0 <= x < (1 << 256))


def grid(seed=5):
    """(k, enc): every special scalar against (a) a few valid points, (b) y-only strings built from special field values
    (most are not on the curve: exercises decode's failure classes and the zero-filled rows next to good ones)."""
    rng = np.random.default_rng(seed)
    scal = special_scalars()
    pts = [bytes(r) for r in C.mul_base(rng.integers(0, 256, (6, 32), np.uint8))] + [O.encode(O.GX, O.GY)]
    vals = special_fp_values()
    ys = [(a, b) for a in vals[:: max(1, len(vals) // 12)] for b in (0, 1, P - 1, vals[len(vals) // 2])]
    encs = pts + [int(a).to_bytes(16, "little") + int(b).to_bytes(16, "little") for a, b in ys]
    ks, es = [], []
    for i, m in enumerate(scal):
        for j, e in enumerate(encs):
            if j < len(pts) or (i + j) % 7 == 0:
                ks.append(int(m).to_bytes(32, "little")); es.append(e)
    k = np.frombuffer(b"".join(ks), np.uint8).reshape(-1, 32).copy()
    enc = np.frombuffer(b"".join(es), np.uint8).reshape(-1, 32).copy()
    return k, enc


def fp2_grid():
    vals = special_fp_values()
    vals = vals + [v | (1 << 127) for v in vals[:8]] + [(1 << 128) - 1, (1 << 128) - 2, P + 1]          # non-canonical inputs too
    pairs = [(a, b) for a in vals for b in vals[:: max(1, len(vals) // 9)]]
    a = np.frombuffer(b"".join(int(x).to_bytes(16, "little") + int(y).to_bytes(16, "little") for x, y in pairs), np.uint8).reshape(-1, 32).copy()
    b = np.roll(a, 5, axis=0).copy()
    return a, b
