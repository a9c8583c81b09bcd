"""CPU oracle for the Curve4Q hot path -- TEST INFRASTRUCTURE, NOT PRODUCT.

A Python-3 restatement of the algorithms of the reference (bifurcation/fourq, Python 2):
    impl/fields.py      GF(p), GF(p^2) with p = 2^127-1, GF(2^255-19)
    impl/curve4q.py     Curve4Q encode/decode, point formulas, windowed + endomorphism scalar
                        multiplication, cofactor Diffie-Hellman
    impl/curve25519.py  X25519 (RFC 7748)
Every function cites the reference lines it follows.  Arithmetic is on Python ints, so this is
only for small cases (about 6 ms per DH); oracle/fourq_oracle.c is the fast restatement.

Parity status: PINNED.  tests/test_oracle_golden.py checks this file against tests/golden/*.json,
which tests/golden/gen_golden.py produced by running the reference's OWN code (read from
/root/reference, py2->py3 syntax rewritten in memory by tests/golden/ref_loader.py) after that code
passed all 64 of its own self-checks here.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.  fourq_b200/ never does: the product has no CPU path.

Differences from the reference that do not change any result:
  * exceptions are replaced by status codes (the *_status functions); the raising wrappers keep the
    reference's messages;
  * decode() does not mutate its argument (curve4q.py:56, fields.py:130 do);
  * the reference's accidental AttributeError (curve4q.py:77, `GFp.two` does not exist) is
    reported as status ST_QUIRK_T0 instead of being fixed, so failure parity stays defined by it.
"""

P127 = (1 << 127) - 1                     # fields.py:5
P25519 = (1 << 255) - 19                  # fields.py:6

# status codes shared with include/fourq_b200.h
ST_OK = 0
ST_RESERVED_BIT = 1      # curve4q.py:52-53   bit 127 of y0 set
ST_NONCANONICAL = 2      # curve4q.py:61-62   y0 == p or y1 == p
ST_QUIRK_T0 = 3          # curve4q.py:76-77   t == 0 -> AttributeError in the reference
ST_NOT_ON_CURVE = 4      # curve4q.py:93-94, 447-448
ST_NEUTRAL = 5           # curve4q.py:459-460

MESSAGES = {
    ST_RESERVED_BIT: "Malformed point: reserved bit is not zero",
    ST_NONCANONICAL: "Malformed point: reserved bit is not zero",      # sic, curve4q.py:62
    ST_QUIRK_T0: "type object 'GFp' has no attribute 'two'",
    ST_NOT_ON_CURVE: "Point not on curve",
    ST_NEUTRAL: "DH computation resulted in neutral point",
}

# ------------------------------------------------------------------ GF(p), p = 2^127 - 1


def fp_add(x, y):            # fields.py:30-33
    return (x + y) % P127


def fp_sub(x, y):            # fields.py:36-39
    return (x - y) % P127


def fp_mul(x, y):            # fields.py:42-45
    return (x * y) % P127


def fp_sqr(x):               # fields.py:48-51
    return (x * x) % P127


def fp_neg(x):               # fields.py:54-57
    return (P127 - x) % P127


def _fp_nsqr(x, n):
    for _ in range(n):
        x = fp_sqr(x)
    return x


def fp_inv(x):
    """x^(2^127-3), the fixed chain of fields.py:67-106 (126 S + 12 M); inv(0) = 0."""
    x3 = fp_mul(x, fp_sqr(x))                       # 2^2 - 1
    xf = fp_mul(x3, _fp_nsqr(x3, 2))                # 2^4 - 1
    x8 = fp_mul(xf, _fp_nsqr(xf, 4))                # 2^8 - 1
    x16 = fp_mul(x8, _fp_nsqr(x8, 8))               # 2^16 - 1
    x32 = fp_mul(x16, _fp_nsqr(x16, 16))            # 2^32 - 1
    t = fp_mul(_fp_nsqr(x32, 32), x32)              # 2^64 - 1
    t = fp_mul(_fp_nsqr(t, 32), x32)                # 2^96 - 1
    t = fp_mul(_fp_nsqr(t, 16), x16)                # 2^112 - 1
    t = fp_mul(_fp_nsqr(t, 8), x8)                  # 2^120 - 1
    t = fp_mul(_fp_nsqr(t, 4), xf)                  # 2^124 - 1
    t = fp_mul(fp_sqr(t), x)                        # 2^125 - 1
    return fp_mul(_fp_nsqr(t, 2), x)                # 2^127 - 3


def fp_invsqrt(x):
    """x^(2^125-1) = x^((p-3)/4), fields.py:110-122 (5-bit sliding chain)."""
    x2 = fp_mul(x, x)
    x5 = fp_mul(fp_mul(x2, x2), x)
    x15 = fp_mul(fp_mul(x5, x5), x5)
    run = fp_mul(fp_mul(x15, x15), x)               # 2^5 - 1
    acc = run
    for _ in range(24):
        run = _fp_nsqr(run, 5)
        acc = fp_mul(run, acc)
    return acc


def fp_to_le(x):             # fields.py:125-126
    return (x % (1 << 128)).to_bytes(16, "little")


def fp_from_le(b):           # fields.py:129-132 (clears bit 127; does NOT mutate here)
    return int.from_bytes(bytes(b), "little") & P127


# ------------------------------------------------------------------ GF(p^2) = GF(p)[i]/(i^2+1)

F2_ZERO, F2_ONE = (0, 0), (1, 0)


def f2_add(a, b):            # fields.py:157-159
    return ((a[0] + b[0]) % P127, (a[1] + b[1]) % P127)


def f2_sub(a, b):            # fields.py:162-164
    return ((a[0] - b[0]) % P127, (a[1] - b[1]) % P127)


def f2_mul(a, b):            # fields.py:167-173
    return ((a[0] * b[0] - a[1] * b[1]) % P127, (a[0] * b[1] + a[1] * b[0]) % P127)


def f2_sqr(a):               # fields.py:176-181
    return ((a[0] * a[0] - a[1] * a[1]) % P127, (2 * a[0] * a[1]) % P127)


def f2_neg(a):               # fields.py:184-186
    return ((P127 - a[0]) % P127, (P127 - a[1]) % P127)


def f2_conj(a):              # fields.py:189-191
    return (a[0], (P127 - a[1]) % P127)


def f2_inv(a):               # fields.py:194-199   conj(a) / (a0^2 + a1^2); inv((0,0)) = (0,0)
    n = fp_inv(fp_add(fp_sqr(a[0]), fp_sqr(a[1])))
    return f2_mul((n, 0), f2_conj(a))


def f2_invsqrt(a):           # fields.py:201-230   the reference's control flow; its two `== -1` tests can never hold
    """a is taken as given (the raw a[1] == 0 test of :204 sees unreduced values); every field op reduces."""
    if a[1] == 0:                                             # :204-209
        t = fp_invsqrt(a[0])
        return (t, 0) if fp_mul(a[0], fp_sqr(t)) == 1 else (0, t)
    n = fp_add(fp_sqr(a[0]), fp_sqr(a[1]))                    # :214
    s = fp_invsqrt(n)                                         # :215
    c = fp_mul(n, s)                                          # :216  (the 'not square' exception of :217-218 is unreachable)
    half = 1 << 126                                           # GFp.half
    delta = fp_mul(fp_add(a[0], c), half)                     # :220
    g = fp_invsqrt(delta)                                     # :221
    h = fp_mul(delta, g)                                      # :222  (the second delta of :223-226 is unreachable)
    return (fp_mul(h, s), fp_neg(fp_mul(fp_mul(fp_mul(a[1], s), g), half)) % P127)       # :228-230


def select_raw(c, x, y, bits=128):
    """fields.py:59-64  y ^ ((mask * c) & (x ^ y)) with mask = 2^512 - 1, on `bits`-bit values: c = 1 -> x, c = 0 -> y, no reduction."""
    return y ^ ((((1 << 512) - 1) * c) & (x ^ y)) & ((1 << bits) - 1)


# ------------------------------------------------------------------ Curve4Q constants (curve4q.py:8-20)

D = (0xe40000000000000142, 0x5e472f846657e0fcb3821488f1fc0c8d)
N = 0x29cbc14e5e0a72f05397829cbc14e5dfbd004dfe0f79992fb2540ec7768ce7
GX = (0x1A3472237C2FB305286592AD7B3833AA, 0x1E1F553F2878AA9C96869FB360AC77F6)
GY = (0x0E3FEE9BA120785AB924A2462BCBB287, 0x6E1C4AF8630E024249A7C344844C8B5C)
NEUTRAL = ((0, 0), (1, 0))


def on_curve(P):             # curve4q.py:23-29    -x^2 + y^2 == 1 + d x^2 y^2
    x2, y2 = f2_sqr(P[0]), f2_sqr(P[1])
    return f2_sub(y2, x2) == f2_add(F2_ONE, f2_mul(f2_mul(D, x2), y2))


def sign(x):                 # curve4q.py:33-39
    return (x[0] >> 126) if x[0] != 0 else (x[1] >> 126)


def encode(x, y):            # curve4q.py:41-46
    out = bytearray(fp_to_le(y[0]) + fp_to_le(y[1]))
    out[31] |= sign(x) << 7
    return bytes(out)


def decode_status(B, spec=False):
    """curve4q.py:49-96 -> (status, (x, y) or None).  Length must be 32 (checked by the caller).
    spec=True follows the draft where the reference crashes: t = 2 (t0 - t3) when t == 0 (draft-ladd-cfrg-4q.md:865-867)."""
    B = bytes(B)
    if len(B) != 32:
        raise ValueError("Malformed point: length {} != 32".format(len(B)))   # curve4q.py:50-51
    if B[15] & 0x80:                                                          # :52
        return ST_RESERVED_BIT, None
    s = B[31] >> 7                                                            # :55
    y0 = fp_from_le(B[:16])                                                   # :58
    y1 = fp_from_le(B[16:])                                                   # :59 (drops the sign bit)
    if y0 >= P127 or y1 >= P127:                                              # :61
        return ST_NONCANONICAL, None
    y = (y0, y1)
    y2 = f2_sqr(y)
    u0, u1 = f2_sub(y2, F2_ONE)                                               # :66
    v0, v1 = f2_add(f2_mul(D, y2), F2_ONE)                                    # :67
    t0 = fp_add(fp_mul(u0, v0), fp_mul(u1, v1))                               # :69  Re(u conj v)
    t1 = fp_sub(fp_mul(u1, v0), fp_mul(u0, v1))                               # :70  Im(u conj v)
    t2 = fp_add(fp_sqr(v0), fp_sqr(v1))                                       # :71  |v|^2
    t3 = fp_add(fp_sqr(t0), fp_sqr(t1))                                       # :72
    t3 = fp_mul(fp_invsqrt(t3), t3)                                           # :73  sqrt(|u conj v|^2)
    t = fp_mul(2, fp_add(t0, t3))                                             # :75
    if t == 0:                                                                # :76-77 reference crashes here
        if not spec:
            return ST_QUIRK_T0, None
        t = fp_mul(2, fp_sub(t0, t3))                                         # what :77 means (draft :865-867)
    a = fp_invsqrt(fp_mul(t, fp_mul(t2, fp_sqr(t2))))                         # :79
    b = fp_mul(fp_mul(a, t2), t)                                              # :80
    x0 = fp_mul(b, 1 << 126)                                                  # :82  (GFp.half)
    x1 = fp_mul(fp_mul(a, t2), t1)                                            # :83
    if t != fp_mul(t2, fp_sqr(b)):                                            # :84
        x0, x1 = x1, x0
    x = (x0, x1)
    if sign(x) != s:                                                          # :88
        x = f2_neg(x)
    if not on_curve((x, y)):                                                  # :91
        x = f2_conj(x)
    if not on_curve((x, y)):                                                  # :93
        return ST_NOT_ON_CURVE, None
    return ST_OK, (x, y)


def decode(B):
    st, P = decode_status(B)
    if st == ST_QUIRK_T0:
        raise AttributeError(MESSAGES[st])
    if st != ST_OK:
        raise Exception(MESSAGES[st])
    return P


# ------------------------------------------------------------------ representations and group law


def affine_to_r1(x, y):      # curve4q.py:100-101
    return (x, y, F2_ONE, x, y)


def r1_to_affine(P):         # curve4q.py:103-106
    zi = f2_inv(P[2])
    return (f2_mul(P[0], zi), f2_mul(P[1], zi))


_TWO_D = f2_mul((2, 0), D)


def r1_to_r2(P):             # curve4q.py:109-116   (X+Y, Y-X, 2Z, 2d Ta Tb)
    X, Y, Z, Ta, Tb = P
    return (f2_add(X, Y), f2_sub(Y, X), f2_add(Z, Z), f2_mul(_TWO_D, f2_mul(Ta, Tb)))


def r1_to_r3(P):             # curve4q.py:119-126   (X+Y, Y-X, Z, Ta Tb)
    X, Y, Z, Ta, Tb = P
    return (f2_add(X, Y), f2_sub(Y, X), Z, f2_mul(Ta, Tb))


def r2_to_r4(P):             # curve4q.py:129-135   (N-D, D+N, E)
    return (f2_sub(P[0], P[1]), f2_add(P[1], P[0]), P[2])


def r2_neg(P):               # curve4q.py:193-195
    return (P[1], P[0], P[2], f2_neg(P[3]))


def dbl(P):                  # curve4q.py:138-152   R1/R4 -> R1
    X, Y, Z = P[0], P[1], P[2]
    A, B = f2_sqr(X), f2_sqr(Y)
    C = f2_mul((2, 0), f2_sqr(Z))
    Dd = f2_add(A, B)
    E = f2_sub(f2_sqr(f2_add(X, Y)), Dd)
    F = f2_sub(B, A)
    G = f2_sub(C, F)
    return (f2_mul(E, G), f2_mul(Dd, F), f2_mul(F, G), E, Dd)


def add_core(P, Q):          # curve4q.py:155-171   R3 + R2 -> R1
    N1, D1, E1, F1 = P
    N2, D2, Z2, T2 = Q
    A, B = f2_mul(D1, D2), f2_mul(N1, N2)
    C, Dd = f2_mul(T2, F1), f2_mul(Z2, E1)
    E, F, G, H = f2_sub(B, A), f2_sub(Dd, C), f2_add(Dd, C), f2_add(B, A)
    return (f2_mul(E, F), f2_mul(G, H), f2_mul(F, G), E, H)


def add(P, Q):               # curve4q.py:174-175   R1 + R2 -> R1
    return add_core(r1_to_r3(P), Q)


# ------------------------------------------------------------------ fixed-window scalar multiplication


def table_windowed(P):       # curve4q.py:179-185   T[i] = [2i+1]P in R2
    twoP = dbl(P)
    T = [r1_to_r2(P)]
    for _ in range(7):
        T.append(r1_to_r2(add(twoP, T[-1])))
    return T


def recode_windowed(m):
    """Digits of curve4q.py:216-226: returns (ind[63], sgn[63]); sgn 1 = positive, 0 = negative."""
    r = m % N
    if r % 2 == 0:
        r += N
    digs = []
    for _ in range(63):
        di = (r % 32) - 16
        digs.append(di)
        r = (r - di) // 16
    digs[62] = r                                                              # :223
    return [(abs(x) - 1) // 2 for x in digs], [1 if x > 0 else 0 for x in digs]


def mul_windowed(m, P, table=None):     # curve4q.py:188-235   R1 -> R1
    T = table if table else table_windowed(P)
    ind, sgn = recode_windowed(m)

    def pick(i):
        return T[ind[i]] if sgn[i] else r2_neg(T[ind[i]])
    Q = r2_to_r4(pick(62))
    for i in range(61, -1, -1):
        Q = dbl(dbl(dbl(dbl(Q))))
        Q = add(Q, pick(i))
    return Q


# ------------------------------------------------------------------ endomorphisms (curve4q.py:240-322)

CTAU = (0x1964de2c3afad20c74dcd57cebce74c3, 0x000000000000000c0000000000000012)
CTAUDUAL = (0x4aa740eb230586529ecaa6d9decdf034, 0x7ffffffffffffff40000000000000011)
CPHI = [
    (0x0000000000000005fffffffffffffff7, 0x2553a0759182c3294f65536cef66f81a),
    (0x00000000000000050000000000000007, 0x62c8caa0c50c62cf334d90e9e28296f9),
    (0x000000000000000f0000000000000015, 0x78df262b6c9b5c982c2cb7154f1df391),
    (0x00000000000000020000000000000003, 0x5084c6491d76342a92440457a7962ea4),
    (0x00000000000000030000000000000003, 0x12440457a7962ea4a1098c923aec6855),
    (0x000000000000000a000000000000000f, 0x459195418a18c59e669b21d3c5052df3),
    (0x00000000000000120000000000000018, 0x0b232a8314318b3ccd3643a78a0a5be7),
    (0x00000000000000180000000000000023, 0x3963bc1c99e2ea1a66c183035f48781a),
    (0x00000000000000aa00000000000000f0, 0x1f529f860316cbe544e251582b5d0ef0),
    (0x00000000000008700000000000000bef, 0x0fd52e9cfe00375b014d3e48976e2505),
]
CPSI = [
    None,
    (0x2af99e9a83d54a02edf07f4767e346ef, 0x00000000000000de000000000000013a),
    (0x00000000000000e40000000000000143, 0x21b8d07b99a81f034c7deb770e03f372),
    (0x00000000000000060000000000000009, 0x4cb26f161d7d69063a6e6abe75e73a61),
    (0x7ffffffffffffff9fffffffffffffff6, 0x334d90e9e28296f9c59195418a18c59e),
]


def _two_sqr(z):
    s = f2_sqr(z)
    return f2_add(s, s)


def tau(P):                  # curve4q.py:258-267
    X, Y, Z = P
    A, B = f2_sqr(X), f2_sqr(Y)
    C, Dd = f2_add(A, B), f2_sub(A, B)
    return (f2_mul(f2_mul(f2_mul(CTAU, X), Y), Dd),
            f2_neg(f2_mul(f2_add(_two_sqr(Z), Dd), C)),
            f2_mul(C, Dd))


def tau_dual(P):             # curve4q.py:269-280   -> R1
    X, Y, Z = P
    A, B = f2_sqr(X), f2_sqr(Y)
    C = f2_add(A, B)
    Ta = f2_sub(B, A)
    Dd = f2_sub(_two_sqr(Z), Ta)
    Tb = f2_mul(f2_mul(CTAUDUAL, X), Y)
    return (f2_mul(Tb, C), f2_mul(Ta, Dd), f2_mul(C, Dd), Ta, Tb)


def upsilon(P):              # curve4q.py:282-302
    X, Y, Z = P
    c = CPHI
    A = f2_mul(f2_mul(c[0], X), Y)
    B = f2_mul(Y, Z)
    C, Dd = f2_sqr(Y), f2_sqr(Z)
    F, G, H = f2_sqr(Dd), f2_sqr(B), f2_sqr(C)
    I = f2_mul(c[1], B)
    J = f2_add(C, f2_mul(c[2], Dd))
    K = f2_add(f2_add(f2_mul(c[8], G), H), f2_mul(c[9], F))
    X2 = f2_conj(f2_mul(f2_mul(A, K), f2_mul(f2_add(I, J), f2_sub(I, J))))
    L = f2_add(C, f2_mul(c[4], Dd))
    M = f2_mul(c[3], B)
    Nn = f2_mul(f2_add(L, M), f2_sub(L, M))
    Y2 = f2_add(f2_add(H, f2_mul(c[6], G)), f2_mul(c[7], F))
    Y2 = f2_conj(f2_mul(f2_mul(f2_mul(c[5], Dd), Nn), Y2))
    Z2 = f2_conj(f2_mul(f2_mul(B, K), Nn))
    return (X2, Y2, Z2)


def chi(P):                  # curve4q.py:304-316
    X, Y, Z = P
    A, B = f2_conj(X), f2_conj(Y)
    C = f2_sqr(f2_conj(Z))
    Dd, F = f2_sqr(A), f2_sqr(B)
    G = f2_mul(B, f2_add(Dd, f2_mul(CPSI[2], C)))
    H = f2_neg(f2_add(Dd, f2_mul(CPSI[4], C)))
    return (f2_mul(f2_mul(f2_mul(CPSI[1], A), C), H),
            f2_mul(G, f2_add(Dd, f2_mul(CPSI[3], C))),
            f2_mul(G, H))


def phi(P):                  # curve4q.py:318-319
    return tau_dual(upsilon(tau(P[:3])))


def psi(P):                  # curve4q.py:321-322
    return tau_dual(chi(tau(P[:3])))


# ------------------------------------------------------------------ decomposition + GLV-SAC recoding

B1 = [0x0906ff27e0a0a196, -0x1363e862c22a2da0, 0x07426031ecc8030f, -0x084f739986b9e651]
B2 = [0x1d495bea84fcc2d4, -0x0000000000000001, 0x0000000000000001, 0x25dbc5bc8dd167d0]
B3 = [0x17abad1d231f0302, 0x02c4211ae388da51, -0x2e4d21c98927c49f, 0x0a9e6f44c02ecd97]
B4 = [0x136e340a9108c83f, 0x3122df2dc3e0ff32, -0x068a49f02aa8a9b5, -0x18d5087896de0aea]
ELL = [0x7fc5bb5c5ea2be5dff75682ace6a6bd66259686e09d1a7d4f,
       0x38fd4b04caa6c0f8a2bd235580f468d8dd1ba1d84dd627afb,
       0x0d038bf8d0bffbaf6c42bd6c965dca9029b291a33678c203c,
       0x31b073877a22d841081cbdc3714983d8212e5666b77e7fdc0]
OFFS_C = [5 * B2[i] - 3 * B3[i] + 2 * B4[i] for i in range(4)]       # curve4q.py:336
OFFS_CP = [OFFS_C[i] + B4[i] for i in range(4)]                      # curve4q.py:337


def decompose(m):            # curve4q.py:339-356   m in [0,2^256) -> four 64-bit scalars, a1 odd
    t = [(L * m) >> 256 for L in ELL]
    a = [(m if i == 0 else 0) - t[0] * B1[i] - t[1] * B2[i] - t[2] * B3[i] - t[3] * B4[i] for i in range(4)]
    ac = [a[i] + OFFS_C[i] for i in range(4)]
    acp = [a[i] + OFFS_CP[i] for i in range(4)]
    return ac if (ac[0] & 1) else acp


def recode_endo(v):          # curve4q.py:358-380   -> (sign bits m[65], digits d[65])
    vv = list(v)
    digits, signs = [0] * 65, [0] * 65
    for i in range(64):
        b1 = (vv[0] >> (i + 1)) & 1
        signs[i] = b1
        for j in (1, 2, 3):
            bj = vv[j] & 1
            digits[i] += bj << (j - 1)
            vv[j] = (vv[j] >> 1) + ((b1 | bj) ^ b1)
    digits[64] = vv[1] + 2 * vv[2] + 4 * vv[3]
    signs[64] = 1
    return signs, digits


def table_endo(P):           # curve4q.py:385-403
    Q = phi(P)
    R = psi(P)
    S = psi(Q)
    Q, R, S = r1_to_r3(Q), r1_to_r3(R), r1_to_r3(S)
    T = [None] * 8
    T[0] = r1_to_r2(P)
    T[1] = r1_to_r2(add_core(Q, T[0]))
    T[2] = r1_to_r2(add_core(R, T[0]))
    T[3] = r1_to_r2(add_core(R, T[1]))
    for i in range(4):
        T[4 + i] = r1_to_r2(add_core(S, T[i]))
    return T


def mul_endo(m, P, table=None):         # curve4q.py:405-442
    T = table if table else table_endo(P)
    s, dg = recode_endo(decompose(m))

    def pick(i):
        return T[dg[i]] if s[i] else r2_neg(T[dg[i]])
    Q = r2_to_r4(pick(64))
    for i in range(63, -1, -1):
        Q = add(dbl(Q), pick(i))
    return Q


# ------------------------------------------------------------------ Diffie-Hellman (curve4q.py:446-468)


def clear_cofactor(P):       # curve4q.py:450-455   [392]P: DBL, ADD, 4 DBL, ADD, 3 DBL
    P0 = affine_to_r1(P[0], P[1])
    base = r1_to_r2(P0)
    Q = add(dbl(P0), base)                      # 3P
    Q = dbl(dbl(dbl(dbl(Q))))                   # 48P
    Q = add(Q, base)                            # 49P
    return dbl(dbl(dbl(Q)))                     # 392P


def dh_status(m, P, mul=mul_windowed, table=None):
    """DH_core as (status, affine or None)."""
    if not on_curve(P):                                                       # :447
        return ST_NOT_ON_CURVE, None
    Q = r1_to_affine(mul(m, clear_cofactor(P), table=table))                  # :457
    if Q == NEUTRAL:                                                          # :459
        return ST_NEUTRAL, None
    return ST_OK, Q


def dh_windowed(m, P, table=None):
    st, Q = dh_status(m, P, mul_windowed, table)
    if st != ST_OK:
        raise Exception(MESSAGES[st])
    return Q


def dh_endo(m, P, table=None):
    st, Q = dh_status(m, P, mul_endo, table)
    if st != ST_OK:
        raise Exception(MESSAGES[st])
    return Q


# ------------------------------------------------------------------ byte-level row functions
# These are what the batched C-ABI entry points compute for one row (include/fourq_b200.h).

ZERO32 = bytes(32)
_G_R1 = affine_to_r1(GX, GY)
_cache = {}


def _table_G():
    if "G" not in _cache:
        _cache["G"] = table_windowed(_G_R1)
    return _cache["G"]


def _table_G392():
    if "G392" not in _cache:
        _cache["G392"] = table_windowed(mul_windowed(392, _G_R1))
    return _cache["G392"]


def le_scalar(k):
    return int.from_bytes(bytes(k), "little")


def row_dh(k, enc_pt, mul=mul_windowed):
    """fq_dh: encode(DH_windowed(k, decode(enc_pt))) -> (32 bytes, status); failed rows are zeros."""
    st, P = decode_status(enc_pt)
    if st != ST_OK:
        return ZERO32, st
    st, Q = dh_status(le_scalar(k), P, mul)
    if st != ST_OK:
        return ZERO32, st
    return encode(Q[0], Q[1]), ST_OK


def row_dh_affine(k, xy, mul=mul_windowed):
    """fq_dh_affine: DH_windowed on an affine point given as 64 bytes x0|x1|y0|y1 -> 64 bytes."""
    P = xy_from_bytes(xy)
    st, Q = dh_status(le_scalar(k), P, mul)
    if st != ST_OK:
        return bytes(64), st
    return xy_to_bytes(Q), ST_OK


def row_dh_base(k):
    """fq_dh_base: encode(DH_windowed(k, G, table=T392)) = [392 k]G  (curve4q.py:743-762)."""
    st, Q = dh_status(le_scalar(k), (GX, GY), mul_windowed, _table_G392())
    if st != ST_OK:
        return ZERO32, st
    return encode(Q[0], Q[1]), ST_OK


def row_mul_base(k):
    """fq_mul_base: encode(R1toAffine(MUL_windowed(k, G, table=table_windowed(G)))) = [k]G
    (curve4q.py:582-584).  [k]G = neutral is NOT an error here (MUL_* has no such check)."""
    Q = r1_to_affine(mul_windowed(le_scalar(k), _G_R1, _table_G()))
    return encode(Q[0], Q[1])


def row_decode(enc, spec=False):
    st, P = decode_status(enc, spec)
    if st != ST_OK:
        return bytes(64), st
    return xy_to_bytes(P), ST_OK


def row_encode(xy):
    P = xy_from_bytes(xy)
    return encode(P[0], P[1])


def xy_to_bytes(P):
    return fp_to_le(P[0][0]) + fp_to_le(P[0][1]) + fp_to_le(P[1][0]) + fp_to_le(P[1][1])


def _u128(b):
    return int.from_bytes(bytes(b), "little")


def xy_from_bytes(b):
    """Affine 64-byte rows carry four 128-bit LE integers, taken mod p like the reference's ints."""
    b = bytes(b)
    return ((_u128(b[0:16]) % P127, _u128(b[16:32]) % P127), (_u128(b[32:48]) % P127, _u128(b[48:64]) % P127))


def f2_from_bytes(b):
    """A GF(p^2) row: two 128-bit LE integers, any value (the reference's ops reduce mod p)."""
    b = bytes(b)
    return (_u128(b[:16]), _u128(b[16:]))


def f2_to_bytes(a):
    return fp_to_le(a[0] % P127) + fp_to_le(a[1] % P127)


def row_fp2(op, a, b=None):
    A = f2_from_bytes(a)
    if op == "mul":
        return f2_to_bytes(f2_mul(A, f2_from_bytes(b)))
    if op == "add":
        return f2_to_bytes(f2_add(A, f2_from_bytes(b)))
    if op == "sub":
        return f2_to_bytes(f2_sub(A, f2_from_bytes(b)))
    if op == "sqr":
        return f2_to_bytes(f2_sqr(A))
    if op == "inv":
        return f2_to_bytes(f2_inv((A[0] % P127, A[1] % P127)))
    if op == "neg":
        return f2_to_bytes(f2_neg((A[0] % P127, A[1] % P127)))
    if op == "conj":
        return f2_to_bytes(f2_conj((A[0] % P127, A[1] % P127)))
    if op == "invsqrt":
        return f2_to_bytes(f2_invsqrt(A))
    raise ValueError(op)


def row_select(c, x, y):
    """GFp.select / GFp2.select (fields.py:59-64, :236-238) on 16- or 32-byte rows with the condition byte c."""
    out = b""
    for h in range(len(x) // 16):
        xi, yi = int.from_bytes(x[16 * h:16 * h + 16], "little"), int.from_bytes(y[16 * h:16 * h + 16], "little")
        out += select_raw(c, xi, yi).to_bytes(16, "little")
    return out


def row_fp(op, a, b=None):
    """GF(p) op on 16-byte little-endian rows; inputs are arbitrary 128-bit ints as in the reference (fields.py:29-122)."""
    x = int.from_bytes(a, "little")
    y = int.from_bytes(b, "little") if b is not None else None
    if op == "add":
        r = fp_add(x, y)
    elif op == "sub":
        r = fp_sub(x, y)
    elif op == "mul":
        r = fp_mul(x, y)
    elif op == "sqr":
        r = fp_sqr(x)
    elif op == "neg":
        r = fp_neg(x % P127) % P127
    elif op == "inv":
        r = fp_inv(x % P127)
    elif op == "invsqrt":
        r = fp_invsqrt(x % P127)
    else:
        raise ValueError(op)
    return fp_to_le(r)


# ------------------------------------------------------------------ X25519 (curve25519.py:17-91)


def x25519_scalar(k):        # curve25519.py:20-25
    b = bytearray(k)
    b[0] &= 248
    b[31] &= 127
    b[31] |= 64
    return int.from_bytes(b, "little")


def x25519_ucoord(u):        # curve25519.py:27-33
    b = bytearray(u)
    b[31] &= 127
    return int.from_bytes(b, "little")


def fp25519_inv(z):          # fields.py:293-362 computes z^(p-2); any chain gives the same value
    return pow(z, P25519 - 2, P25519)


def row_f25519(op, a, b=None):
    """GFp25519.add/sub/mul/sqr/inv (fields.py:267-362) on 32-byte little-endian rows; inputs are arbitrary 256-bit ints as in the
    reference (it reduces mod p), the output is the canonical value."""
    x = int.from_bytes(a, "little")
    y = int.from_bytes(b, "little") if b is not None else None
    if op == "add":
        r = (x + y) % P25519
    elif op == "sub":
        r = (x - y) % P25519
    elif op == "mul":
        r = (x * y) % P25519
    elif op == "sqr":
        r = (x * x) % P25519
    elif op == "inv":
        r = fp25519_inv(x)
    else:
        raise ValueError(op)
    return r.to_bytes(32, "little")


def x25519_ladder(k, u):     # curve25519.py:43-80 with bits = 255, a24 = 121665
    p = P25519
    x1, x2, z2, x3, z3, swap = u, 1, 0, u, 1, 0
    for t in range(254, -1, -1):
        kt = (k >> t) & 1
        if swap ^ kt:
            x2, x3, z2, z3 = x3, x2, z3, z2
        swap = kt
        A, B = (x2 + z2) % p, (x2 - z2) % p
        AA, BB = A * A % p, B * B % p
        E = (AA - BB) % p
        C, Dd = (x3 + z3) % p, (x3 - z3) % p
        DA, CB = Dd * A % p, C * B % p
        x3 = (DA + CB) ** 2 % p
        z3 = x1 * ((DA - CB) ** 2 % p) % p
        x2 = AA * BB % p
        z2 = E * ((AA + 121665 * E) % p) % p
    if swap:
        x2, z2 = x3, z3
    return x2 * fp25519_inv(z2) % p


def x25519(k, u):            # curve25519.py:88-91
    return (x25519_ladder(x25519_scalar(k), x25519_ucoord(u)) % P25519).to_bytes(32, "little")
