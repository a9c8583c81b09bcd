/* fourq_oracle.c -- CPU oracle for the Curve4Q hot path in plain C.  TEST INFRASTRUCTURE, NOT PRODUCT.
 *
 * A restatement of the reference's algorithms (bifurcation/fourq, impl/fields.py and impl/curve4q.py) on unsigned
 * __int128, fast enough (~0.1 ms per Diffie-Hellman) to check full BASELINE-size batches bit for bit.  Every function
 * cites the reference lines it follows.  Built by oracle/Makefile into oracle/_build/libfourq_oracle.so.
 *
 * Parity status: PINNED.  tests/test_oracle_c.py checks this library against tests/golden/{fields,fp,codec,mul}.json (produced
 * by the reference's own code, tests/golden/gen_golden.py) and against oracle/fourq_oracle.py on random rows.
 *
 * Only tests/ (and, as the checker only, bench.py / __graft_entry__.smoke()) may load this library; fourq_b200/ never does.
 * Status codes: include/fourq_b200.h (FQ_ST_*).  Failed rows are zero-filled, as in the product and the Python oracle.
 */
#include <stddef.h>
#include <stdint.h>
#include <string.h>

typedef unsigned __int128 u128;
typedef uint64_t u64;

#define ST_OK 0
#define ST_RESERVED_BIT 1
#define ST_NONCANONICAL 2
#define ST_QUIRK_T0 3
#define ST_NOT_ON_CURVE 4
#define ST_NEUTRAL 5

static const u128 P = (((u128)1) << 127) - 1;                       /* fields.py:5 */
#define U128(hi, lo) ((((u128)(hi)) << 64) | (u128)(lo))

/* ------------------------------------------------------------------ GF(p), fields.py:29-132 (inputs and outputs < p) */
static u128 fp_red(u128 x) { x = (x & P) + (x >> 127); if (x >= P) x -= P; return x; }     /* any 128-bit value mod p */
static u128 fp_add(u128 x, u128 y) { u128 s = x + y; if (s >= P) s -= P; return s; }       /* :30-33 */
static u128 fp_sub(u128 x, u128 y) { return x >= y ? x - y : x + P - y; }                  /* :36-39 */
static u128 fp_neg(u128 x) { return x == 0 ? 0 : P - x; }                                  /* :54-57 */
static u128 fp_mul(u128 x, u128 y) {                                                       /* :42-45 */
  u64 x0 = (u64)x, x1 = (u64)(x >> 64), y0 = (u64)y, y1 = (u64)(y >> 64);
  u128 p00 = (u128)x0 * y0, p01 = (u128)x0 * y1, p10 = (u128)x1 * y0, p11 = (u128)x1 * y1;
  /* product = p11 2^128 + (p01 + p10) 2^64 + p00;  2^128 = 2 (mod p) */
  u128 mid = p01 + p10;                        /* < 2^127: x1, y1 < 2^63 */
  u128 lo = p00 + (mid << 64);
  u128 carry = lo < p00;
  u128 hi = p11 + (mid >> 64) + carry;         /* < 2^126 */
  return fp_add(fp_red(lo), fp_red(hi << 1));
}
static u128 fp_sqr(u128 x) { return fp_mul(x, x); }                                        /* :48-51 */
static u128 fp_nsqr(u128 x, int n) { while (n--) x = fp_sqr(x); return x; }
/* :67-106, x^(2^127 - 3) by the reference's chain */
static u128 fp_inv(u128 x) {
  u128 x3 = fp_mul(fp_sqr(x), x), xf = fp_mul(fp_nsqr(x3, 2), x3), x8 = fp_mul(fp_nsqr(xf, 4), xf);
  u128 x16 = fp_mul(fp_nsqr(x8, 8), x8), x32 = fp_mul(fp_nsqr(x16, 16), x16);
  u128 t = fp_mul(fp_nsqr(x32, 32), x32);
  t = fp_mul(fp_nsqr(t, 32), x32); t = fp_mul(fp_nsqr(t, 16), x16); t = fp_mul(fp_nsqr(t, 8), x8);
  t = fp_mul(fp_nsqr(t, 4), xf); t = fp_mul(fp_sqr(t), x);
  return fp_mul(fp_nsqr(t, 2), x);
}
/* :110-122, x^(2^125 - 1) */
static u128 fp_invsqrt(u128 x) {
  u128 r = x;
  for (int i = 0; i < 124; i++) r = fp_mul(fp_sqr(r), x);
  return r;
}

/* ------------------------------------------------------------------ GF(p^2), fields.py:134-238 */
typedef struct { u128 re, im; } f2;
static const f2 F2_ZERO = {0, 0}, F2_ONE = {1, 0};
static f2 f2_add(f2 a, f2 b) { f2 r = {fp_add(a.re, b.re), fp_add(a.im, b.im)}; return r; }       /* :157 */
static f2 f2_sub(f2 a, f2 b) { f2 r = {fp_sub(a.re, b.re), fp_sub(a.im, b.im)}; return r; }       /* :162 */
static f2 f2_mul(f2 a, f2 b) {                                                                     /* :167-173 */
  f2 r = {fp_sub(fp_mul(a.re, b.re), fp_mul(a.im, b.im)), fp_add(fp_mul(a.re, b.im), fp_mul(a.im, b.re))};
  return r;
}
static f2 f2_sqr(f2 a) {                                                                           /* :176-181 */
  f2 r = {fp_mul(fp_add(a.re, a.im), fp_sub(a.re, a.im)), fp_mul(fp_add(a.re, a.re), a.im)};
  return r;
}
static f2 f2_neg(f2 a) { f2 r = {fp_neg(a.re), fp_neg(a.im)}; return r; }                          /* :184 */
static f2 f2_conj(f2 a) { f2 r = {a.re, fp_neg(a.im)}; return r; }                                 /* :189 */
static f2 f2_inv(f2 a) {                                                                           /* :194-199 */
  u128 n = fp_inv(fp_add(fp_sqr(a.re), fp_sqr(a.im)));
  f2 r = {fp_mul(a.re, n), fp_mul(fp_neg(a.im), n)};
  return r;
}
static int f2_eq(f2 a, f2 b) { return a.re == b.re && a.im == b.im; }

/* ------------------------------------------------------------------ bytes */
static u128 ld128(const uint8_t* b) { u64 lo, hi; memcpy(&lo, b, 8); memcpy(&hi, b + 8, 8); return U128(hi, lo); }
static void st128(uint8_t* b, u128 x) { u64 lo = (u64)x, hi = (u64)(x >> 64); memcpy(b, &lo, 8); memcpy(b + 8, &hi, 8); }

/* ------------------------------------------------------------------ curve constants, curve4q.py:9-20 */
static const f2 CURVE_D = {U128(0x00000000000000e4ull, 0x0000000000000142ull), U128(0x5e472f846657e0fcull, 0xb3821488f1fc0c8dull)};
static const f2 GX = {U128(0x1A3472237C2FB305ull, 0x286592AD7B3833AAull), U128(0x1E1F553F2878AA9Cull, 0x96869FB360AC77F6ull)};
static const f2 GY = {U128(0x0E3FEE9BA120785Aull, 0xB924A2462BCBB287ull), U128(0x6E1C4AF8630E0242ull, 0x49A7C344844C8B5Cull)};
/* N, little-endian 64-bit limbs (curve4q.py:12) */
static const u64 CURVE_N[4] = {0x2fb2540ec7768ce7ull, 0xdfbd004dfe0f7999ull, 0xf05397829cbc14e5ull, 0x0029cbc14e5e0a72ull};

typedef struct { f2 x, y; } aff;
typedef struct { f2 X, Y, Z, Ta, Tb; } r1;
typedef struct { f2 N, D, E, F; } r2;            /* also used for R3 = (X+Y, Y-X, Z, T) */

/* curve4q.py:23-29 */
static int on_curve(f2 x, f2 y) {
  f2 x2 = f2_sqr(x), y2 = f2_sqr(y);
  return f2_eq(f2_sub(y2, x2), f2_add(F2_ONE, f2_mul(f2_mul(CURVE_D, x2), y2)));
}
static unsigned sign_of(f2 x) { return (unsigned)((x.re != 0 ? x.re : x.im) >> 126); }             /* :33-39 */
static void encode(f2 x, f2 y, uint8_t* out) {                                                      /* :41-46 */
  st128(out, y.re); st128(out + 16, y.im);
  out[31] |= (uint8_t)(sign_of(x) << 7);
}
/* curve4q.py:49-96 */
static int decode(const uint8_t* B, aff* out) {
  if (B[15] & 0x80) return ST_RESERVED_BIT;                                                        /* :52 */
  unsigned s = B[31] >> 7;                                                                         /* :55 */
  u128 y0 = ld128(B) & P, y1 = ld128(B + 16) & P;                                                  /* :58-59, fromLittleEndian clears bit 127 */
  if (y0 >= P || y1 >= P) return ST_NONCANONICAL;                                                  /* :61 */
  f2 y = {y0, y1};
  f2 y2 = f2_sqr(y);
  f2 u = f2_sub(y2, F2_ONE);                                                                       /* :66 */
  f2 v = f2_add(f2_mul(CURVE_D, y2), F2_ONE);                                                      /* :67 */
  u128 t0 = fp_add(fp_mul(u.re, v.re), fp_mul(u.im, v.im));                                        /* :69 */
  u128 t1 = fp_sub(fp_mul(u.im, v.re), fp_mul(u.re, v.im));                                        /* :70 */
  u128 t2 = fp_add(fp_sqr(v.re), fp_sqr(v.im));                                                    /* :71 */
  u128 t3 = fp_add(fp_sqr(t0), fp_sqr(t1));                                                        /* :72 */
  t3 = fp_mul(fp_invsqrt(t3), t3);                                                                 /* :73 */
  u128 t = fp_mul(2, fp_add(t0, t3));                                                              /* :75 */
  if (t == 0) return ST_QUIRK_T0;                                                                  /* :76-77: the reference raises */
  u128 a = fp_invsqrt(fp_mul(t, fp_mul(t2, fp_sqr(t2))));                                          /* :79 */
  u128 b = fp_mul(fp_mul(a, t2), t);                                                               /* :80 */
  u128 x0 = fp_mul(b, ((u128)1) << 126);                                                           /* :82 */
  u128 x1 = fp_mul(fp_mul(a, t2), t1);                                                             /* :83 */
  if (t != fp_mul(t2, fp_sqr(b))) { u128 w = x0; x0 = x1; x1 = w; }                                /* :84-85 */
  f2 x = {x0, x1};
  if (sign_of(x) != s) x = f2_neg(x);                                                              /* :88-89 */
  if (!on_curve(x, y)) x = f2_conj(x);                                                             /* :91-92 */
  if (!on_curve(x, y)) return ST_NOT_ON_CURVE;                                                     /* :93-94 */
  out->x = x; out->y = y;
  return ST_OK;
}

/* ------------------------------------------------------------------ representations and group law, curve4q.py:100-175 */
static r1 affine_to_r1(f2 x, f2 y) { r1 p = {x, y, F2_ONE, x, y}; return p; }                      /* :100 */
static aff r1_to_affine(r1 p) { f2 zi = f2_inv(p.Z); aff a = {f2_mul(p.X, zi), f2_mul(p.Y, zi)}; return a; }   /* :103-106 */
static r2 r1_to_r2(r1 p) {                                                                          /* :109-116 */
  f2 two = {2, 0};
  r2 r = {f2_add(p.X, p.Y), f2_sub(p.Y, p.X), f2_add(p.Z, p.Z), f2_mul(f2_mul(two, CURVE_D), f2_mul(p.Ta, p.Tb))};
  return r;
}
static r2 r1_to_r3(r1 p) { r2 r = {f2_add(p.X, p.Y), f2_sub(p.Y, p.X), p.Z, f2_mul(p.Ta, p.Tb)}; return r; }  /* :119-126 */
static r1 r2_to_r4(r2 p) { r1 r = {f2_sub(p.N, p.D), f2_add(p.D, p.N), p.E, F2_ZERO, F2_ZERO}; return r; }   /* :129-135 */
static r2 r2_neg(r2 p) { r2 r = {p.D, p.N, p.E, f2_neg(p.F)}; return r; }                          /* :193-195 */
static r1 dbl(r1 p) {                                                                               /* :138-152 */
  f2 two = {2, 0};
  f2 A = f2_sqr(p.X), B = f2_sqr(p.Y), C = f2_mul(two, f2_sqr(p.Z)), D = f2_add(A, B);
  f2 E = f2_sub(f2_sqr(f2_add(p.X, p.Y)), D), F = f2_sub(B, A), G = f2_sub(C, F);
  r1 r = {f2_mul(E, G), f2_mul(D, F), f2_mul(F, G), E, D};
  return r;
}
static r1 add_core(r2 p, r2 q) {                                                                    /* :155-171, p in R3, q in R2 */
  f2 A = f2_mul(p.D, q.D), B = f2_mul(p.N, q.N), C = f2_mul(q.F, p.F), D = f2_mul(q.E, p.E);
  f2 E = f2_sub(B, A), F = f2_sub(D, C), G = f2_add(D, C), H = f2_add(B, A);
  r1 r = {f2_mul(E, F), f2_mul(G, H), f2_mul(F, G), E, H};
  return r;
}
static r1 add(r1 p, r2 q) { return add_core(r1_to_r3(p), q); }                                     /* :174-175 */

/* ------------------------------------------------------------------ 256-bit scalars (little-endian 64-bit limbs) */
typedef struct { u64 v[5]; } sc;                 /* one spare limb for shifted copies of N */
static int sc_ge(const sc* a, const sc* b) {
  for (int i = 4; i >= 0; i--) { if (a->v[i] != b->v[i]) return a->v[i] > b->v[i]; }
  return 1;
}
static void sc_sub(sc* a, const sc* b) {
  u64 borrow = 0;
  for (int i = 0; i < 5; i++) { u128 d = (u128)a->v[i] - b->v[i] - borrow; a->v[i] = (u64)d; borrow = (u64)(d >> 64) & 1; }
}
static void sc_add(sc* a, const sc* b) {
  u64 carry = 0;
  for (int i = 0; i < 5; i++) { u128 s = (u128)a->v[i] + b->v[i] + carry; a->v[i] = (u64)s; carry = (u64)(s >> 64); }
}
static sc sc_shl(const sc* a, int k) {           /* 0 <= k < 64 */
  sc r;
  for (int i = 4; i >= 0; i--) r.v[i] = (a->v[i] << k) | ((k && i) ? a->v[i - 1] >> (64 - k) : 0);
  return r;
}
/* m mod N (curve4q.py:217) by shift-and-subtract: m < 2^256, N > 2^245 */
static sc sc_mod_n(sc m) {
  sc n = {{CURVE_N[0], CURVE_N[1], CURVE_N[2], CURVE_N[3], 0}};
  for (int k = 11; k >= 0; k--) { sc t = sc_shl(&n, k); if (sc_ge(&m, &t)) sc_sub(&m, &t); }
  return m;
}

/* ------------------------------------------------------------------ fixed-window scalar multiplication, curve4q.py:179-235 */
static void table_windowed(r1 p, r2 T[8]) {                                                         /* :179-185 */
  r1 q = dbl(p);
  T[0] = r1_to_r2(p);
  for (int i = 1; i < 8; i++) T[i] = r1_to_r2(add(q, T[i - 1]));
}
/* :216-226.  digits d[i] = (r mod 32) - 16, r = (r - d[i]) / 16, 63 times; then d[62] = r */
static void recode_windowed(const uint8_t* k, int ind[63], int sgn[63]) {
  sc m; memcpy(m.v, k, 32); m.v[4] = 0;
  sc r = sc_mod_n(m);
  if ((r.v[0] & 1) == 0) { sc n = {{CURVE_N[0], CURVE_N[1], CURVE_N[2], CURVE_N[3], 0}}; sc_add(&r, &n); }   /* :218-219 */
  int d[63];
  for (int i = 0; i < 63; i++) {
    int di = (int)(r.v[0] & 31) - 16;
    d[i] = di;
    /* r = (r - di) / 16: r is odd, di is odd, so r - di is a multiple of 2; the reference's integer division is exact */
    sc t = {{0, 0, 0, 0, 0}};
    if (di >= 0) { t.v[0] = (u64)di; sc_sub(&r, &t); } else { t.v[0] = (u64)(-di); sc_add(&r, &t); }
    for (int j = 0; j < 5; j++) r.v[j] = (r.v[j] >> 4) | (j < 4 ? r.v[j + 1] << 60 : 0);
  }
  d[62] = (int)r.v[0];                                                                             /* :223 (r is a small integer here) */
  for (int i = 0; i < 63; i++) { int a = d[i] < 0 ? -d[i] : d[i]; ind[i] = (a - 1) / 2; sgn[i] = d[i] > 0; }   /* :224-226 */
}
static r1 mul_windowed(const uint8_t* k, r1 p, const r2* table) {                                   /* :188-235 */
  r2 Tl[8];
  const r2* T = table;
  if (!T) { table_windowed(p, Tl); T = Tl; }                                                       /* :210-212 */
  int ind[63], sgn[63];
  recode_windowed(k, ind, sgn);
  r1 q = r2_to_r4(sgn[62] ? T[ind[62]] : r2_neg(T[ind[62]]));                                      /* :229 */
  for (int i = 61; i >= 0; i--) {
    q = dbl(dbl(dbl(dbl(q))));                                                                     /* :231 */
    q = add(q, sgn[i] ? T[ind[i]] : r2_neg(T[ind[i]]));                                            /* :232-233 */
  }
  return q;
}

/* ------------------------------------------------------------------ Diffie-Hellman, curve4q.py:446-465 */
static int dh_core(const uint8_t* k, aff P, aff* out) {
  if (!on_curve(P.x, P.y)) return ST_NOT_ON_CURVE;                                                 /* :447-448 */
  r1 q = affine_to_r1(P.x, P.y);
  r2 b = r1_to_r2(q);
  q = dbl(q); q = add(q, b);                                                                       /* :450-455: [392]P */
  q = dbl(dbl(dbl(dbl(q)))); q = add(q, b);
  q = dbl(dbl(dbl(q)));
  r1 r = mul_windowed(k, q, NULL);                                                                 /* :457 */
  aff a = r1_to_affine(r);
  if (f2_eq(a.x, F2_ZERO) && f2_eq(a.y, F2_ONE)) return ST_NEUTRAL;                                /* :459-460 */
  *out = a;
  return ST_OK;
}

/* ================================================================== exported (ctypes) */

/* fq_dh: encode(DH_windowed(k, decode(enc))) */
int fqo_dh(const uint8_t* k, const uint8_t* enc, uint8_t* out) {
  aff P, Q;
  memset(out, 0, 32);
  int st = decode(enc, &P);
  if (st != ST_OK) return st;
  st = dh_core(k, P, &Q);
  if (st != ST_OK) return st;
  encode(Q.x, Q.y, out);
  return ST_OK;
}
void fqo_dh_batch(const uint8_t* k, const uint8_t* enc, uint8_t* out, uint8_t* status, size_t n) {
  for (size_t i = 0; i < n; i++) status[i] = (uint8_t)fqo_dh(k + 32 * i, enc + 32 * i, out + 32 * i);
}
/* fq_dh_affine: DH_windowed on x0|x1|y0|y1 (any 128-bit values, reduced mod p) -> 64 bytes */
void fqo_dh_affine_batch(const uint8_t* k, const uint8_t* xy, uint8_t* out, uint8_t* status, size_t n) {
  for (size_t i = 0; i < n; i++) {
    const uint8_t* b = xy + 64 * i;
    aff P = {{fp_red(ld128(b)), fp_red(ld128(b + 16))}, {fp_red(ld128(b + 32)), fp_red(ld128(b + 48))}}, Q;
    memset(out + 64 * i, 0, 64);
    int st = dh_core(k + 32 * i, P, &Q);
    status[i] = (uint8_t)st;
    if (st == ST_OK) { st128(out + 64 * i, Q.x.re); st128(out + 64 * i + 16, Q.x.im); st128(out + 64 * i + 32, Q.y.re); st128(out + 64 * i + 48, Q.y.im); }
  }
}
/* fq_mul_base: encode(R1toAffine(MUL_windowed(k, G, table_windowed(G)))) = [k]G (curve4q.py:582-584) */
void fqo_mul_base_batch(const uint8_t* k, uint8_t* out, size_t n) {
  r1 g = affine_to_r1(GX, GY);
  r2 T[8];
  table_windowed(g, T);
  for (size_t i = 0; i < n; i++) {
    aff a = r1_to_affine(mul_windowed(k + 32 * i, g, T));
    memset(out + 32 * i, 0, 32);
    encode(a.x, a.y, out + 32 * i);
  }
}
/* fq_dh_base: encode(DH_windowed(k, G, table=T392)) = [392 k]G with the neutral check (curve4q.py:743-762) */
void fqo_dh_base_batch(const uint8_t* k, uint8_t* out, uint8_t* status, size_t n) {
  uint8_t genc[32];
  memset(genc, 0, 32);
  encode(GX, GY, genc);
  for (size_t i = 0; i < n; i++) status[i] = (uint8_t)fqo_dh(k + 32 * i, genc, out + 32 * i);
}
/* fq_point_on_curve */
void fqo_on_curve_batch(const uint8_t* xy, uint8_t* ok, size_t n) {
  for (size_t i = 0; i < n; i++) {
    const uint8_t* b = xy + 64 * i;
    f2 x = {fp_red(ld128(b)), fp_red(ld128(b + 16))}, y = {fp_red(ld128(b + 32)), fp_red(ld128(b + 48))};
    ok[i] = (uint8_t)on_curve(x, y);
  }
}
/* fq_decode / fq_encode */
void fqo_decode_batch(const uint8_t* enc, uint8_t* xy, uint8_t* status, size_t n) {
  for (size_t i = 0; i < n; i++) {
    aff P;
    int st = decode(enc + 32 * i, &P);
    status[i] = (uint8_t)st;
    memset(xy + 64 * i, 0, 64);
    if (st == ST_OK) { st128(xy + 64 * i, P.x.re); st128(xy + 64 * i + 16, P.x.im); st128(xy + 64 * i + 32, P.y.re); st128(xy + 64 * i + 48, P.y.im); }
  }
}
void fqo_encode_batch(const uint8_t* xy, uint8_t* enc, size_t n) {
  for (size_t i = 0; i < n; i++) {
    const uint8_t* b = xy + 64 * i;
    f2 x = {fp_red(ld128(b)), fp_red(ld128(b + 16))}, y = {fp_red(ld128(b + 32)), fp_red(ld128(b + 48))};
    memset(enc + 32 * i, 0, 32);
    encode(x, y, enc + 32 * i);
  }
}
/* GF(p^2) ops on 32-byte rows: op 0 mul, 1 sqr, 2 inv, 3 add, 4 sub, 5 neg, 6 conj (any 128-bit halves, reduced mod p) */
int fqo_fp2_batch(int op, const uint8_t* a, const uint8_t* b, uint8_t* out, size_t n) {
  for (size_t i = 0; i < n; i++) {
    f2 x = {fp_red(ld128(a + 32 * i)), fp_red(ld128(a + 32 * i + 16))}, y = F2_ZERO, r;
    if (b) { y.re = fp_red(ld128(b + 32 * i)); y.im = fp_red(ld128(b + 32 * i + 16)); }
    switch (op) {
      case 0: r = f2_mul(x, y); break;
      case 1: r = f2_sqr(x); break;
      case 2: r = f2_inv(x); break;
      case 3: r = f2_add(x, y); break;
      case 4: r = f2_sub(x, y); break;
      case 5: r = f2_neg(x); break;
      case 6: r = f2_conj(x); break;
      default: return -1;
    }
    st128(out + 32 * i, r.re); st128(out + 32 * i + 16, r.im);
  }
  return 0;
}
/* GF(p) ops on 16-byte rows: op 0 mul, 1 sqr, 2 inv, 3 add, 4 sub, 5 neg, 6 invsqrt */
int fqo_fp_batch(int op, const uint8_t* a, const uint8_t* b, uint8_t* out, size_t n) {
  for (size_t i = 0; i < n; i++) {
    u128 x = fp_red(ld128(a + 16 * i)), y = b ? fp_red(ld128(b + 16 * i)) : 0, r;
    switch (op) {
      case 0: r = fp_mul(x, y); break;
      case 1: r = fp_sqr(x); break;
      case 2: r = fp_inv(x); break;
      case 3: r = fp_add(x, y); break;
      case 4: r = fp_sub(x, y); break;
      case 5: r = fp_neg(x); break;
      case 6: r = fp_invsqrt(x); break;
      default: return -1;
    }
    st128(out + 16 * i, r);
  }
  return 0;
}
