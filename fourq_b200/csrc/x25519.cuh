// x25519.cuh -- batched X25519 (RFC 7748), the comparison curve of the reference (impl/curve25519.py:17-91 with
// GFp25519 of impl/fields.py:240-362).  GF(2^255-19) elements are 8 x 32-bit limbs kept "loose" (any value < 2^256,
// congruent mod p); 2^256 = 38 (mod p).  Multiplication: 64 IMAD.WIDE in even/odd 64-bit lattices, then 8 more for
// the 38 * hi fold.  One thread = one ladder (255 steps, cswap by masks, curve25519.py:51-76).
#pragma once
#include "arith.cuh"

struct f25 { u32 v[8]; };

FQ_FN f25 f25_small(u32 x) { f25 r; r.v[0] = x; FQ_UNROLL for (int i = 1; i < 8; i++) r.v[i] = 0; return r; }

// r + 38*c for a carry/borrow correction c in {0,1}; cannot overflow twice
FQ_FN f25 f25_fix_carry(f25 r, u32 c) {
  u32 k = 38u * c;
  r.v[0] = add_cc(r.v[0], k);
  FQ_UNROLL
  for (int i = 1; i < 8; i++) r.v[i] = addc_cc(r.v[i], 0);
  u32 c2 = addc(0, 0);
  r.v[0] += 38u * c2;               // after a wrap the value is < 38, no further carry
  return r;
}
// fields.py:267-270
FQ_FN f25 f25_add(const f25& a, const f25& b) {
  f25 r;
  r.v[0] = add_cc(a.v[0], b.v[0]);
  FQ_UNROLL
  for (int i = 1; i < 8; i++) r.v[i] = addc_cc(a.v[i], b.v[i]);
  return f25_fix_carry(r, addc(0, 0));
}
// fields.py:273-276.  a - b wraps to a - b + 2^256 = a - b + 38 (mod p): take the 38 back.
FQ_FN f25 f25_sub(const f25& a, const f25& b) {
  f25 r;
  r.v[0] = sub_cc(a.v[0], b.v[0]);
  FQ_UNROLL
  for (int i = 1; i < 8; i++) r.v[i] = subc_cc(a.v[i], b.v[i]);
  u32 bw = 0u - subc(0, 0);         // 1 if borrow
  u32 k = 38u * bw;
  r.v[0] = sub_cc(r.v[0], k);
  FQ_UNROLL
  for (int i = 1; i < 8; i++) r.v[i] = subc_cc(r.v[i], 0);
  u32 bw2 = 0u - subc(0, 0);
  r.v[0] -= 38u * bw2;              // after a second wrap the value is >= 2^256 - 38, no further borrow
  return r;
}
// fields.py:259-264: m all ones swaps
FQ_FN void f25_cswap(u32 m, f25& x, f25& y) {
  FQ_UNROLL
  for (int i = 0; i < 8; i++) { u32 t = m & (x.v[i] ^ y.v[i]); x.v[i] ^= t; y.v[i] ^= t; }
}

// 9-limb value t (t[8] small) -> loose 8-limb: t[0..7] + 38 * t[8]
FQ_FN f25 f25_fold9(const u32* t) {
  f25 r;
  u32 lo, hi;
  mul_wide(lo, hi, t[8], 38u);
  r.v[0] = add_cc(t[0], lo); r.v[1] = addc_cc(t[1], hi);
  FQ_UNROLL
  for (int i = 2; i < 8; i++) r.v[i] = addc_cc(t[i], 0);
  return f25_fix_carry(r, addc(0, 0));
}

// 16-limb product m -> loose 8-limb: m[0..7] + 38 * m[8..15] (2^256 = 38 mod p), 8 IMAD.WIDE
FQ_FN f25 f25_reduce16(const u32* m) {
  u32 t[9], P[9];
  FQ_UNROLL
  for (int k = 0; k < 8; k++) t[k] = m[k];
  t[0] = mad_lo_cc(m[8], 38u, t[0]); t[1] = madc_hi_cc(m[8], 38u, t[1]);
  t[2] = madc_lo_cc(m[10], 38u, t[2]); t[3] = madc_hi_cc(m[10], 38u, t[3]);
  t[4] = madc_lo_cc(m[12], 38u, t[4]); t[5] = madc_hi_cc(m[12], 38u, t[5]);
  t[6] = madc_lo_cc(m[14], 38u, t[6]); t[7] = madc_hi_cc(m[14], 38u, t[7]);
  t[8] = addc(0, 0);
  mul_wide(P[1], P[2], m[9], 38u); mul_wide(P[3], P[4], m[11], 38u);
  mul_wide(P[5], P[6], m[13], 38u); mul_wide(P[7], P[8], m[15], 38u);
  t[1] = add_cc(t[1], P[1]);
  FQ_UNROLL
  for (int k = 2; k < 8; k++) t[k] = addc_cc(t[k], P[k]);
  t[8] = addc(t[8], P[8]);
  return f25_fold9(t);
}

// fields.py:279-282
FQ_FN f25 f25_mul(const f25& a, const f25& b) {
  // E[k] holds limb k of the even lattice, O[k] limb k+1 of the odd lattice
  u32 E[16], O[16];
  FQ_UNROLL
  for (int i = 0; i < 16; i++) { E[i] = 0; O[i] = 0; }
  FQ_UNROLL
  for (int i = 0; i < 8; i++) {
    // products a_j * b_i at limb position i + j
    FQ_UNROLL
    for (int par = 0; par < 2; par++) {
      // par = parity of j handled by this chain
      const bool even_pos = ((i + par) & 1) == 0;
      u32* A = even_pos ? E : O;
      const int base = even_pos ? (i + par) : (i + par - 1);
      FQ_UNROLL
      for (int jj = 0; jj < 4; jj++) {
        const int j = 2 * jj + par, k = base + 2 * jj;
        if (jj == 0) { A[k] = mad_lo_cc(a.v[j], b.v[i], A[k]); A[k + 1] = madc_hi_cc(a.v[j], b.v[i], A[k + 1]); }
        else { A[k] = madc_lo_cc(a.v[j], b.v[i], A[k]); A[k + 1] = madc_hi_cc(a.v[j], b.v[i], A[k + 1]); }
      }
      if (base + 8 < 16) A[base + 8] = addc(A[base + 8], 0);
    }
  }
  // merge: m = E + (O << 32)
  u32 m[16];
  m[0] = E[0];
  m[1] = add_cc(E[1], O[0]);
  FQ_UNROLL
  for (int k = 2; k < 15; k++) m[k] = addc_cc(E[k], O[k - 1]);
  m[15] = addc(E[15], O[14]);
  return f25_reduce16(m);
}

// fields.py:285-288 as a true squaring: the 28 cross products a_i a_j (i < j) in the two lattices, doubled by a one-bit shift
// of the merged 16-limb sum, plus the 8 squares (which land on disjoint limb pairs): 36 IMAD.WIDE instead of 64.
FQ_FN f25 f25_sqr(const f25& a) {
  u32 E[16], O[16];
  FQ_UNROLL
  for (int i = 0; i < 16; i++) { E[i] = 0; O[i] = 0; }
  FQ_UNROLL
  for (int i = 0; i < 7; i++) {
    FQ_UNROLL
    for (int par = 0; par < 2; par++) {
      // partners a_j with j > i and j = par (mod 2); their products sit at limb position i + j: one carry chain
      const bool even_pos = ((i + par) & 1) == 0;
      u32* A = even_pos ? E : O;
      const int j0 = (((i + 1) & 1) == par) ? i + 1 : i + 2;          // smallest j > i of this parity
      const int first_jj = (j0 - par) / 2;
      if (first_jj > 3) continue;
      FQ_UNROLL
      for (int jj = 0; jj < 4; jj++) {
        const int j = 2 * jj + par;
        if (jj < first_jj) continue;
        const int k = even_pos ? (i + j) : (i + j - 1);
        if (jj == first_jj) { A[k] = mad_lo_cc(a.v[j], a.v[i], A[k]); A[k + 1] = madc_hi_cc(a.v[j], a.v[i], A[k + 1]); }
        else { A[k] = madc_lo_cc(a.v[j], a.v[i], A[k]); A[k + 1] = madc_hi_cc(a.v[j], a.v[i], A[k + 1]); }
      }
      const int kend = (even_pos ? (i + 6 + par) : (i + 6 + par - 1)) + 2;      // the pair after the last one (j = 6 + par)
      if (kend < 16) A[kend] = addc(A[kend], 0);
    }
  }
  // cross sum c = E + (O << 32), then 2 c
  u32 c[16], m[16];
  c[0] = E[0];
  c[1] = add_cc(E[1], O[0]);
  FQ_UNROLL
  for (int k = 2; k < 15; k++) c[k] = addc_cc(E[k], O[k - 1]);
  c[15] = addc(E[15], O[14]);
  m[0] = c[0] << 1;
  FQ_UNROLL
  for (int k = 1; k < 16; k++) m[k] = shl_pair(c[k - 1], c[k], 1);
  // + squares
  u32 s[16];
  FQ_UNROLL
  for (int i = 0; i < 8; i++) mul_wide(s[2 * i], s[2 * i + 1], a.v[i], a.v[i]);
  m[0] = add_cc(m[0], s[0]);
  FQ_UNROLL
  for (int k = 1; k < 15; k++) m[k] = addc_cc(m[k], s[k]);
  m[15] = addc(m[15], s[15]);
  return f25_reduce16(m);
}


// a * 121665 (curve25519.py:76, a24)
FQ_FN f25 f25_mul_a24(const f25& a) {
  u32 t[9], P[9];
  const u32 c = 121665u;
  mul_wide(t[0], t[1], a.v[0], c); mul_wide(t[2], t[3], a.v[2], c); mul_wide(t[4], t[5], a.v[4], c); mul_wide(t[6], t[7], a.v[6], c);
  mul_wide(P[1], P[2], a.v[1], c); mul_wide(P[3], P[4], a.v[3], c); mul_wide(P[5], P[6], a.v[5], c); mul_wide(P[7], P[8], a.v[7], c);
  t[1] = add_cc(t[1], P[1]);
  FQ_UNROLL
  for (int k = 2; k < 8; k++) t[k] = addc_cc(t[k], P[k]);
  t[8] = addc(0, P[8]);
  return f25_fold9(t);
}

FQ_FN f25 f25_nsqr(f25 x, int n) {
  FQ_NOUNROLL
  for (int i = 0; i < n; i++) x = f25_sqr(x);
  return x;
}
// fields.py:293-362: z^(p-2) = z^(2^255 - 21), same chain shape (254 S + 11 M)
FQ_FN f25 f25_inv(const f25& z) {
  f25 z2 = f25_sqr(z);
  f25 z9 = f25_mul(f25_nsqr(z2, 2), z);
  f25 z11 = f25_mul(z9, z2);
  f25 z2_5_0 = f25_mul(f25_sqr(z11), z9);
  f25 z2_10_0 = f25_mul(f25_nsqr(z2_5_0, 5), z2_5_0);
  f25 z2_20_0 = f25_mul(f25_nsqr(z2_10_0, 10), z2_10_0);
  f25 z2_40_0 = f25_mul(f25_nsqr(z2_20_0, 20), z2_20_0);
  f25 z2_50_0 = f25_mul(f25_nsqr(z2_40_0, 10), z2_10_0);
  f25 z2_100_0 = f25_mul(f25_nsqr(z2_50_0, 50), z2_50_0);
  f25 z2_200_0 = f25_mul(f25_nsqr(z2_100_0, 100), z2_100_0);
  f25 z2_250_0 = f25_mul(f25_nsqr(z2_200_0, 50), z2_50_0);
  return f25_mul(f25_nsqr(z2_250_0, 5), z11);
}

// loose -> canonical representative in [0, p)
FQ_FN f25 f25_canon(f25 r) {
  // fold bit 255: r = (r mod 2^255) + 19 * (r >> 255)  < 2^255 + 19
  u32 top = r.v[7] >> 31;
  r.v[7] &= 0x7fffffffu;
  r.v[0] = add_cc(r.v[0], 19u * top);
  FQ_UNROLL
  for (int i = 1; i < 7; i++) r.v[i] = addc_cc(r.v[i], 0);
  r.v[7] = addc(r.v[7], 0);
  // r >= p  <=>  r + 19 >= 2^255
  f25 s;
  s.v[0] = add_cc(r.v[0], 19u);
  FQ_UNROLL
  for (int i = 1; i < 7; i++) s.v[i] = addc_cc(r.v[i], 0);
  s.v[7] = addc(r.v[7], 0);
  u32 ge = 0u - (s.v[7] >> 31);     // all ones if r >= p; then r - p = (r + 19) mod 2^255
  s.v[7] &= 0x7fffffffu;
  FQ_UNROLL
  for (int i = 0; i < 8; i++) r.v[i] = (s.v[i] & ge) | (r.v[i] & ~ge);
  return r;
}

// curve25519.py:17-80 up to (and not including) the final inversion: clamp, mask, 255 ladder steps, final swap.
// The result is x2 / z2.
FQ_FN void x25519_ladder(const u32* kw, const u32* uw, f25& x2, f25& z2) {
  u32 k[8];
  FQ_UNROLL
  for (int i = 0; i < 8; i++) k[i] = kw[i];
  k[0] &= 0xfffffff8u; k[7] = (k[7] & 0x7fffffffu) | 0x40000000u;           // curve25519.py:20-25
  f25 x1;
  FQ_UNROLL
  for (int i = 0; i < 8; i++) x1.v[i] = uw[i];
  x1.v[7] &= 0x7fffffffu;                                                    // curve25519.py:27-33
  f25 x3 = x1, z3 = f25_small(1);
  x2 = f25_small(1); z2 = f25_small(0);
  u32 swap = 0;
  FQ_NOUNROLL
  for (int t = 254; t >= 0; t--) {                                           // curve25519.py:51-76
    u32 kt = (k[t >> 5] >> (t & 31)) & 1;
    u32 m = 0u - (swap ^ kt);
    f25_cswap(m, x2, x3); f25_cswap(m, z2, z3);
    swap = kt;
    f25 A = f25_add(x2, z2), B = f25_sub(x2, z2);
    f25 AA = f25_sqr(A), BB = f25_sqr(B);
    f25 E = f25_sub(AA, BB);
    f25 C = f25_add(x3, z3), D = f25_sub(x3, z3);
    f25 DA = f25_mul(D, A), CB = f25_mul(C, B);
    x3 = f25_sqr(f25_add(DA, CB));
    z3 = f25_mul(x1, f25_sqr(f25_sub(DA, CB)));
    x2 = f25_mul(AA, BB);
    z2 = f25_mul(E, f25_add(AA, f25_mul_a24(E)));
  }
  u32 m = 0u - swap;
  f25_cswap(m, x2, x3); f25_cswap(m, z2, z3);                                 // curve25519.py:78-79
}

// curve25519.py:88-91: k, u, out = 8 little-endian words each
FQ_FN void row_x25519(const u32* kw, const u32* uw, u32* out) {
  f25 x2, z2;
  x25519_ladder(kw, uw, x2, z2);
  f25 r = f25_canon(f25_mul(x2, f25_inv(z2)));                                // curve25519.py:80, 35-39
  FQ_UNROLL
  for (int i = 0; i < 8; i++) out[i] = r.v[i];
}

// ---------------------------------------------------------------- shared inversion of the final x2 / z2 (batchinv.cuh)
// scratch: x2 then z2, each as two quads per row, component-major ([4][npad] uint4)
FQ_FN void st_f25(uint4* p, size_t npad, const f25& a) {
  p[0] = make_uint4(a.v[0], a.v[1], a.v[2], a.v[3]); p[npad] = make_uint4(a.v[4], a.v[5], a.v[6], a.v[7]);
}
FQ_FN f25 ld_f25(const uint4* p, size_t npad) {
  uint4 a = p[0], b = p[npad];
  f25 r; r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w; r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
  return r;
}

// x2 * z2^(p-2) (curve25519.py:78-80) and the canonical little-endian output (:35-39) for FQ_BATCHINV_ROWS rows per thread
struct F25Ops {
  typedef f25 elem;
  static FQ_MFN f25 one() { return f25_small(1); }
  static FQ_MFN f25 mul(const f25& a, const f25& b) { return f25_mul(a, b); }
  static FQ_MFN f25 inv(const f25& a) { return f25_inv(a); }
};
// a loose value that is 0 mod p (0, p or 2p) is replaced by one; zero = all ones if it was
FQ_FN f25 f25_one_if_zero(f25 v, u32& zero) {
  f25 c = f25_canon(v);
  u32 nz = 0;
  FQ_UNROLL
  for (int i = 0; i < 8; i++) nz |= c.v[i];
  zero = nz == 0 ? 0xffffffffu : 0u;
  FQ_UNROLL
  for (int i = 0; i < 8; i++) v.v[i] = (v.v[i] & ~zero) | ((i == 0 ? 1u : 0u) & zero);
  return v;
}
struct X25519FinIO {
  const uint4* scratch; size_t npad; uint4* out; size_t n, t, stride;
  FQ_MFN f25 z(int j, u32& zero) const {
    const size_t row = t + (size_t)j * stride;
    f25 v = f25_small(1);
    if (row < n) v = ld_f25(scratch + 2 * npad + row, npad);
    return f25_one_if_zero(v, zero);
  }
  FQ_MFN void park(int j, const f25& acc) const {
    const size_t row = t + (size_t)j * stride;
    if (row < n) st_f25(out + 2 * row, 1, acc);
  }
  FQ_MFN f25 parked(int j) const {
    const size_t row = t + (size_t)j * stride;
    return row < n ? ld_f25(out + 2 * row, 1) : f25_small(1);
  }
  FQ_MFN void emit(int j, const f25& zi, u32 zero) const {
    const size_t row = t + (size_t)j * stride;
    if (row >= n) return;
    f25 r = f25_canon(f25_mul(ld_f25(scratch + row, npad), zi));
    out[2 * row] = make_uint4(r.v[0] & ~zero, r.v[1] & ~zero, r.v[2] & ~zero, r.v[3] & ~zero);
    out[2 * row + 1] = make_uint4(r.v[4] & ~zero, r.v[5] & ~zero, r.v[6] & ~zero, r.v[7] & ~zero);
  }
};

// ---------------------------------------------------------------- GFp25519.add/sub/mul/sqr/inv on 32-byte rows
// (fields.py:267-362; the GFp25519 column of compare.py:14-49 compare_fields).  Any 256-bit input (the reference reduces ints
// mod p), canonical little-endian output.
enum { FQ_F25OP_MUL = 0, FQ_F25OP_SQR = 1, FQ_F25OP_INV = 2, FQ_F25OP_ADD = 3, FQ_F25OP_SUB = 4 };
FQ_FN f25 f25_from_words(const u32* w) { f25 r; FQ_UNROLL for (int i = 0; i < 8; i++) r.v[i] = w[i]; return r; }
template <int OP> FQ_FN void row_f25_op(const u32* a, const u32* b, u32* out) {
  const f25 x = f25_from_words(a);
  f25 r;
  if (OP == FQ_F25OP_MUL) r = f25_mul(x, f25_from_words(b));
  else if (OP == FQ_F25OP_ADD) r = f25_add(x, f25_from_words(b));
  else if (OP == FQ_F25OP_SUB) r = f25_sub(x, f25_from_words(b));
  else if (OP == FQ_F25OP_SQR) r = f25_sqr(x);
  else r = f25_inv(x);
  r = f25_canon(r);
  FQ_UNROLL
  for (int i = 0; i < 8; i++) out[i] = r.v[i];
}
// GFp25519.inv for FQ_BATCHINV_ROWS rows per thread with one z^(p-2) chain; inv(0) = 0 as the reference's chain gives
struct F25InvIO {
  const uint4* a; uint4* out; size_t n, t, stride;
  FQ_MFN f25 z(int j, u32& zero) const {
    const size_t row = t + (size_t)j * stride;
    f25 v = f25_small(1);
    if (row < n) v = ld_f25(a + 2 * row, 1);
    return f25_one_if_zero(v, zero);
  }
  FQ_MFN void park(int j, const f25& acc) const {
    const size_t row = t + (size_t)j * stride;
    if (row < n) st_f25(out + 2 * row, 1, acc);
  }
  FQ_MFN f25 parked(int j) const {
    const size_t row = t + (size_t)j * stride;
    return row < n ? ld_f25(out + 2 * row, 1) : f25_small(1);
  }
  FQ_MFN void emit(int j, const f25& zi, u32 zero) const {
    const size_t row = t + (size_t)j * stride;
    if (row >= n) return;
    f25 r = f25_canon(zi);
    out[2 * row] = make_uint4(r.v[0] & ~zero, r.v[1] & ~zero, r.v[2] & ~zero, r.v[3] & ~zero);
    out[2 * row + 1] = make_uint4(r.v[4] & ~zero, r.v[5] & ~zero, r.v[6] & ~zero, r.v[7] & ~zero);
  }
};
