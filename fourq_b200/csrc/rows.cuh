// rows.cuh -- what each batched entry point computes for ONE row, on register-resident words.
// Shared by the CUDA kernels (kernels.cu) and by the CPU instruction-level simulation used in tests/hostsim.
// Byte conventions (include/fourq_b200.h): GF(p^2) element = LE128(re)|LE128(im); affine point = x|y (64 B);
// encoded point and scalar = 32 B.  Words are the little-endian u32 view of those bytes.
#pragma once
#include "endo.cuh"

// any two 128-bit values -> tight GF(p^2) element (the reference's ops accept unreduced ints: fields.py:157-181)
FQ_FN fp2 row_load_fp2(const u32* w) {
  return fp2_set(fp_from_u128(fp_set(w[0], w[1], w[2], w[3])), fp_from_u128(fp_set(w[4], w[5], w[6], w[7])));
}
FQ_FN void row_store_fp2(u32* w, const fp2& a) {
  fp2 c = fp2_canon(a);
  FQ_UNROLL
  for (int i = 0; i < 4; i++) { w[i] = c.re.v[i]; w[4 + i] = c.im.v[i]; }
}
FQ_FN scal row_load_scalar(const u32* w) {
  scal k;
  FQ_UNROLL
  for (int i = 0; i < 8; i++) k.v[i] = w[i];
  return k;
}

enum { FQ_OP_MUL = 0, FQ_OP_SQR = 1, FQ_OP_INV = 2, FQ_OP_ADD = 3, FQ_OP_SUB = 4, FQ_OP_NEG = 5, FQ_OP_CONJ = 6 };

template <int OP> FQ_FN void row_fp2_op(const u32* a, const u32* b, u32* out) {
  fp2 x = row_load_fp2(a), r;
  if (OP == FQ_OP_MUL) r = fp2_mul(x, row_load_fp2(b));
  else if (OP == FQ_OP_ADD) r = fp2_add(x, row_load_fp2(b));
  else if (OP == FQ_OP_SUB) r = fp2_sub(x, row_load_fp2(b));
  else if (OP == FQ_OP_SQR) r = fp2_sqr(x);
  else if (OP == FQ_OP_INV) r = fp2_inv(x);
  else if (OP == FQ_OP_NEG) r = fp2_neg(x);
  else r = fp2_conj(x);
  row_store_fp2(out, r);
}

// GF(p) ops on 16-byte rows (fields.py:29-122 GFp.add/sub/mul/sqr/neg/inv/invsqrt): any 128-bit input (the reference reduces
// ints mod p), canonical output
enum { FQ_FPOP_MUL = 0, FQ_FPOP_SQR = 1, FQ_FPOP_INV = 2, FQ_FPOP_ADD = 3, FQ_FPOP_SUB = 4, FQ_FPOP_NEG = 5, FQ_FPOP_INVSQRT = 6 };
template <int OP> FQ_FN void row_fp_op(const u32* a, const u32* b, u32* out) {
  fp x = fp_from_u128(fp_set(a[0], a[1], a[2], a[3])), r;
  if (OP == FQ_FPOP_MUL) r = fp_mul(x, fp_from_u128(fp_set(b[0], b[1], b[2], b[3])));
  else if (OP == FQ_FPOP_ADD) r = fp_add(x, fp_from_u128(fp_set(b[0], b[1], b[2], b[3])));
  else if (OP == FQ_FPOP_SUB) r = fp_sub(x, fp_from_u128(fp_set(b[0], b[1], b[2], b[3])));
  else if (OP == FQ_FPOP_SQR) r = fp_sqr(x);
  else if (OP == FQ_FPOP_INV) r = fp_inv(x);
  else if (OP == FQ_FPOP_NEG) r = fp_neg(x);
  else r = fp_invsqrt(x);
  r = fp_canon(r);
  FQ_UNROLL
  for (int i = 0; i < 4; i++) out[i] = r.v[i];
}

// fields.py:201-230 GFp2.invsqrt, with the reference's control flow: the raw test a[1] == 0 (:204) picks the GF(p) branch, every
// other row the norm branch; the reference's two comparisons `== -1` (:217, :223) can never be true (GFp.mul returns values in
// [0, p)), so neither the 'not square' exception nor the second delta is ever taken.  8 words in, 8 canonical words out.
FQ_FN void row_fp2_invsqrt(const u32* a, u32* out) {
  const bool a1_raw_zero = (a[4] | a[5] | a[6] | a[7]) == 0;
  const fp a0 = fp_from_u128(fp_set(a[0], a[1], a[2], a[3])), a1 = fp_from_u128(fp_set(a[4], a[5], a[6], a[7]));
  fp x0, x1;
  if (a1_raw_zero) {
    fp t = fp_canon(fp_invsqrt_c(a0));                                          // :205
    const bool is_one = fp_eq_canon(fp_canon(fp_mul(a0, fp_sqr(t))), fp_one());   // :206
    x0 = is_one ? t : fp_zero(); x1 = is_one ? fp_zero() : t;                    // :207-209
  } else {
    fp n = fp_add(fp_sqr(a0), fp_sqr(a1));                                      // :214
    fp sv = fp_invsqrt_c(n);                                                    // :215
    fpb S = fp_prep(sv);
    fp c = fp_mul_prep(n, S);                                                   // :216
    fp delta = fp_half(fp_add(a0, c));                                          // :220
    fp g = fp_invsqrt_c(delta);                                                 // :221
    fp h = fp_mul(delta, g);                                                    // :222
    x0 = fp_mul_prep(h, S);                                                     // :228
    x1 = fp_neg(fp_half(fp_mul(fp_mul_prep(a1, S), g)));                        // :229
  }
  x0 = fp_canon(x0); x1 = fp_canon(x1);
  FQ_UNROLL
  for (int i = 0; i < 4; i++) { out[i] = x0.v[i]; out[4 + i] = x1.v[i]; }
}

// fields.py:59-64 GFp.select / :236-238 GFp2.select on raw words: out = y ^ ((mask * c) & (x ^ y)) with mask = 2^512 - 1, i.e.
// mask * c = -c modulo 2^128.  W words per row half (4), HALVES halves per row that share the condition byte c.
template <int HALVES> FQ_FN void row_select(u32 c, const u32* x, const u32* y, u32* out) {
  const u32 m0 = 0u - c, mh = c ? 0xffffffffu : 0u;           // limbs of -c mod 2^128: the borrow fills the upper limbs
  FQ_UNROLL
  for (int h = 0; h < HALVES; h++) {
    FQ_UNROLL
    for (int i = 0; i < 4; i++) { const u32 m = i == 0 ? m0 : mh; out[4 * h + i] = y[4 * h + i] ^ (m & (x[4 * h + i] ^ y[4 * h + i])); }
  }
}

// decode: 8 words -> 16 words (x0|x1|y0|y1), zero-filled on failure
template <bool SPEC = false> FQ_FN u32 row_decode(const u32* enc, u32* xy) {
  fp2 x, y;
  u32 st = pt_decode<SPEC>(enc, x, y);
  if (st != FQ_ST_OK) { x = fp2_zero(); y = fp2_zero(); }
  row_store_fp2(xy, x); row_store_fp2(xy + 8, y);
  return st;
}
// encode: 16 words -> 8 words.  Coordinates are reduced mod p first.
FQ_FN void row_encode(const u32* xy, u32* enc) {
  fp2 x = fp2_canon(row_load_fp2(xy)), y = fp2_canon(row_load_fp2(xy + 8));
  pt_encode(x, y, enc);
}

FQ_FN void row_zero(u32* w, int n) { for (int i = 0; i < n; i++) w[i] = 0; }

// fq_dh / fq_dh_endo (AFFINE = false): decode(enc) -> DH_windowed | DH_endo -> encode, 8 words in, 8 out.
// fq_dh_affine / fq_dh_endo_affine (AFFINE = true): DH_* on an affine point, 16 words in, 16 out.
// Three phases (see dh.cuh); a row that fails validation still runs them (on a harmless value) so that every thread of
// a CTA reaches the barriers the kernel puts between phases; its status is kept and its output zero-filled.
template <bool ENDO, bool AFFINE> FQ_FN u32 row_dh_setup(const u32* k, const u32* pt, const TabView& T, DhState& D) {
  fp2 x, y;
  u32 st;
  if (AFFINE) {
    x = fp2_canon(row_load_fp2(pt)); y = fp2_canon(row_load_fp2(pt + 8));
    st = pt_on_curve(x, y) ? FQ_ST_OK : FQ_ST_NOT_ON_CURVE;                         // curve4q.py:447
  } else {
    st = pt_decode(pt, x, y);
  }
  if (ENDO) dh_setup_endo(row_load_scalar(k), x, y, T, D); else dh_setup_windowed(row_load_scalar(k), x, y, T, D);
  return st;
}
template <bool ENDO, bool STRICT = FQ_STRICT_DEFAULT> FQ_FN ptR1 row_dh_loop(const TabView& T, DhState& D) { return ENDO ? dh_loop_endo<STRICT>(T, D) : dh_loop_windowed<STRICT>(T, D); }

// fq_mul_base / fq_dh_base and the endo variants: MUL_windowed | MUL_endo on the base point's table (tab: the 64 quads of the
// table, [entry][quad], in shared memory on the device); the result stays projective (R1) for k_dh_finish
template <bool ENDO, bool STRICT = FQ_STRICT_DEFAULT> FQ_FN ptR1 row_fixed_base_r1(const u32* k, uint4* tab) {
  TabView T; T.base = tab; T.stride = 1;
  if (ENDO) { SelectBroadcast<STRICT> sel; sel.T = T; return mul_endo(row_load_scalar(k), sel); }
  SelectBroadcast<STRICT> sel; sel.T = T;
  return mul_windowed(row_load_scalar(k), sel);
}

// Fixed-base tables: out[0..255] = table_windowed(G) (curve4q.py:582), out[256..511] = table_windowed([392]G)
// (curve4q.py:758-759), out[512..767] = table_endo(G), out[768..1023] = table_endo([392]G) (curve4q.py:760).  scratch: 56 uint4.
FQ_FN void row_build_base_tables(u32* out, uint4* scratch) {
  TabView T; T.base = scratch; T.stride = 1;
  for (int which = 0; which < 4; which++) {
    ptR1 B = ((which & 1) == 0) ? pt_from_affine(curve_gx(), curve_gy()) : pt_clear_cofactor(curve_gx(), curve_gy());
    ptR2 T7 = (which < 2) ? tab_build(T, B) : endo_tab_build(T, B);
    for (int e = 0; e < 7; e++) { ptR2 P = tab_load(T, e); r2_to_words(P, out + which * 256 + e * 32); }
    r2_to_words(T7, out + which * 256 + 7 * 32);
  }
}
