// codec.cuh -- Curve4Q point compression (curve4q.py:33-46) and decompression + validation (curve4q.py:49-96).
// decode mirrors the reference step by step so that every failure class, including the reference's accidental
// AttributeError when t == 0 (curve4q.py:76-77), maps to the same rows.  Status codes: include/fourq_b200.h.
#pragma once
#include "point.cuh"

#define FQ_ST_OK 0
#define FQ_ST_RESERVED_BIT 1
#define FQ_ST_NONCANONICAL 2
#define FQ_ST_QUIRK_T0 3
#define FQ_ST_NOT_ON_CURVE 4
#define FQ_ST_NEUTRAL 5

// curve4q.py:33-39 on canonical x
FQ_FN u32 pt_sign(const fp2& x) {
  u32 nz = x.re.v[0] | x.re.v[1] | x.re.v[2] | x.re.v[3];
  return (nz != 0) ? (x.re.v[3] >> 30) : (x.im.v[3] >> 30);
}

// curve4q.py:41-46.  x, y canonical.  out = 8 little-endian words.
FQ_FN void pt_encode(const fp2& x, const fp2& y, u32* out) {
  FQ_UNROLL
  for (int i = 0; i < 4; i++) { out[i] = y.re.v[i]; out[4 + i] = y.im.v[i]; }
  out[7] |= pt_sign(x) << 31;
}

// curve4q.py:49-96.  in = 8 little-endian words of the 32-byte string.  Returns the status; x, y canonical when OK.
// SPEC = false (default everywhere): bit-compatible with the reference, which raises AttributeError when t == 0
// (curve4q.py:76-77 calls a GFp.two that does not exist) -> FQ_ST_QUIRK_T0.  SPEC = true (fq_decode_spec only): what the
// draft specifies there, t = 2 (t0 - t3) (draft-ladd-cfrg-4q.md:865-867), so the encodings of (0, +-1) and (+-i, 0) decode.
template <bool SPEC = false> FQ_FN u32 pt_decode(const u32* in, fp2& x, fp2& y) {
  u32 st = FQ_ST_OK;
  if (in[3] >> 31) st = FQ_ST_RESERVED_BIT;                                   // :52  B[15] & 0x80
  u32 s = in[7] >> 31;                                                        // :55
  y.re = fp_set(in[0], in[1], in[2], in[3] & FQ_P3);                          // :58  (fromLittleEndian clears bit 127)
  y.im = fp_set(in[4], in[5], in[6], in[7] & FQ_P3);                          // :59
  u32 y0p = in[0] & in[1] & in[2] & (in[3] | 0x80000000u);
  u32 y1p = in[4] & in[5] & in[6] & (in[7] | 0x80000000u);
  if (st == FQ_ST_OK && (y0p == 0xffffffffu || y1p == 0xffffffffu)) st = FQ_ST_NONCANONICAL;   // :61

  fp2 y2 = fp2_sqr_c(y);                                                       // :65
  fp2 u = fp2_sub(y2, fp2_one());                                             // :66
  const fp2 dy2 = fp2_mul_c(curve_d(), y2);
  fp2 v = fp2_add(dy2, fp2_one());                                            // :67
  fpb V0 = fp_prep(v.re), V1 = fp_prep(v.im);
  fp t0 = fp_add(fp_mul_prep(u.re, V0), fp_mul_prep(u.im, V1));               // :69
  fp t1 = fp_sub(fp_mul_prep(u.im, V0), fp_mul_prep(u.re, V1));               // :70
  fp t2 = fp_add(fp_sqr(v.re), fp_sqr(v.im));                                 // :71
  fp t3 = fp_add(fp_sqr(t0), fp_sqr(t1));                                     // :72
  t3 = fp_mul(fp_invsqrt_c(t3), t3);                                            // :73
  fp t = fp_dbl(fp_add(t0, t3));                                              // :75
  if (SPEC) { if (fp_is_zero(t)) t = fp_dbl(fp_sub(t0, t3)); }                // draft :865-867
  else if (st == FQ_ST_OK && fp_is_zero(t)) st = FQ_ST_QUIRK_T0;              // :76-77 (the reference raises here)
  fpb T2 = fp_prep(t2);
  fp a = fp_invsqrt_c(fp_mul(fp_mul_prep(fp_sqr(t2), T2), t));                  // :79
  fp at2 = fp_mul_prep(a, T2);
  fp b = fp_mul(at2, t);                                                      // :80
  fp x0 = fp_half(b);                                                         // :82
  fp x1 = fp_mul(at2, t1);                                                    // :83
  bool same = fp_eq_canon(fp_canon(t), fp_canon(fp_mul_prep(fp_sqr(b), T2))); // :84
  u32 msw = same ? 0u : 0xffffffffu;
  x.re = fp_canon(fp_select(msw, x1, x0));                                    // :85
  x.im = fp_canon(fp_select(msw, x0, x1));
  u32 mneg = (pt_sign(x) != s) ? 0xffffffffu : 0u;                            // :88-89
  x = fp2_canon(fp2_select(mneg, fp2_neg(x), x));
  y = fp2_canon(y);
  // :91-94: PointOnCurve (curve4q.py:23-29) of both candidates, x and conj(x).  y^2 and d y^2 are at hand from :65-67 and
  // conj(x)^2 = conj(x^2), so the two tests -x^2 + y^2 == 1 + (d y^2) x^2 cost one squaring and two multiplications.
  const fp2 x2 = fp2_sqr_c(x), x2c = fp2_conj(x2);
  bool on1 = fp2_eq(fp2_sub(y2, x2), fp2_add(fp2_one(), fp2_mul_c(dy2, x2)));
  fp2 xc = fp2_canon(fp2_conj(x));
  bool on2 = fp2_eq(fp2_sub(y2, x2c), fp2_add(fp2_one(), fp2_mul_c(dy2, x2c)));
  x = fp2_select(on1 ? 0xffffffffu : 0u, x, xc);
  if (st == FQ_ST_OK && !(on1 | on2)) st = FQ_ST_NOT_ON_CURVE;
  return st;
}
