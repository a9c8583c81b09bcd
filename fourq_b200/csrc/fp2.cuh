// fp2.cuh -- GF(p^2) = GF(p)[i]/(i^2+1), p = 2^127-1.  Replaces impl/fields.py:134-238 (GFp2.*).
// Karatsuba (3 GF(p) products = 48 IMAD.WIDE) with lazy reduction: the three products stay in 6-limb unreduced
// form, are combined there, and only the two results are folded (draft-ladd-cfrg-4q.md:227-230 suggests the same
// operation counts).  Squaring uses (a0+a1)(a0-a1), 2 a0 a1 = 32 IMAD.WIDE.
#pragma once
#include "fp.cuh"

struct fp2 { fp re, im; };

// prepared right-hand operand of fp2_mul: rotations of re, im and re+im
struct fp2b { fpb re, im, sum; };

FQ_FN fp2 fp2_set(const fp& re, const fp& im) { fp2 r; r.re = re; r.im = im; return r; }
FQ_FN fp2 fp2_zero() { return fp2_set(fp_zero(), fp_zero()); }
FQ_FN fp2 fp2_one() { return fp2_set(fp_one(), fp_zero()); }

FQ_FN fp2 fp2_add(const fp2& a, const fp2& b) { return fp2_set(fp_add(a.re, b.re), fp_add(a.im, b.im)); }     // fields.py:157
FQ_FN fp2 fp2_sub(const fp2& a, const fp2& b) { return fp2_set(fp_sub(a.re, b.re), fp_sub(a.im, b.im)); }     // fields.py:162
FQ_FN fp2 fp2_neg(const fp2& a) { return fp2_set(fp_neg(a.re), fp_neg(a.im)); }                               // fields.py:184
FQ_FN fp2 fp2_conj(const fp2& a) { return fp2_set(a.re, fp_neg(a.im)); }                                      // fields.py:189
FQ_FN fp2 fp2_dbl(const fp2& a) { return fp2_set(fp_dbl(a.re), fp_dbl(a.im)); }
FQ_FN fp2 fp2_canon(const fp2& a) { return fp2_set(fp_canon(a.re), fp_canon(a.im)); }
FQ_FN bool fp2_eq_canon(const fp2& a, const fp2& b) { return fp_eq_canon(a.re, b.re) & fp_eq_canon(a.im, b.im); }
FQ_FN bool fp2_eq(const fp2& a, const fp2& b) { return fp2_eq_canon(fp2_canon(a), fp2_canon(b)); }
FQ_FN fp2 fp2_select(u32 m, const fp2& x, const fp2& y) { return fp2_set(fp_select(m, x.re, y.re), fp_select(m, x.im, y.im)); }  // fields.py:237

#ifdef FQ_FP2_SCHOOLBOOK
// Variant: four GF(p) products in two accumulators, (a0 b0 + (p - a1) b1, a0 b1 + a1 b0): 64 IMAD.WIDE but no
// combination step (no merges of partial products, no multiword subtractions) and no prepared sum.  a must be tight.
FQ_FN fp2b fp2_prep(const fp2& b) {
  fp2b B;
  B.re = fp_prep(b.re); B.im = fp_prep(b.im); B.sum = B.re;
  return B;
}
FQ_FN fp2 fp2_mul_prep(const fp2& a, const fp2b& B) {
  facc A0, A1;
  facc_mul<true>(A0, a.re, B.re); facc_mul<false>(A0, fp_neg(a.im), B.im);
  facc_mul<true>(A1, a.re, B.im); facc_mul<false>(A1, a.im, B.re);
  fp2 r;
  r.re = fp_fold(facc_merge(A0));
  r.im = fp_fold(facc_merge(A1));
  return r;
}
#else
FQ_FN fp2b fp2_prep(const fp2& b) {
  fp2b B;
  B.re = fp_prep(b.re); B.im = fp_prep(b.im); B.sum = fp_prep(fp_add(b.re, b.im));
  return B;
}

// fields.py:167-173.  (a0 b0 - a1 b1, (a0+a1)(b0+b1) - a0 b0 - a1 b1).
// Offsets: t0 starts at 2^36 p >= t1 (< 2^162); t2 starts at 2^38 p >= t0 + t1; their difference is 0 mod p.
FQ_FN fp2 fp2_mul_prep(const fp2& a, const fp2b& B) {
  facc A0, A1, A2;
  facc_init_kp<36>(A0); facc_mul<false>(A0, a.re, B.re);
  facc_mul<true>(A1, a.im, B.im);
  facc_init_kp<38>(A2); facc_mul<false>(A2, fp_add_loose(a.re, a.im), B.sum);
  fpw t0 = facc_merge(A0), t1 = facc_merge(A1), t2 = facc_merge(A2);
  fp2 r;
  r.re = fp_fold(fpw_sub(t0, t1));
  r.im = fp_fold(fpw_sub(fpw_sub(t2, t0), t1));
  return r;
}
#endif
FQ_FN fp2 fp2_mul(const fp2& a, const fp2& b) { return fp2_mul_prep(a, fp2_prep(b)); }

// fields.py:176-181.  ((a0+a1)(a0-a1), 2 a0 a1)
FQ_FN fp2 fp2_sqr(const fp2& a) {
  fp2 r;
  r.re = fp_mul_prep(fp_add_loose(a.re, a.im), fp_prep(fp_sub(a.re, a.im)));
  r.im = fp_mul_prep(fp_add_loose(a.re, a.re), fp_prep(a.im));
  return r;
}

// fields.py:194-199.  conj(a) / (a0^2 + a1^2); inv(0) = 0
FQ_FN fp2 fp2_inv(const fp2& a) {
  fp n = fp_inv(fp_add(fp_sqr(a.re), fp_sqr(a.im)));
  fpb N = fp_prep(n);
  return fp2_set(fp_mul_prep(a.re, N), fp_mul_prep(fp_neg(a.im), N));
}

// ---------------------------------------------------------------- out-of-line copies
// The once-per-row setup code (decode, cofactor clearing, endomorphisms, table build) is long and straight-line; inlining
// every multiplication there makes the kernel several hundred KiB of code and the instruction cache thrash.  These
// copies are real calls (arguments and result in registers); the hot loops keep using the inlined versions.
#ifdef FQ_HOSTSIM
#define FQ_CALL static inline
#else
#define FQ_CALL static __device__ __noinline__
#endif
FQ_CALL fp2 fp2_mul_c(fp2 a, fp2 b) { return fp2_mul(a, b); }
FQ_CALL fp2 fp2_sqr_c(fp2 a) { return fp2_sqr(a); }
FQ_CALL fp fp_invsqrt_c(fp a) { return fp_invsqrt(a); }
