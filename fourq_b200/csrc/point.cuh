// point.cuh -- Curve4Q point representations and the complete twisted-Edwards group law, inlined per thread.
// Follows impl/curve4q.py:100-175 (AffineToR1, R1toR2, R1toR3, R2toR4, DBL, ADD_core, ADD) and the cofactor
// clearing of DH_core (curve4q.py:450-455).
//   R1 = (X, Y, Z, Ta, Tb), T = Ta*Tb      R2 = (X+Y, Y-X, 2Z, 2dT)      R3 = (X+Y, Y-X, Z, T)      R4 = (X, Y, Z)
#pragma once
#include "fp2.cuh"

struct ptR1 { fp2 X, Y, Z, Ta, Tb; };
struct ptR2 { fp2 N, D, E, F; };
struct ptR3p { fp2b N, D, E, F; };     // R3 with every coordinate prepared as a right-hand multiplication operand

// curve4q.py:9, and 2d (curve4q.py:115 recomputes it per call)
FQ_FN fp2 curve_d() { return fp2_set(fp_set(0x00000142u, 0x00000000u, 0x000000e4u, 0x00000000u), fp_set(0xf1fc0c8du, 0xb3821488u, 0x6657e0fcu, 0x5e472f84u)); }
FQ_FN fp2 curve_2d() { return fp2_set(fp_set(0x00000284u, 0x00000000u, 0x000001c8u, 0x00000000u), fp_set(0xe3f8191bu, 0x67042911u, 0xccafc1f9u, 0x3c8e5f08u)); }
// curve4q.py:19-20
FQ_FN fp2 curve_gx() { return fp2_set(fp_set(0x7b3833aau, 0x286592adu, 0x7c2fb305u, 0x1a347223u), fp_set(0x60ac77f6u, 0x96869fb3u, 0x2878aa9cu, 0x1e1f553fu)); }
FQ_FN fp2 curve_gy() { return fp2_set(fp_set(0x2bcbb287u, 0xb924a246u, 0xa120785au, 0x0e3fee9bu), fp_set(0x844c8b5cu, 0x49a7c344u, 0x630e0242u, 0x6e1c4af8u)); }

// curve4q.py:23-29:  y^2 - x^2 == 1 + d x^2 y^2
FQ_CALL bool pt_on_curve(fp2 x, fp2 y) {
  fp2 x2 = fp2_sqr_c(x), y2 = fp2_sqr_c(y);
  fp2 lhs = fp2_sub(y2, x2);
  fp2 rhs = fp2_add(fp2_one(), fp2_mul_c(fp2_mul_c(curve_d(), x2), y2));
  return fp2_eq(lhs, rhs);
}

// curve4q.py:100-101
FQ_FN ptR1 pt_from_affine(const fp2& x, const fp2& y) { ptR1 P; P.X = x; P.Y = y; P.Z = fp2_one(); P.Ta = x; P.Tb = y; return P; }

// curve4q.py:109-116
FQ_FN ptR2 pt_r1_to_r2(const ptR1& P) {
  ptR2 R;
  R.N = fp2_add(P.X, P.Y); R.D = fp2_sub(P.Y, P.X); R.E = fp2_dbl(P.Z);
  R.F = fp2_mul(fp2_mul(P.Ta, P.Tb), curve_2d());
  return R;
}
// curve4q.py:119-126, coordinates prepared for repeated use as right-hand operands
FQ_FN ptR3p pt_r1_to_r3p(const ptR1& P) {
  ptR3p R;
  R.N = fp2_prep(fp2_add(P.X, P.Y)); R.D = fp2_prep(fp2_sub(P.Y, P.X)); R.E = fp2_prep(P.Z);
  R.F = fp2_prep(fp2_mul(P.Ta, P.Tb));
  return R;
}
// curve4q.py:129-135 (as R1 with Ta, Tb unset: only DBL may follow)
FQ_FN ptR1 pt_r2_to_r4(const ptR2& P) {
  ptR1 Q; Q.X = fp2_sub(P.N, P.D); Q.Y = fp2_add(P.D, P.N); Q.Z = P.E; Q.Ta = fp2_zero(); Q.Tb = fp2_zero();
  return Q;
}
// curve4q.py:193-195, under a mask (all ones = negate): swap N,D and negate F
FQ_FN ptR2 pt_r2_cneg(u32 m, const ptR2& P) {
  ptR2 R;
  R.N = fp2_select(m, P.D, P.N); R.D = fp2_select(m, P.N, P.D); R.E = P.E;
  R.F.re = fp_set(P.F.re.v[0] ^ m, P.F.re.v[1] ^ m, P.F.re.v[2] ^ m, P.F.re.v[3] ^ (m & FQ_P3));
  R.F.im = fp_set(P.F.im.v[0] ^ m, P.F.im.v[1] ^ m, P.F.im.v[2] ^ m, P.F.im.v[3] ^ (m & FQ_P3));
  return R;
}

// curve4q.py:138-152.  R1/R4 -> R1, 4 S + 3 M (the reference's multiplication by the constant 2 is a rotation here)
FQ_FN void pt_dbl(ptR1& Q) {
  fp2 A = fp2_sqr(Q.X); FQ_SCHED_FENCE();
  fp2 B = fp2_sqr(Q.Y); FQ_SCHED_FENCE();
  fp2 C = fp2_dbl(fp2_sqr(Q.Z)); FQ_SCHED_FENCE();
  fp2 D = fp2_add(A, B);
  fp2 E = fp2_sub(fp2_sqr(fp2_add(Q.X, Q.Y)), D); FQ_SCHED_FENCE();
  fp2 F = fp2_sub(B, A);
  fp2 G = fp2_sub(C, F);
  fp2b Gp = fp2_prep(G), Fp = fp2_prep(F);
  Q.X = fp2_mul_prep(E, Gp); FQ_SCHED_FENCE();
  Q.Y = fp2_mul_prep(D, Fp); FQ_SCHED_FENCE();
  Q.Z = fp2_mul_prep(F, Gp); FQ_SCHED_FENCE();
  Q.Ta = E; Q.Tb = D;
}

// curve4q.py:155-171.  R3 (prepared) + R2 -> R1, 7 M
FQ_FN ptR1 pt_add_core(const ptR3p& P, const ptR2& S) {
  fp2 A = fp2_mul_prep(S.D, P.D); FQ_SCHED_FENCE();
  fp2 B = fp2_mul_prep(S.N, P.N); FQ_SCHED_FENCE();
  fp2 C = fp2_mul_prep(S.F, P.F); FQ_SCHED_FENCE();
  fp2 D = fp2_mul_prep(S.E, P.E); FQ_SCHED_FENCE();
  fp2 E = fp2_sub(B, A), F = fp2_sub(D, C), G = fp2_add(D, C), H = fp2_add(B, A);
  fp2b Fp = fp2_prep(F), Gp = fp2_prep(G);
  ptR1 R;
  R.X = fp2_mul_prep(E, Fp); FQ_SCHED_FENCE();
  R.Y = fp2_mul_prep(H, Gp); FQ_SCHED_FENCE();
  R.Z = fp2_mul_prep(G, Fp); FQ_SCHED_FENCE();
  R.Ta = E; R.Tb = H;
  return R;
}
// curve4q.py:174-175
FQ_FN ptR1 pt_add(const ptR1& Q, const ptR2& S) { return pt_add_core(pt_r1_to_r3p(Q), S); }

// Out-of-line copies of the group law for the once-per-row setup code (see fp2.cuh); operands pass through memory.
struct ptR3 { fp2 N, D, E, F; };        // R3 = (X+Y, Y-X, Z, T), not prepared
#ifdef FQ_SMALL_POINT_OPS
// experiment (tools/kexp/prep_ab.cu): the out-of-line group law built from calls to the out-of-line field routines -- small
// bodies, more calls: 4 % slower (3.15 vs 3.03 ms per 2^20 rows)
FQ_CALL ptR1 pt_dbl_v(ptR1 Q) {
  fp2 A = fp2_sqr_c(Q.X), B = fp2_sqr_c(Q.Y), C = fp2_dbl(fp2_sqr_c(Q.Z));
  fp2 D = fp2_add(A, B), E = fp2_sub(fp2_sqr_c(fp2_add(Q.X, Q.Y)), D), F = fp2_sub(B, A), G = fp2_sub(C, F);
  ptR1 R; R.X = fp2_mul_c(E, G); R.Y = fp2_mul_c(D, F); R.Z = fp2_mul_c(F, G); R.Ta = E; R.Tb = D;
  return R;
}
#else
FQ_CALL ptR1 pt_dbl_v(ptR1 q) { pt_dbl(q); return q; }
#endif
FQ_FN void pt_dbl_c(ptR1* Q) { *Q = pt_dbl_v(*Q); }
FQ_FN void pt_r1_to_r3_c(ptR3* R, const ptR1* P) {                                     // curve4q.py:119-126
  R->N = fp2_add(P->X, P->Y); R->D = fp2_sub(P->Y, P->X); R->E = P->Z; R->F = fp2_mul_c(P->Ta, P->Tb);
}
#ifdef FQ_SMALL_POINT_OPS
FQ_CALL ptR1 pt_add_core_v(ptR3 P, ptR2 S) {
  fp2 A = fp2_mul_c(S.D, P.D), B = fp2_mul_c(S.N, P.N), C = fp2_mul_c(S.F, P.F), D = fp2_mul_c(S.E, P.E);
  fp2 E = fp2_sub(B, A), F = fp2_sub(D, C), G = fp2_add(D, C), H = fp2_add(B, A);
  ptR1 R; R.X = fp2_mul_c(E, F); R.Y = fp2_mul_c(H, G); R.Z = fp2_mul_c(G, F); R.Ta = E; R.Tb = H;
  return R;
}
#else
FQ_CALL ptR1 pt_add_core_v(ptR3 P, ptR2 S) {
  ptR3p Pp; Pp.N = fp2_prep(P.N); Pp.D = fp2_prep(P.D); Pp.E = fp2_prep(P.E); Pp.F = fp2_prep(P.F);
  return pt_add_core(Pp, S);
}
#endif
FQ_FN void pt_add_core_c(ptR1* out, const ptR3* P, const ptR2* S) { *out = pt_add_core_v(*P, *S); }
FQ_FN void pt_r1_to_r2_c(ptR2* R, const ptR1* P) {
  R->N = fp2_add(P->X, P->Y); R->D = fp2_sub(P->Y, P->X); R->E = fp2_dbl(P->Z);
  R->F = fp2_mul_c(fp2_mul_c(P->Ta, P->Tb), curve_2d());
}

// curve4q.py:450-455.  [392]P = 8 * 49 P: DBL, ADD, 4 DBL, ADD, 3 DBL
FQ_FN ptR1 pt_clear_cofactor(const fp2& x, const fp2& y) {
  ptR1 Q = pt_from_affine(x, y);
  ptR2 B; pt_r1_to_r2_c(&B, &Q);
  FQ_NOUNROLL
  for (int i = 0; i < 10; i++) {              // steps 1 and 6 are the additions
    if (i == 1 || i == 6) { ptR3 Q3; pt_r1_to_r3_c(&Q3, &Q); pt_add_core_c(&Q, &Q3, &B); }
    else pt_dbl_c(&Q);
  }
  return Q;
}

// curve4q.py:103-106; returns canonical coordinates
FQ_FN void pt_to_affine(const ptR1& P, fp2& x, fp2& y) {
  fp2b Zi = fp2_prep(fp2_inv(P.Z));
  x = fp2_canon(fp2_mul_prep(P.X, Zi)); y = fp2_canon(fp2_mul_prep(P.Y, Zi));
}
