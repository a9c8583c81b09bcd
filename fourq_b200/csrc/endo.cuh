// endo.cuh -- the endomorphism-accelerated scalar multiplication of the reference:
// phi, psi (impl/curve4q.py:258-322), 4-dimensional decomposition and GLV-SAC recoding (curve4q.py:339-380),
// table_endo / MUL_endo (curve4q.py:385-442).  64 x (DBL + ADD) instead of 62 x (4 DBL + ADD): 58,284 instead of
// 103,836 multiply-adds per DH with bit-identical results (test_dh, curve4q.py:706-762, asserts DH_endo == DH_windowed).
#pragma once
#include "dh.cuh"

// ---------------------------------------------------------------- constants (curve4q.py:240-256)
FQ_FN fp2 endo_ctau() { return fp2_set(fp_set(0xebce74c3u, 0x74dcd57cu, 0x3afad20cu, 0x1964de2cu), fp_set(0x00000012u, 0x00000000u, 0x0000000cu, 0x00000000u)); }
FQ_FN fp2 endo_ctaudual() { return fp2_set(fp_set(0xdecdf034u, 0x9ecaa6d9u, 0x23058652u, 0x4aa740ebu), fp_set(0x00000011u, 0x00000000u, 0xfffffff4u, 0x7fffffffu)); }
FQ_FN fp2 endo_cphi0() { return fp2_set(fp_set(0xfffffff7u, 0xffffffffu, 0x00000005u, 0x00000000u), fp_set(0xef66f81au, 0x4f65536cu, 0x9182c329u, 0x2553a075u)); }
FQ_FN fp2 endo_cphi1() { return fp2_set(fp_set(0x00000007u, 0x00000000u, 0x00000005u, 0x00000000u), fp_set(0xe28296f9u, 0x334d90e9u, 0xc50c62cfu, 0x62c8caa0u)); }
FQ_FN fp2 endo_cphi2() { return fp2_set(fp_set(0x00000015u, 0x00000000u, 0x0000000fu, 0x00000000u), fp_set(0x4f1df391u, 0x2c2cb715u, 0x6c9b5c98u, 0x78df262bu)); }
FQ_FN fp2 endo_cphi3() { return fp2_set(fp_set(0x00000003u, 0x00000000u, 0x00000002u, 0x00000000u), fp_set(0xa7962ea4u, 0x92440457u, 0x1d76342au, 0x5084c649u)); }
FQ_FN fp2 endo_cphi4() { return fp2_set(fp_set(0x00000003u, 0x00000000u, 0x00000003u, 0x00000000u), fp_set(0x3aec6855u, 0xa1098c92u, 0xa7962ea4u, 0x12440457u)); }
FQ_FN fp2 endo_cphi5() { return fp2_set(fp_set(0x0000000fu, 0x00000000u, 0x0000000au, 0x00000000u), fp_set(0xc5052df3u, 0x669b21d3u, 0x8a18c59eu, 0x45919541u)); }
FQ_FN fp2 endo_cphi6() { return fp2_set(fp_set(0x00000018u, 0x00000000u, 0x00000012u, 0x00000000u), fp_set(0x8a0a5be7u, 0xcd3643a7u, 0x14318b3cu, 0x0b232a83u)); }
FQ_FN fp2 endo_cphi7() { return fp2_set(fp_set(0x00000023u, 0x00000000u, 0x00000018u, 0x00000000u), fp_set(0x5f48781au, 0x66c18303u, 0x99e2ea1au, 0x3963bc1cu)); }
FQ_FN fp2 endo_cphi8() { return fp2_set(fp_set(0x000000f0u, 0x00000000u, 0x000000aau, 0x00000000u), fp_set(0x2b5d0ef0u, 0x44e25158u, 0x0316cbe5u, 0x1f529f86u)); }
FQ_FN fp2 endo_cphi9() { return fp2_set(fp_set(0x00000befu, 0x00000000u, 0x00000870u, 0x00000000u), fp_set(0x976e2505u, 0x014d3e48u, 0xfe00375bu, 0x0fd52e9cu)); }
FQ_FN fp2 endo_cpsi1() { return fp2_set(fp_set(0x67e346efu, 0xedf07f47u, 0x83d54a02u, 0x2af99e9au), fp_set(0x0000013au, 0x00000000u, 0x000000deu, 0x00000000u)); }
FQ_FN fp2 endo_cpsi2() { return fp2_set(fp_set(0x00000143u, 0x00000000u, 0x000000e4u, 0x00000000u), fp_set(0x0e03f372u, 0x4c7deb77u, 0x99a81f03u, 0x21b8d07bu)); }
FQ_FN fp2 endo_cpsi3() { return fp2_set(fp_set(0x00000009u, 0x00000000u, 0x00000006u, 0x00000000u), fp_set(0x75e73a61u, 0x3a6e6abeu, 0x1d7d6906u, 0x4cb26f16u)); }
FQ_FN fp2 endo_cpsi4() { return fp2_set(fp_set(0xfffffff6u, 0xffffffffu, 0xfffffff9u, 0x7fffffffu), fp_set(0x8a18c59eu, 0xc5919541u, 0xe28296f9u, 0x334d90e9u)); }

struct pt3 { fp2 X, Y, Z; };

// tau, tau_dual and chi are evaluated three, three and two times per row (phi(P), psi(P), psi(phi(P))).  Inlined at every use:
// as out-of-line routines (-DFQ_ENDO_CALLS: one copy of each, arguments in registers) the prepare kernel is 1.5 % slower
// (3.07 vs 3.03 ms per 2^20 rows: 160 B more stack, tools/kexp/prep_ab.cu).
#if defined(FQ_ENDO_CALLS) && !defined(FQ_HOSTSIM)
#define FQ_ENDO_FN FQ_CALL
#else
#define FQ_ENDO_FN FQ_FN
#endif

FQ_FN fp2 fp2_two_sqr(const fp2& z) { return fp2_dbl(fp2_sqr_c(z)); }

// curve4q.py:258-267
FQ_ENDO_FN pt3 endo_tau(pt3 P) {
  fp2 A = fp2_sqr_c(P.X), B = fp2_sqr_c(P.Y);
  fp2 C = fp2_add(A, B), D = fp2_sub(A, B);
  pt3 R;
  R.X = fp2_mul_c(fp2_mul_c(fp2_mul_c(endo_ctau(), P.X), P.Y), D);
  R.Y = fp2_neg(fp2_mul_c(fp2_add(fp2_two_sqr(P.Z), D), C));
  R.Z = fp2_mul_c(C, D);
  return R;
}
// curve4q.py:269-280 -> R1
FQ_ENDO_FN ptR1 endo_tau_dual(pt3 P) {
  fp2 A = fp2_sqr_c(P.X), B = fp2_sqr_c(P.Y);
  fp2 C = fp2_add(A, B);
  ptR1 R;
  R.Ta = fp2_sub(B, A);
  fp2 D = fp2_sub(fp2_two_sqr(P.Z), R.Ta);
  R.Tb = fp2_mul_c(fp2_mul_c(endo_ctaudual(), P.X), P.Y);
  R.X = fp2_mul_c(R.Tb, C); R.Y = fp2_mul_c(R.Ta, D); R.Z = fp2_mul_c(D, C);
  return R;
}
// curve4q.py:282-302
FQ_ENDO_FN pt3 endo_upsilon(pt3 P) {
  fp2 A = fp2_mul_c(fp2_mul_c(endo_cphi0(), P.X), P.Y);
  fp2 B = fp2_mul_c(P.Y, P.Z);
  fp2 C = fp2_sqr_c(P.Y), D = fp2_sqr_c(P.Z);
  fp2 F = fp2_sqr_c(D), G = fp2_sqr_c(B), H = fp2_sqr_c(C);
  fp2 I = fp2_mul_c(endo_cphi1(), B);
  fp2 J = fp2_add(C, fp2_mul_c(endo_cphi2(), D));
  fp2 K = fp2_add(fp2_add(fp2_mul_c(endo_cphi8(), G), H), fp2_mul_c(endo_cphi9(), F));
  pt3 R;
  R.X = fp2_conj(fp2_mul_c(fp2_mul_c(A, K), fp2_mul_c(fp2_add(I, J), fp2_sub(I, J))));
  fp2 L = fp2_add(C, fp2_mul_c(endo_cphi4(), D));
  fp2 M = fp2_mul_c(endo_cphi3(), B);
  fp2 Nn = fp2_mul_c(fp2_add(L, M), fp2_sub(L, M));
  fp2 Y2 = fp2_add(fp2_add(H, fp2_mul_c(endo_cphi6(), G)), fp2_mul_c(endo_cphi7(), F));
  R.Y = fp2_conj(fp2_mul_c(fp2_mul_c(fp2_mul_c(endo_cphi5(), D), Nn), Y2));
  R.Z = fp2_conj(fp2_mul_c(fp2_mul_c(B, K), Nn));
  return R;
}
// curve4q.py:304-316
FQ_ENDO_FN pt3 endo_chi(pt3 P) {
  fp2 A = fp2_conj(P.X), B = fp2_conj(P.Y);
  fp2 C = fp2_sqr_c(fp2_conj(P.Z));
  fp2 D = fp2_sqr_c(A);
  fp2 G = fp2_mul_c(B, fp2_add(D, fp2_mul_c(endo_cpsi2(), C)));
  fp2 H = fp2_neg(fp2_add(D, fp2_mul_c(endo_cpsi4(), C)));
  pt3 R;
  R.X = fp2_mul_c(fp2_mul_c(fp2_mul_c(endo_cpsi1(), A), C), H);
  R.Y = fp2_mul_c(G, fp2_add(D, fp2_mul_c(endo_cpsi3(), C)));
  R.Z = fp2_mul_c(G, H);
  return R;
}
FQ_FN pt3 pt3_of(const ptR1& P) { pt3 R; R.X = P.X; R.Y = P.Y; R.Z = P.Z; return R; }
FQ_FN ptR1 endo_phi(const ptR1& P) { return endo_tau_dual(endo_upsilon(endo_tau(pt3_of(P)))); }   // curve4q.py:318-319
FQ_FN ptR1 endo_psi(const ptR1& P) { return endo_tau_dual(endo_chi(endo_tau(pt3_of(P)))); }       // curve4q.py:321-322

// ---------------------------------------------------------------- decomposition (curve4q.py:326-356)
// t_i = floor(L_i m / 2^256) is needed mod 2^64 only (the four results are < 2^64): limbs 8 and 9 of the 7 x 8 limb
// product, by column sums with a 3-word carry-save accumulator.
// One out-of-line body shared by the four constants (the fully inlined version was 4 x 170 straight-line instructions).
FQ_CALL u64 endo_mulhi_256(u32 l0, u32 l1, u32 l2, u32 l3, u32 l4, u32 l5, u32 l6, scal m) {
  const u32 L[7] = {l0, l1, l2, l3, l4, l5, l6};
  u32 c0 = 0, c1 = 0, c2 = 0, out8 = 0, out9 = 0;
  FQ_UNROLL
  for (int k = 0; k < 10; k++) {
    FQ_UNROLL
    for (int i = 0; i < 7; i++) {
      const int j = k - i;
      if (j >= 0 && j < 8) { c0 = mad_lo_cc(L[i], m.v[j], c0); c1 = madc_hi_cc(L[i], m.v[j], c1); c2 = addc(c2, 0); }
    }
    if (k == 8) out8 = c0;
    if (k == 9) out9 = c0;
    c0 = c1; c1 = c2; c2 = 0;
  }
  return ((u64)out9 << 32) | out8;
}

struct scal4 { u64 v[4]; };

FQ_FN scal4 endo_decompose(const scal& m) {
  // L1..L4 (curve4q.py:333-336), little-endian 32-bit limbs (L3 has 6: its 7th is 0)
  // lattice basis and offsets mod 2^64 (curve4q.py:326-337; negative entries wrapped)
  const u64 B1[4] = {0x0906ff27e0a0a196ull, 0xec9c179d3dd5d260ull, 0x07426031ecc8030full, 0xf7b08c66794619afull};
  const u64 B2[4] = {0x1d495bea84fcc2d4ull, 0xffffffffffffffffull, 0x0000000000000001ull, 0x25dbc5bc8dd167d0ull};
  const u64 B3[4] = {0x17abad1d231f0302ull, 0x02c4211ae388da51ull, 0xd1b2de3676d83b61ull, 0x0a9e6f44c02ecd97ull};
  const u64 B4[4] = {0x136e340a9108c83full, 0x3122df2dc3e0ff32ull, 0xf975b60fd557564bull, 0xe72af7876921f516ull};
  const u64 C[4] = {0x72482c5251a4559cull, 0x59f95b0add276f6cull, 0x7dd2d17c4625fa78ull, 0x6bc57def56ce8877ull};
  const u64 CP[4] = {0x85b6605ce2ad1ddbull, 0x8b1c3a38a1086e9eull, 0x7748878c1b7d50c3ull, 0x52f07576bff07d8dull};
  u64 t1 = endo_mulhi_256(0x9d1a7d4fu, 0x259686e0u, 0xe6a6bd66u, 0xf75682acu, 0xea2be5dfu, 0xfc5bb5c5u, 0x00000007u, m);
  u64 t2 = endo_mulhi_256(0xdd627afbu, 0xd1ba1d84u, 0x0f468d8du, 0x2bd23558u, 0xaa6c0f8au, 0x8fd4b04cu, 0x00000003u, m);
  u64 t3 = endo_mulhi_256(0x678c203cu, 0x9b291a33u, 0x65dca902u, 0xc42bd6c9u, 0x0bffbaf6u, 0xd038bf8du, 0x00000000u, m);
  u64 t4 = endo_mulhi_256(0x77e7fdc0u, 0x12e5666bu, 0x14983d82u, 0x81cbdc37u, 0xa22d8410u, 0x1b073877u, 0x00000003u, m);
  scal4 a;
  FQ_UNROLL
  for (int j = 0; j < 4; j++) {
    u64 x = (j == 0) ? (((u64)m.v[1] << 32) | m.v[0]) : 0ull;
    a.v[j] = x - t1 * B1[j] - t2 * B2[j] - t3 * B3[j] - t4 * B4[j];
  }
  u64 odd = 0ull - ((a.v[0] + C[0]) & 1ull);             // curve4q.py:354-355: pick a + c if its first entry is odd
  scal4 v;
  FQ_UNROLL
  for (int j = 0; j < 4; j++) v.v[j] = a.v[j] + ((C[j] & odd) | (CP[j] & ~odd));
  return v;
}

// ---------------------------------------------------------------- GLV-SAC recoding (curve4q.py:358-380)
// Digit i (0..63) is packed as nibble i of S: bit 3 = sign (1 = positive), bits 0..2 = table index.  Returns d[64].
FQ_FN u32 endo_recode(const scal4& vin, scal& S) {
  u64 v0s = vin.v[0] >> 1, v1 = vin.v[1], v2 = vin.v[2], v3 = vin.v[3];     // v0s: bit i+1 of v0 is its bit 0 in step i (0 for i = 63)
  FQ_UNROLL
  for (int w = 0; w < 8; w++) S.v[w] = 0;
  FQ_NOUNROLL
  for (int w = 0; w < 8; w++) {                  // rolled: the unrolled 64 steps were 2,000 straight-line instructions
    u32 word = 0;
    FQ_UNROLL
    for (int n = 0; n < 8; n++) {
      u32 b1 = (u32)(v0s & 1ull);
      v0s >>= 1;
      u32 e1 = (u32)(v1 & 1ull), e2 = (u32)(v2 & 1ull), e3 = (u32)(v3 & 1ull);
      word |= ((b1 << 3) | e1 | (e2 << 1) | (e3 << 2)) << (4 * n);
      u32 nb1 = b1 ^ 1u;
      v1 = (v1 >> 1) + (u64)(nb1 & e1); v2 = (v2 >> 1) + (u64)(nb1 & e2); v3 = (v3 >> 1) + (u64)(nb1 & e3);
    }
    FQ_UNROLL
    for (int i = 0; i < 7; i++) S.v[i] = S.v[i + 1];          // word w ends up in S.v[w] after the 8 rounds
    S.v[7] = word;
  }
  return (u32)(v1 + 2ull * v2 + 4ull * v3);
}
FQ_FN void endo_next_digit(scal& s, u32& idx, u32& neg) {
  u32 nib = s.v[7] >> 28;
  neg = (nib >> 3) - 1;                    // sign bit 1 = positive
  idx = nib & 7;
  FQ_UNROLL
  for (int i = 7; i > 0; i--) s.v[i] = shl_pair(s.v[i - 1], s.v[i], 4);
  s.v[0] <<= 4;
}

// ---------------------------------------------------------------- table_endo / MUL_endo
// curve4q.py:385-403: T[0] = P, T[1] = P+Q, T[2] = P+R, T[3] = P+Q+R, T[4+i] = T[i] + S with Q = phi(P), R = psi(P),
// S = psi(phi(P)); entries in R2, 0..6 to shared memory, returns T[7].
FQ_FN ptR2 endo_tab_build(const TabView& T, const ptR1& P) {
  const pt3 tP = endo_tau(pt3_of(P));        // tau(P) is the first step of both phi(P) and psi(P) (curve4q.py:318-322): evaluated once
  ptR1 Qp = endo_tau_dual(endo_upsilon(tP)); // phi(P)
  ptR3 A3;                                   // the left operand of each addition, in R3
  ptR2 T0, Ti;
  ptR1 S;
  pt_r1_to_r2_c(&T0, &P);
  tab_store(T, 0, T0);
  pt_r1_to_r3_c(&A3, &Qp);
  pt_add_core_c(&S, &A3, &T0); pt_r1_to_r2_c(&Ti, &S); tab_store(T, 1, Ti);          // T[1] = Q + P
  S = endo_tau_dual(endo_chi(tP));           // psi(P)
  pt_r1_to_r3_c(&A3, &S);
  pt_add_core_c(&S, &A3, &Ti); pt_r1_to_r2_c(&Ti, &S); tab_store(T, 3, Ti);          // T[3] = R + T[1]
  pt_add_core_c(&S, &A3, &T0); pt_r1_to_r2_c(&Ti, &S); tab_store(T, 2, Ti);          // T[2] = R + P
  S = endo_psi(Qp);
  pt_r1_to_r3_c(&A3, &S);
  FQ_NOUNROLL
  for (int i = 0; i < 4; i++) {                                                       // T[4+i] = S + T[i]
    Ti = tab_load(T, i);
    pt_add_core_c(&S, &A3, &Ti); pt_r1_to_r2_c(&Ti, &S);
    if (i < 3) tab_store(T, 4 + i, Ti);
  }
  return Ti;
}

// curve4q.py:405-436: decompose + recode
FQ_FN MulPlan plan_endo(const scal& k) {
  MulPlan pl;
  pl.first = endo_recode(endo_decompose(k), pl.S);      // s[64] = 1: the leading digit is positive
  return pl;
}
// curve4q.py:437-442
template <class SELECT> FQ_FN ptR1 loop_endo(MulPlan& pl, SELECT select) {
  ptR1 Q = pt_r2_to_r4(select(pl.first));
  FQ_NOUNROLL
  for (int i = 63; i >= 0; i--) {
    pt_dbl(Q);
    u32 idx, neg;
    endo_next_digit(pl.S, idx, neg);
    Q = pt_add(Q, pt_r2_cneg(neg, select(idx)));
  }
  return Q;
}
template <class SELECT> FQ_FN ptR1 mul_endo(const scal& k, SELECT select) {
  MulPlan pl = plan_endo(k);
  return loop_endo(pl, select);
}

// DH_endo (curve4q.py:467-468) after validation: the phases of dh.cuh with table_endo / MUL_endo
FQ_FN void dh_setup_endo(const scal& k, const fp2& x, const fp2& y, const TabView& T, DhState& D) {
  ptR1 Q = pt_clear_cofactor(x, y);
  D.T7 = endo_tab_build(T, Q);
  D.plan = plan_endo(k);
}
template <bool STRICT = FQ_STRICT_DEFAULT> FQ_FN ptR1 dh_loop_endo(const TabView& T, DhState& D) {
  SelectShared<STRICT> sel; sel.T = T; sel.T7 = D.T7;
  return loop_endo(D.plan, sel);
}
