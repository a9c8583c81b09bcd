// dh.cuh -- one thread = one scalar multiplication.  Table handling, the fixed-window main loop of MUL_windowed
// (curve4q.py:188-235), and the Diffie-Hellman shell DH_core (curve4q.py:446-462).
//
// Table residency.  T[i] = [2i+1]P in R2 is 8 x 4 x 32 B = 1 KiB per thread.  Entries 0..6 live in shared memory as
// 128-bit words laid out [entry][quad][thread] (consecutive threads -> consecutive 16 B: conflict-free LDS.128/STS.128),
// entry 7 stays in registers: 7 x 128 B x 128 threads = 112 KiB per CTA, two CTAs per SM.
// Selection issues the loads of ALL entries and keeps one (tab_select: strict scan by default); the sign is applied with masks.
#pragma once
#include "codec.cuh"
#include "scalar.cuh"

// per-thread view of the shared-memory table: quad q of entry e is base[(e*8+q)*stride]
struct TabView { uint4* base; u32 stride; };

FQ_FN void tab_put(const TabView& T, int e, int q, const fp& a) { T.base[(e * 8 + q) * T.stride] = make_uint4(a.v[0], a.v[1], a.v[2], a.v[3]); }
FQ_FN fp tab_get(const TabView& T, int e, int q) { uint4 w = T.base[(e * 8 + q) * T.stride]; return fp_set(w.x, w.y, w.z, w.w); }

FQ_FN void tab_store(const TabView& T, int e, const ptR2& P) {
  tab_put(T, e, 0, P.N.re); tab_put(T, e, 1, P.N.im); tab_put(T, e, 2, P.D.re); tab_put(T, e, 3, P.D.im);
  tab_put(T, e, 4, P.E.re); tab_put(T, e, 5, P.E.im); tab_put(T, e, 6, P.F.re); tab_put(T, e, 7, P.F.im);
}
FQ_FN ptR2 tab_load(const TabView& T, int e) {
  ptR2 P;
  P.N = fp2_set(tab_get(T, e, 0), tab_get(T, e, 1)); P.D = fp2_set(tab_get(T, e, 2), tab_get(T, e, 3));
  P.E = fp2_set(tab_get(T, e, 4), tab_get(T, e, 5)); P.F = fp2_set(tab_get(T, e, 6), tab_get(T, e, 7));
  return P;
}

// Constant-time table selection.  Every thread issues the loads of ALL entries in the same order from addresses that do not
// depend on the digit; the digit only decides what each load keeps.  Two variants, compiled side by side and chosen per
// launch (template parameter STRICT; fq_set_select_mode / FQ_STRICT_SELECT at run time):
//   strict scan (library default)   every lane loads every entry and each word goes through one SEL per entry (56 LDS.128 + 224
//                            SEL): no data-dependent memory activity of any kind; ladder time flat to 0.05 % over scalar
//                            distributions (profiles/r02_ct_timing.jsonl).  This is what draft-ladd-cfrg-4q.md:653-656, :753-755 ask for.
//   masked loads (opt-in)    `@p ld.shared.v4` -- a lane whose predicate is off transfers nothing and keeps its registers, so
//                            the select costs no ALU instruction at all (56 LDS.128 per select): the ladder is 1 % faster, the
//                            fixed-base comb kernel (short iterations) 8 %.  The number of shared-memory wavefronts of a load then
//                            depends on which lanes are on: a batch in which all 32 rows of every warp pick the same entry runs the
//                            ladder 0.4 % faster than one in which they differ (profiles/r01_ct_timing.jsonl).  For non-secret
//                            scalars only.
// No secret-dependent branch or address in either; the layout [entry][quad][thread] keeps lane L on banks 4L..4L+3 for
// every entry, so there is no digit-dependent bank conflict.  (The template default below only concerns code that does not
// pass STRICT explicitly: the CPU simulation and the experiments in tools/kexp.)
#ifdef FQ_STRICT_SELECT
#define FQ_STRICT_DEFAULT true
#else
#define FQ_STRICT_DEFAULT false
#endif

// w = c ? *p : w for one 16-byte quad; p points into shared memory
template <bool STRICT = FQ_STRICT_DEFAULT> FQ_FN void quad_take(const uint4* p, fp& w, bool c) {
#if defined(FQ_HOSTSIM)
  if (c) w = fp_set(p->x, p->y, p->z, p->w);
#else
  if (STRICT) {
    u32 a0, a1, a2, a3;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a0), "=r"(a1), "=r"(a2), "=r"(a3) : "r"((u32)__cvta_generic_to_shared(p)));
    w.v[0] = c ? a0 : w.v[0]; w.v[1] = c ? a1 : w.v[1]; w.v[2] = c ? a2 : w.v[2]; w.v[3] = c ? a3 : w.v[3];
  } else {
    asm volatile("{ .reg .pred p; setp.ne.u32 p, %5, 0; @p ld.shared.v4.u32 {%0,%1,%2,%3}, [%4]; }"
                 : "+r"(w.v[0]), "+r"(w.v[1]), "+r"(w.v[2]), "+r"(w.v[3]) : "r"((u32)__cvta_generic_to_shared(p)), "r"((u32)c));
  }
#endif
}
template <bool STRICT = FQ_STRICT_DEFAULT> FQ_FN void tab_take(const TabView& T, int e, int q, fp& w, bool c) {
  quad_take<STRICT>(T.base + (e * 8 + q) * T.stride, w, c);
}
template <bool STRICT = FQ_STRICT_DEFAULT> FQ_FN void tab_take_r2(const TabView& T, int e, ptR2& w, bool c) {
  tab_take<STRICT>(T, e, 0, w.N.re, c); tab_take<STRICT>(T, e, 1, w.N.im, c); tab_take<STRICT>(T, e, 2, w.D.re, c); tab_take<STRICT>(T, e, 3, w.D.im, c);
  tab_take<STRICT>(T, e, 4, w.E.re, c); tab_take<STRICT>(T, e, 5, w.E.im, c); tab_take<STRICT>(T, e, 6, w.F.re, c); tab_take<STRICT>(T, e, 7, w.F.im, c);
}

// T[idx]: starts from the register-resident entry 7 and scans entries 0..6 in shared memory
template <bool STRICT = FQ_STRICT_DEFAULT> FQ_FN ptR2 tab_select(const TabView& T, const ptR2& T7, u32 idx) {
  ptR2 S = T7;
  FQ_UNROLL
  for (int e = 0; e < 7; e++) tab_take_r2<STRICT>(T, e, S, idx == (u32)e);
  return S;
}

// curve4q.py:179-185: T[i] = [2i+1]P, i = 0..7; returns T[7]
FQ_FN ptR2 tab_build(const TabView& T, const ptR1& P) {
  ptR1 P2 = P;
  pt_dbl_c(&P2);
  ptR3 P23; pt_r1_to_r3_c(&P23, &P2);
  ptR2 Ti; pt_r1_to_r2_c(&Ti, &P);
  tab_store(T, 0, Ti);
  FQ_NOUNROLL
  for (int i = 1; i < 8; i++) {
    ptR1 S; pt_add_core_c(&S, &P23, &Ti);
    pt_r1_to_r2_c(&Ti, &S);
    if (i < 7) tab_store(T, i, Ti);
  }
  return Ti;
}

// The scalar-dependent plan of a multiplication: the packed digit register and the table index of the leading digit.
struct MulPlan { scal S; u32 first; };

// curve4q.py:216-226
FQ_FN MulPlan plan_windowed(const scal& k) {
  MulPlan pl; pl.S = scal_digits_init(scal_reduce_odd(k)); pl.first = 0u;          // digit 62 is always +1
  return pl;
}
// curve4q.py:229-233 with the digits of scalar.cuh.  SELECT(idx) returns T[idx] in constant time.
template <class SELECT> FQ_FN ptR1 loop_windowed(MulPlan& pl, SELECT select) {
  ptR1 Q = pt_r2_to_r4(select(pl.first));
  FQ_NOUNROLL
  for (int i = 61; i >= 0; i--) {
    FQ_NOUNROLL
    for (int j = 0; j < 4; j++) pt_dbl(Q);
    u32 idx, neg;
    scal_next_digit(pl.S, idx, neg);
    Q = pt_add(Q, pt_r2_cneg(neg, select(idx)));
  }
  return Q;
}
// curve4q.py:188-235
template <class SELECT> FQ_FN ptR1 mul_windowed(const scal& k, SELECT select) {
  MulPlan pl = plan_windowed(k);
  return loop_windowed(pl, select);
}

template <bool STRICT = FQ_STRICT_DEFAULT> struct SelectShared {
  TabView T; ptR2 T7;
  FQ_MFN ptR2 operator()(u32 idx) const { return tab_select<STRICT>(T, T7, idx); }
};

// DH_core (curve4q.py:446-462) after the point has been validated, in three phases so that a kernel can keep all warps
// of a CTA in the same phase (instruction-cache locality): setup ([392]P, table, scalar plan), loop ([k]Q), finish
// (affine, neutral check).
struct DhState { ptR2 T7; MulPlan plan; };

FQ_FN void dh_setup_windowed(const scal& k, const fp2& x, const fp2& y, const TabView& T, DhState& D) {
  ptR1 Q = pt_clear_cofactor(x, y);
  D.T7 = tab_build(T, Q);
  D.plan = plan_windowed(k);
}
template <bool STRICT = FQ_STRICT_DEFAULT> FQ_FN ptR1 dh_loop_windowed(const TabView& T, DhState& D) {
  SelectShared<STRICT> sel; sel.T = T; sel.T7 = D.T7;
  return loop_windowed(D.plan, sel);
}

// fixed base (the reference's table_windowed / table_endo shape): ONE table of 8 entries x 8 quads shared by all threads
// of a CTA, held in shared memory as [entry][quad] (stride 1).  Every thread reads the same address (broadcast), entry 7 is
// loaded unconditionally and entries 0..6 under the digit's predicate, exactly like tab_select.
template <bool STRICT = FQ_STRICT_DEFAULT> struct SelectBroadcast {
  TabView T;
  FQ_MFN ptR2 operator()(u32 idx) const {
    ptR2 S = tab_load(T, 7);
    FQ_UNROLL
    for (int e = 0; e < 7; e++) tab_take_r2<STRICT>(T, e, S, idx == (u32)e);
    return S;
  }
};
FQ_FN void r2_to_words(const ptR2& P, u32* w) {
  const fp* f[8] = {&P.N.re, &P.N.im, &P.D.re, &P.D.im, &P.E.re, &P.E.im, &P.F.re, &P.F.im};
  FQ_UNROLL
  for (int q = 0; q < 8; q++) { FQ_UNROLL for (int j = 0; j < 4; j++) w[q * 4 + j] = f[q]->v[j]; }
}
