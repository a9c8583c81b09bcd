// hostsim.cpp -- TEST INFRASTRUCTURE ONLY.  Compiles the device headers of fourq_b200/csrc with -DFQ_HOSTSIM, where every
// PTX primitive of arith.cuh is emulated instruction by instruction (including the carry flag), and exposes the per-row
// routines of rows.cuh to the CPU tests.  This lets `pytest -m "not gpu"` check the exact limb-level algorithms against
// the golden vectors without a GPU.  It is NOT linked into, loaded by, or a fallback for libfourq_b200.so.
#define FQ_HOSTSIM 1
#include <cstring>
#include <cstddef>
#include "../../fourq_b200/csrc/rows.cuh"
#include "../../fourq_b200/csrc/x25519.cuh"
#include "../../fourq_b200/csrc/endo.cuh"
#include "../../fourq_b200/csrc/comb.cuh"
#include "../../fourq_b200/csrc/batchinv.cuh"
#include <vector>

namespace fqsim { thread_local u32 cc = 0; }

static u32 g_tabs[1024];
static bool g_tabs_ready = false;
static void ensure_tabs() {
  if (!g_tabs_ready) { uint4 scratch[56]; row_build_base_tables(g_tabs, scratch); g_tabs_ready = true; }
}

// ---------------------------------------------------------------- the kernels' data flow, thread by thread
// Projective results of a batch as the kernels hand them over: (X, Y, Z) component-major [6][npad], one meta word per row.
struct SimScratch {
  std::vector<uint4> R; std::vector<u32> meta; size_t npad;
  explicit SimScratch(size_t n) : npad((n + 127) / 128 * 128) { R.resize(6 * npad); meta.assign(npad, 0); }
  void put(size_t row, const ptR1& P, u32 st) {
    stq4(&R[row], P.X.re); stq4(&R[npad + row], P.X.im); stq4(&R[2 * npad + row], P.Y.re); stq4(&R[3 * npad + row], P.Y.im);
    stq4(&R[4 * npad + row], P.Z.re); stq4(&R[5 * npad + row], P.Z.im);
    meta[row] = st << 8;
  }
};
// k_dh_finish (kernels_dh.cuh): grid of 64-thread CTAs, thread t owns rows t, t + stride, ... (rows_per_thread of them)
template <bool AFFINE, bool CHECK> static void sim_finish(const SimScratch& sc, uint8_t* out, uint8_t* status, size_t n, int rows_per_thread) {
  const size_t groups = (n + rows_per_thread - 1) / rows_per_thread, threads = (groups + 63) / 64 * 64;
  for (size_t t = 0; t < threads; t++) {
    FinishIO<AFFINE, CHECK> io;
    io.R = sc.R.data(); io.meta = sc.meta.data(); io.npad = sc.npad; io.out = reinterpret_cast<uint4*>(out); io.status = status; io.n = n;
    io.stride = threads; io.t = t;
    batch_invert<Fp2Ops>(io, rows_per_thread);
  }
}
// k_dh_prep -> k_dh_ladder -> k_dh_finish for one batch (row_dh_setup / row_dh_loop of rows.cuh, FinishIO of batchinv.cuh)
template <bool ENDO, bool AFFINE> static int sim_dh_pipeline(const uint8_t* k, const uint8_t* pt, uint8_t* out, uint8_t* status, size_t n) {
  SimScratch sc(n);
  uint4 tab[56]; TabView T; T.base = tab; T.stride = 1;
  for (size_t i = 0; i < n; i++) {
    u32 wk[8], wp[16]; memcpy(wk, k + 32 * i, 32); memcpy(wp, pt + (AFFINE ? 64 : 32) * i, AFFINE ? 64 : 32);
    DhState D;
    u32 st = row_dh_setup<ENDO, AFFINE>(wk, wp, T, D);
    sc.put(i, row_dh_loop<ENDO>(T, D), st);
  }
  sim_finish<AFFINE, true>(sc, out, status, n, FQ_BATCHINV_ROWS);
  return 0;
}

extern "C" {

int sim_fp2_op(int op, const uint8_t* a, const uint8_t* b, uint8_t* out, size_t n) {
  for (size_t i = 0; i < n; i++) {
    u32 wa[8], wb[8] = {0}, wo[8];
    memcpy(wa, a + 32 * i, 32);
    if (b) memcpy(wb, b + 32 * i, 32);
    switch (op) {
      case FQ_OP_MUL: row_fp2_op<FQ_OP_MUL>(wa, wb, wo); break;
      case FQ_OP_SQR: row_fp2_op<FQ_OP_SQR>(wa, wb, wo); break;
      case FQ_OP_INV: row_fp2_op<FQ_OP_INV>(wa, wb, wo); break;
      case FQ_OP_ADD: row_fp2_op<FQ_OP_ADD>(wa, wb, wo); break;
      case FQ_OP_SUB: row_fp2_op<FQ_OP_SUB>(wa, wb, wo); break;
      case FQ_OP_NEG: row_fp2_op<FQ_OP_NEG>(wa, wb, wo); break;
      case FQ_OP_CONJ: row_fp2_op<FQ_OP_CONJ>(wa, wb, wo); break;
      default: return -1;
    }
    memcpy(out + 32 * i, wo, 32);
  }
  return 0;
}

// GF(p) ops of the fq_fp_op entry point: op = FQ_FPOP_*, 16-byte rows
int sim_fp_row_op(int op, const uint8_t* a, const uint8_t* b, uint8_t* out, size_t n) {
  for (size_t i = 0; i < n; i++) {
    u32 wa[4], wb[4] = {0, 0, 0, 0}, wo[4];
    memcpy(wa, a + 16 * i, 16);
    if (b) memcpy(wb, b + 16 * i, 16);
    switch (op) {
      case FQ_FPOP_MUL: row_fp_op<FQ_FPOP_MUL>(wa, wb, wo); break;
      case FQ_FPOP_SQR: row_fp_op<FQ_FPOP_SQR>(wa, wb, wo); break;
      case FQ_FPOP_INV: row_fp_op<FQ_FPOP_INV>(wa, wb, wo); break;
      case FQ_FPOP_ADD: row_fp_op<FQ_FPOP_ADD>(wa, wb, wo); break;
      case FQ_FPOP_SUB: row_fp_op<FQ_FPOP_SUB>(wa, wb, wo); break;
      case FQ_FPOP_NEG: row_fp_op<FQ_FPOP_NEG>(wa, wb, wo); break;
      case FQ_FPOP_INVSQRT: row_fp_op<FQ_FPOP_INVSQRT>(wa, wb, wo); break;
      default: return -1;
    }
    memcpy(out + 16 * i, wo, 16);
  }
  return 0;
}

// which: 0 = inv, 1 = invsqrt, 2 = dbl, 3 = half; 16-byte rows, input tight
int sim_fp_op(int which, const uint8_t* a, uint8_t* out, size_t n) {
  for (size_t i = 0; i < n; i++) {
    u32 w[4]; memcpy(w, a + 16 * i, 16);
    fp x = fp_from_u128(fp_set(w[0], w[1], w[2], w[3]));
    fp r = which == 0 ? fp_inv(x) : which == 1 ? fp_invsqrt(x) : which == 2 ? fp_dbl(x) : fp_half(x);
    r = fp_canon(r);
    memcpy(out + 16 * i, r.v, 16);
  }
  return 0;
}

int sim_decode(const uint8_t* enc, uint8_t* xy, uint8_t* status, size_t n) {
  for (size_t i = 0; i < n; i++) {
    u32 we[8], wo[16]; memcpy(we, enc + 32 * i, 32);
    status[i] = (uint8_t)row_decode(we, wo);
    memcpy(xy + 64 * i, wo, 64);
  }
  return 0;
}
int sim_decode_spec(const uint8_t* enc, uint8_t* xy, uint8_t* status, size_t n) {
  for (size_t i = 0; i < n; i++) {
    u32 we[8], wo[16]; memcpy(we, enc + 32 * i, 32);
    status[i] = (uint8_t)row_decode<true>(we, wo);
    memcpy(xy + 64 * i, wo, 64);
  }
  return 0;
}
int sim_encode(const uint8_t* xy, uint8_t* enc, size_t n) {
  for (size_t i = 0; i < n; i++) {
    u32 wi[16], wo[8]; memcpy(wi, xy + 64 * i, 64);
    row_encode(wi, wo);
    memcpy(enc + 32 * i, wo, 32);
  }
  return 0;
}
int sim_dh(const uint8_t* k, const uint8_t* enc, uint8_t* out, uint8_t* status, size_t n) { return sim_dh_pipeline<false, false>(k, enc, out, status, n); }
int sim_dh_affine(const uint8_t* k, const uint8_t* xy, uint8_t* out, uint8_t* status, size_t n) { return sim_dh_pipeline<false, true>(k, xy, out, status, n); }
int sim_dh_endo(const uint8_t* k, const uint8_t* enc, uint8_t* out, uint8_t* status, size_t n) { return sim_dh_pipeline<true, false>(k, enc, out, status, n); }
int sim_dh_endo_affine(const uint8_t* k, const uint8_t* xy, uint8_t* out, uint8_t* status, size_t n) { return sim_dh_pipeline<true, true>(k, xy, out, status, n); }

// k_fixed_base + k_dh_finish: bit 0 of dh = [392 k]G with the neutral check, bit 1 = MUL_endo instead of MUL_windowed
int sim_fixed_base(int dh, const uint8_t* k, uint8_t* out, uint8_t* status, size_t n) {
  ensure_tabs();
  const int endo = dh >> 1; dh &= 1;
  SimScratch sc(n);
  for (size_t i = 0; i < n; i++) {
    u32 wk[8]; memcpy(wk, k + 32 * i, 32);
    uint4* tab = (uint4*)(g_tabs + (endo ? 512 : 0) + (dh ? 256 : 0));
    sc.put(i, endo ? row_fixed_base_r1<true>(wk, tab) : row_fixed_base_r1<false>(wk, tab), 0);
  }
  if (dh) sim_finish<false, true>(sc, out, status, n, FQ_BATCHINV_ROWS); else sim_finish<false, false>(sc, out, status, n, FQ_BATCHINV_ROWS);
  return 0;
}
// k_comb + k_dh_finish (per-digit fixed-base tables, comb.cuh): dh = 0 -> [k]G, 1 -> [392 k]G with the neutral check
static u32 g_comb[2 * FQ_COMB_WORDS];
static bool g_comb_ready = false;
int sim_comb(int dh, const uint8_t* k, uint8_t* out, uint8_t* status, size_t n) {
  if (!g_comb_ready) {
    for (int which = 0; which < 2; which++)
      for (int i = 0; i < FQ_COMB_DIGITS; i++) comb_build_digit(which, i, g_comb + which * FQ_COMB_WORDS + i * FQ_COMB_DIGIT_WORDS);
    g_comb_ready = true;
  }
  SimScratch sc(n);
  for (size_t i = 0; i < n; i++) {
    u32 wk[8]; memcpy(wk, k + 32 * i, 32);
    scal sk; for (int j = 0; j < 8; j++) sk.v[j] = wk[j];
    sc.put(i, mul_comb(sk, g_comb + (dh ? FQ_COMB_WORDS : 0)), 0);
  }
  if (dh) sim_finish<false, true>(sc, out, status, n, FQ_BATCHINV_ROWS); else sim_finish<false, false>(sc, out, status, n, FQ_BATCHINV_ROWS);
  return 0;
}
// k_dh_finish alone on caller-supplied projective rows: xyz = X | Y | Z (96 bytes per row, any limbs), st_in = status so far
int sim_finish_rows(const uint8_t* xyz, const uint8_t* st_in, uint8_t* out, uint8_t* status, size_t n, int rows_per_thread, int check_neutral) {
  SimScratch sc(n);
  for (size_t i = 0; i < n; i++) {
    u32 w[24]; memcpy(w, xyz + 96 * i, 96);
    ptR1 P; P.X = row_load_fp2(w); P.Y = row_load_fp2(w + 8); P.Z = row_load_fp2(w + 16); P.Ta = P.Tb = fp2_zero();
    sc.put(i, P, st_in[i]);
  }
  if (check_neutral) sim_finish<false, true>(sc, out, status, n, rows_per_thread); else sim_finish<false, false>(sc, out, status, n, rows_per_thread);
  return 0;
}
// k_fp2_inv_batched (kernels.cu): 64-thread CTAs, rows_per_thread rows per inversion
int sim_fp2_inv_batched(const uint8_t* a, uint8_t* out, size_t n, int rows_per_thread) {
  const size_t groups = (n + rows_per_thread - 1) / rows_per_thread, threads = (groups + 63) / 64 * 64;
  for (size_t t = 0; t < threads; t++) {
    Fp2InvIO io; io.a = reinterpret_cast<const uint4*>(a); io.out = reinterpret_cast<uint4*>(out); io.n = n; io.stride = threads; io.t = t;
    batch_invert<Fp2Ops>(io, rows_per_thread);
  }
  return 0;
}
// k_x25519 + k_x25519_finish (x25519.cu)
int sim_x25519_batched(const uint8_t* k, const uint8_t* u, uint8_t* out, size_t n, int rows_per_thread) {
  const size_t npad = (n + 127) / 128 * 128;
  std::vector<uint4> scratch(4 * npad);
  for (size_t i = 0; i < n; i++) {
    u32 wk[8], wu[8]; memcpy(wk, k + 32 * i, 32); memcpy(wu, u + 32 * i, 32);
    f25 x2, z2;
    x25519_ladder(wk, wu, x2, z2);
    st_f25(&scratch[i], npad, x2); st_f25(&scratch[2 * npad + i], npad, z2);
  }
  const size_t groups = (n + rows_per_thread - 1) / rows_per_thread, threads = (groups + 63) / 64 * 64;
  for (size_t t = 0; t < threads; t++) {
    X25519FinIO io; io.scratch = scratch.data(); io.npad = npad; io.out = reinterpret_cast<uint4*>(out); io.n = n; io.stride = threads; io.t = t;
    batch_invert<F25Ops>(io, rows_per_thread);
  }
  return 0;
}
// k_f25_op / k_f25_inv_batched (x25519.cu): GFp25519 ops on 32-byte rows, op = FQ_F25OP_*
int sim_f25_op(int op, const uint8_t* a, const uint8_t* b, uint8_t* out, size_t n) {
  for (size_t i = 0; i < n; i++) {
    u32 wa[8], wb[8] = {0}, wo[8]; memcpy(wa, a + 32 * i, 32);
    if (b) memcpy(wb, b + 32 * i, 32);
    switch (op) {
      case FQ_F25OP_MUL: row_f25_op<FQ_F25OP_MUL>(wa, wb, wo); break;
      case FQ_F25OP_SQR: row_f25_op<FQ_F25OP_SQR>(wa, wb, wo); break;
      case FQ_F25OP_INV: row_f25_op<FQ_F25OP_INV>(wa, wb, wo); break;
      case FQ_F25OP_ADD: row_f25_op<FQ_F25OP_ADD>(wa, wb, wo); break;
      case FQ_F25OP_SUB: row_f25_op<FQ_F25OP_SUB>(wa, wb, wo); break;
      default: return 1;
    }
    memcpy(out + 32 * i, wo, 32);
  }
  return 0;
}
int sim_f25_inv_batched(const uint8_t* a, uint8_t* out, size_t n, int rows_per_thread) {
  const size_t groups = (n + rows_per_thread - 1) / rows_per_thread, threads = (groups + 63) / 64 * 64;
  for (size_t t = 0; t < threads; t++) {
    F25InvIO io; io.a = reinterpret_cast<const uint4*>(a); io.out = reinterpret_cast<uint4*>(out); io.n = n; io.stride = threads; io.t = t;
    batch_invert<F25Ops>(io, rows_per_thread);
  }
  return 0;
}
// digits of the recoding: idx[62], neg[62] for i = 61..0 (in pop order), plus the reduced odd scalar (32 bytes)
int sim_recode(const uint8_t* k, uint8_t* idx, uint8_t* neg, uint8_t* reduced) {
  u32 wk[8]; memcpy(wk, k, 32);
  scal r = scal_reduce_odd(row_load_scalar(wk));
  memcpy(reduced, r.v, 32);
  scal S = scal_digits_init(r);
  for (int i = 0; i < 62; i++) { u32 a, b; scal_next_digit(S, a, b); idx[i] = (uint8_t)a; neg[i] = (uint8_t)(b & 1); }
  return 0;
}
// endomorphism pieces: which 0 = phi, 1 = psi on affine 64-byte points -> affine
int sim_endo_map(int which, const uint8_t* xy, uint8_t* out, size_t n) {
  for (size_t i = 0; i < n; i++) {
    u32 wi[16], wo[16]; memcpy(wi, xy + 64 * i, 64);
    fp2 x = fp2_canon(row_load_fp2(wi)), y = fp2_canon(row_load_fp2(wi + 8)), ox, oy;
    ptR1 P = pt_from_affine(x, y);
    ptR1 R = which == 0 ? endo_phi(P) : endo_psi(P);
    pt_to_affine(R, ox, oy);
    row_store_fp2(wo, ox); row_store_fp2(wo + 8, oy);
    memcpy(out + 64 * i, wo, 64);
  }
  return 0;
}
// decompose: k (32 B) -> 4 x u64; recode: digits idx[65], sign[65] in index order 0..64
int sim_endo_scalar(const uint8_t* k, uint64_t* v, uint8_t* idx, uint8_t* sign) {
  u32 wk[8]; memcpy(wk, k, 32);
  scal4 d = endo_decompose(row_load_scalar(wk));
  for (int j = 0; j < 4; j++) v[j] = d.v[j];
  scal S; u32 d64 = endo_recode(d, S);
  idx[64] = (uint8_t)d64; sign[64] = 1;
  for (int i = 63; i >= 0; i--) { u32 a, b; endo_next_digit(S, a, b); idx[i] = (uint8_t)a; sign[i] = (uint8_t)((b & 1) ^ 1); }
  return 0;
}
int sim_x25519(const uint8_t* k, const uint8_t* u, uint8_t* out, size_t n) {
  for (size_t i = 0; i < n; i++) {
    u32 wk[8], wu[8], wo[8]; memcpy(wk, k + 32 * i, 32); memcpy(wu, u + 32 * i, 32);
    row_x25519(wk, wu, wo);
    memcpy(out + 32 * i, wo, 32);
  }
  return 0;
}
// fields.py GFp2.invsqrt :201-230 (rows.cuh row_fp2_invsqrt)
int sim_fp2_invsqrt(const uint8_t* a, uint8_t* out, size_t n) {
  for (size_t i = 0; i < n; i++) { u32 wa[8], wo[8]; memcpy(wa, a + 32 * i, 32); row_fp2_invsqrt(wa, wo); memcpy(out + 32 * i, wo, 32); }
  return 0;
}
// fields.py GFp.select :59-64 (halves = 1, 16-byte rows) / GFp2.select :236-238 (halves = 2, 32-byte rows)
int sim_select(int halves, const uint8_t* c, const uint8_t* x, const uint8_t* y, uint8_t* out, size_t n) {
  for (size_t i = 0; i < n; i++) {
    u32 wx[8], wy[8], wo[8];
    memcpy(wx, x + 16 * halves * i, 16 * halves); memcpy(wy, y + 16 * halves * i, 16 * halves);
    if (halves == 2) row_select<2>(c[i], wx, wy, wo); else row_select<1>(c[i], wx, wy, wo);
    memcpy(out + 16 * halves * i, wo, 16 * halves);
  }
  return 0;
}
// curve4q.py PointOnCurve :23-29 on x | y rows (kernels.cu k_on_curve)
int sim_on_curve(const uint8_t* xy, uint8_t* ok, size_t n) {
  for (size_t i = 0; i < n; i++) {
    u32 wi[16]; memcpy(wi, xy + 64 * i, 64);
    fp2 x = fp2_canon(row_load_fp2(wi)), y = fp2_canon(row_load_fp2(wi + 8));
    ok[i] = pt_on_curve(x, y) ? 1 : 0;
  }
  return 0;
}
}  // extern "C"
