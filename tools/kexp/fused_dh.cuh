// fused_dh.cuh -- EXPERIMENT ONLY (tools/kexp/split.cu): the whole variable-base DH row as ONE kernel, the shape the library had
// before the prepare / ladder / finish split (fourq_b200/csrc/kernels_dh.cuh).  Not compiled into libfourq_b200.so.
#pragma once
#include "../../fourq_b200/csrc/kernels_dh.cuh"
#include "../../fourq_b200/csrc/comb.cuh"

FQ_FN u32 dh_finish(const ptR1& R, fp2& ox, fp2& oy) {
  pt_to_affine(R, ox, oy);
  bool neutral = fp2_eq_canon(ox, fp2_zero()) & fp2_eq_canon(oy, fp2_one());      // curve4q.py:459
  return neutral ? FQ_ST_NEUTRAL : FQ_ST_OK;
}

// [k]B for the base point whose table is T; returns canonical affine.  MUL_windowed(k, ., table) + R1toAffine.
template <bool STRICT = FQ_STRICT_DEFAULT> FQ_FN void mul_fixed_base(const scal& k, const TabView& T, fp2& ox, fp2& oy) {
  SelectBroadcast<STRICT> sel; sel.T = T;
  ptR1 R = mul_windowed(k, sel);
  pt_to_affine(R, ox, oy);
}

template <bool AFFINE> FQ_FN u32 row_dh_finish(u32 st, const ptR1& R, u32* out) {
  fp2 ox, oy;
  u32 st2 = dh_finish(R, ox, oy);
  if (st == FQ_ST_OK) st = st2;
  if (AFFINE) { if (st == FQ_ST_OK) { row_store_fp2(out, ox); row_store_fp2(out + 8, oy); } else row_zero(out, 16); }
  else { if (st == FQ_ST_OK) pt_encode(ox, oy, out); else row_zero(out, 8); }
  return st;
}
template <bool ENDO> FQ_FN u32 row_dh(const u32* k, const u32* enc, u32* out, const TabView& T) {
  DhState D;
  u32 st = row_dh_setup<ENDO, false>(k, enc, T, D);
  return row_dh_finish<false>(st, row_dh_loop<ENDO>(T, D), out);
}
template <bool ENDO> FQ_FN u32 row_dh_affine(const u32* k, const u32* xy, u32* out, const TabView& T) {
  DhState D;
  u32 st = row_dh_setup<ENDO, true>(k, xy, T, D);
  return row_dh_finish<true>(st, row_dh_loop<ENDO>(T, D), out);
}
// fq_mul_base (CHECK_NEUTRAL = false): encode([k]G);  fq_dh_base (true): encode([k][392]G) with the neutral check
// tab: the 64 quads of the base point's table, [entry][quad], in shared memory on the device
template <bool CHECK_NEUTRAL, bool ENDO, bool STRICT = FQ_STRICT_DEFAULT> FQ_FN u32 row_fixed_base(const u32* k, uint4* tab, u32* out) {
  fp2 ox, oy;
  TabView T; T.base = tab; T.stride = 1;
  if (ENDO) { SelectBroadcast<STRICT> sel; sel.T = T; pt_to_affine(mul_endo(row_load_scalar(k), sel), ox, oy); }
  else mul_fixed_base<STRICT>(row_load_scalar(k), T, ox, oy);
  u32 st = FQ_ST_OK;
  if (CHECK_NEUTRAL && (fp2_eq_canon(ox, fp2_zero()) & fp2_eq_canon(oy, fp2_one()))) st = FQ_ST_NEUTRAL;
  if (st == FQ_ST_OK) pt_encode(ox, oy, out); else row_zero(out, 8);
  return st;
}


// fq_mul_base_comb (CHECK_NEUTRAL = false): encode([k]G);  fq_dh_base_comb (true): encode([k][392]G), neutral rejected
template <bool CHECK_NEUTRAL> FQ_FN u32 row_comb(const u32* k, const u32* tab, u32* out) {
  scal s;
  FQ_UNROLL
  for (int i = 0; i < 8; i++) s.v[i] = k[i];
  fp2 ox, oy;
  pt_to_affine(mul_comb(s, tab), ox, oy);
  u32 st = FQ_ST_OK;
  if (CHECK_NEUTRAL && (fp2_eq_canon(ox, fp2_zero()) & fp2_eq_canon(oy, fp2_one()))) st = FQ_ST_NEUTRAL;   // curve4q.py:459
  if (st == FQ_ST_OK) pt_encode(ox, oy, out);
  else { FQ_UNROLL for (int i = 0; i < 8; i++) out[i] = 0; }
  return st;
}


template <bool AFFINE, bool ENDO> __global__ void __launch_bounds__(FQ_DH_THREADS, 2)
k_dh(const void* __restrict__ k, const void* __restrict__ pt, void* __restrict__ out, unsigned char* __restrict__ status, size_t n) {
  extern __shared__ uint4 smem[];
  TabView T; T.base = smem + threadIdx.x; T.stride = FQ_DH_THREADS;
  const size_t ntiles = (n + FQ_DH_THREADS - 1) / FQ_DH_THREADS;
  for (size_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    size_t row = tile * FQ_DH_THREADS + threadIdx.x;
    const bool live = row < n;
    if (!live) row = n - 1;                          // tail threads recompute the last row and store nothing
    DhState D;
    u32 st;
    {
      u32 wk[8], wp[AFFINE ? 16 : 8];
      ld8(k, row, wk);
      if (AFFINE) { ld8(pt, 2 * row, wp); ld8(pt, 2 * row + 1, wp + 8); } else ld8(pt, row, wp);
      st = row_dh_setup<ENDO, AFFINE>(wk, wp, T, D);
    }
    ptR1 R = row_dh_loop<ENDO>(T, D);
    u32 wo[AFFINE ? 16 : 8];
    st = row_dh_finish<AFFINE>(st, R, wo);
    if (live) {
      status[row] = (unsigned char)st;
      if (AFFINE) { st8(out, 2 * row, wo); st8(out, 2 * row + 1, wo + 8); } else st8(out, row, wo);
    }
  }
}

