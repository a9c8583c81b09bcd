// kernels_dh.cuh -- variable-base Diffie-Hellman kernels (fq_dh, fq_dh_affine, fq_dh_endo, fq_dh_endo_affine).
// Included by kernels_dh_windowed.cu and kernels_dh_endo.cu (one translation unit per algorithm: they compile in parallel).
//
// Shape.  One thread = one row, three kernels per batch, each sized for what bounds it (profiles/):
//   k_dh_prep    decode + validate, [392]P, (endomorphisms,) table T[0..7] in R2, scalar plan.  Straight-line code with
//                out-of-line field/point routines, no shared memory, <= 170 registers: three CTAs of 128 threads per SM.
//                The table goes to a scratch buffer in HBM laid out [entry][quad][row] (1 KiB per row, coalesced).
//   k_dh_ladder  copies the row's table into shared memory ([entry][quad][thread], 112 KiB per CTA, entry 7 in
//                registers) and runs the 64 x (DBL + ADD) or 62 x (4 DBL + ADD) loop; two CTAs per SM (shared memory
//                and ~200 registers both allow 256 threads).  Writes (X, Y, Z) to scratch.
//   k_dh_finish  each thread normalises FQ_FIN_ROWS rows with ONE inversion (Montgomery's trick: prefix products, one
//                x^(p-2) chain, back-substitution), checks for the neutral point, encodes, writes result and status.
// All three are bound by integer-instruction issue (IMAD.WIDE.U32 at half rate plus its carry handling); the split lets
// the setup run with 12 warps per SM instead of 8 and removes 3/4 of the inversions.  The scratch traffic (2.3 KiB per
// row) is ~3 % of HBM bandwidth at the measured row rate.  k_dh (below) is the same row as ONE kernel; it is kept for
// A/B experiments (tools/kexp/split.cu) and is not instantiated by the library.
#pragma once
#include "kernels.h"
#include "kio.cuh"

#define FQ_DH_THREADS 128
#define FQ_DH_SMEM (56 * 16 * FQ_DH_THREADS)       // 7 table entries x 8 quads x 16 B per thread = 112 KiB per CTA

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

// ---------------------------------------------------------------- the three-kernel pipeline

#define FQ_FIN_ROWS 4                                // rows per thread that share one inversion in k_dh_finish
#define FQ_DH_SCRATCH_QUADS (64 + 2 + 6)             // table | digit register | (X, Y, Z), 16 B each
#define FQ_DH_MAX_BATCH ((size_t)1 << 22)            // rows per launch group (keeps the u32 table indices in range)

// scratch of one batch: npad rows (a multiple of 128); every array is [component][row] so that warps move 512 B lines
struct DhScratch { uint4* tab; uint4* plan; uint4* R; u32* meta; size_t npad; };

static inline size_t dh_scratch_bytes(size_t rows) {
  size_t npad = (rows + FQ_DH_THREADS - 1) / FQ_DH_THREADS * FQ_DH_THREADS;
  return npad * (FQ_DH_SCRATCH_QUADS * 16 + 4);
}
static inline DhScratch dh_scratch_view(void* base, size_t rows) {
  DhScratch sc;
  sc.npad = (rows + FQ_DH_THREADS - 1) / FQ_DH_THREADS * FQ_DH_THREADS;
  sc.tab = reinterpret_cast<uint4*>(base); sc.plan = sc.tab + 64 * sc.npad; sc.R = sc.plan + 2 * sc.npad;
  sc.meta = reinterpret_cast<u32*>(sc.R + 6 * sc.npad);
  return sc;
}

__device__ __forceinline__ void stq(uint4* p, const fp& a) { *p = make_uint4(a.v[0], a.v[1], a.v[2], a.v[3]); }
__device__ __forceinline__ fp ldq(const uint4* p) { uint4 w = *p; return fp_set(w.x, w.y, w.z, w.w); }

// three CTAs per SM with the endomorphisms (a fourth would spill), four without (measured: 3.26 vs 3.47 ms, 2.15 vs 2.21 ms)
template <bool AFFINE, bool ENDO> __global__ void __launch_bounds__(FQ_DH_THREADS, ENDO ? 3 : 4)
k_dh_prep(const void* __restrict__ k, const void* __restrict__ pt, DhScratch sc, size_t n) {
  const size_t row = (size_t)blockIdx.x * FQ_DH_THREADS + threadIdx.x;
  const size_t src = row < n ? row : n - 1;          // tail threads recompute the last row (their scratch slots exist)
  TabView T; T.base = sc.tab + row; T.stride = (u32)sc.npad;
  u32 wk[8], wp[AFFINE ? 16 : 8];
  ld8(k, src, wk);
  if (AFFINE) { ld8(pt, 2 * src, wp); ld8(pt, 2 * src + 1, wp + 8); } else ld8(pt, src, wp);
  DhState D;
  u32 st = row_dh_setup<ENDO, AFFINE>(wk, wp, T, D);
  tab_store(T, 7, D.T7);
  sc.plan[row] = make_uint4(D.plan.S.v[0], D.plan.S.v[1], D.plan.S.v[2], D.plan.S.v[3]);
  sc.plan[sc.npad + row] = make_uint4(D.plan.S.v[4], D.plan.S.v[5], D.plan.S.v[6], D.plan.S.v[7]);
  sc.meta[row] = D.plan.first | (st << 8);
}

template <bool ENDO, bool STRICT> __global__ void __launch_bounds__(FQ_DH_THREADS, 2)
k_dh_ladder(DhScratch sc) {
  extern __shared__ uint4 smem[];
  const size_t row = (size_t)blockIdx.x * FQ_DH_THREADS + threadIdx.x;
  TabView T; T.base = smem + threadIdx.x; T.stride = FQ_DH_THREADS;
  const uint4* g = sc.tab + row;
#pragma unroll 8
  for (int i = 0; i < 56; i++) T.base[i * FQ_DH_THREADS] = g[(size_t)i * sc.npad];
  DhState D;
  { TabView G; G.base = sc.tab + row; G.stride = (u32)sc.npad; D.T7 = tab_load(G, 7); }
  const uint4 s0 = sc.plan[row], s1 = sc.plan[sc.npad + row];
  D.plan.S.v[0] = s0.x; D.plan.S.v[1] = s0.y; D.plan.S.v[2] = s0.z; D.plan.S.v[3] = s0.w;
  D.plan.S.v[4] = s1.x; D.plan.S.v[5] = s1.y; D.plan.S.v[6] = s1.z; D.plan.S.v[7] = s1.w;
  D.plan.first = sc.meta[row] & 0xffu;
  ptR1 R = row_dh_loop<ENDO, STRICT>(T, D);
  uint4* o = sc.R + row;
  stq(o, R.X.re); stq(o + sc.npad, R.X.im); stq(o + 2 * sc.npad, R.Y.re); stq(o + 3 * sc.npad, R.Y.im);
  stq(o + 4 * sc.npad, R.Z.re); stq(o + 5 * sc.npad, R.Z.im);
}

// R1toAffine (curve4q.py:103-106) for FQ_FIN_ROWS rows of a thread with one GF(p^2) inversion, the neutral check of DH_core
// (curve4q.py:459), encode (curve4q.py:41-46).  Thread t owns rows t, t + stride, ...  Z is never 0 on the curve (complete
// formulas); a zero (only possible on rows that already failed validation) is replaced by 1 so it cannot poison the
// shared product.
// CHECK_NEUTRAL = false (fq_mul_base_comb): MUL_* has no failure path, the neutral point is encoded like any other and
// status may be null.
template <bool AFFINE, bool CHECK_NEUTRAL> __global__ void __launch_bounds__(FQ_DH_THREADS)
k_dh_finish(DhScratch sc, void* __restrict__ out, unsigned char* __restrict__ status, size_t n) {
  const size_t stride = (size_t)gridDim.x * FQ_DH_THREADS;
  const size_t t = (size_t)blockIdx.x * FQ_DH_THREADS + threadIdx.x;
  fp2 pre[FQ_FIN_ROWS];
  fp2 acc = fp2_one();
  FQ_UNROLL
  for (int j = 0; j < FQ_FIN_ROWS; j++) {
    const size_t row = t + j * stride;
    fp2 z = fp2_one();
    if (row < n) { const uint4* o = sc.R + row; z = fp2_set(ldq(o + 4 * sc.npad), ldq(o + 5 * sc.npad)); }
    if (fp_is_zero(z.re) & fp_is_zero(z.im)) z = fp2_one();
    acc = (j == 0) ? z : fp2_mul_c(acc, z);
    pre[j] = acc;
  }
  fp2 inv = fp2_inv(acc);
  FQ_UNROLL
  for (int j = FQ_FIN_ROWS - 1; j >= 0; j--) {
    const size_t row = t + j * stride;
    fp2 z = fp2_one(), X = fp2_zero(), Y = fp2_one();
    if (row < n) {
      const uint4* o = sc.R + row;
      X = fp2_set(ldq(o), ldq(o + sc.npad)); Y = fp2_set(ldq(o + 2 * sc.npad), ldq(o + 3 * sc.npad));
      z = fp2_set(ldq(o + 4 * sc.npad), ldq(o + 5 * sc.npad));
    }
    if (fp_is_zero(z.re) & fp_is_zero(z.im)) z = fp2_one();
    fp2 zi = (j == 0) ? inv : fp2_mul_c(inv, pre[j > 0 ? j - 1 : 0]);
    if (j > 0) inv = fp2_mul_c(inv, z);
    fp2 ox = fp2_canon(fp2_mul_c(X, zi)), oy = fp2_canon(fp2_mul_c(Y, zi));
    if (row < n) {
      u32 st = sc.meta[row] >> 8;
      const bool neutral = fp2_eq_canon(ox, fp2_zero()) & fp2_eq_canon(oy, fp2_one());      // curve4q.py:459
      if (CHECK_NEUTRAL && st == FQ_ST_OK && neutral) st = FQ_ST_NEUTRAL;
      u32 wo[AFFINE ? 16 : 8];
      if (AFFINE) { if (st == FQ_ST_OK) { row_store_fp2(wo, ox); row_store_fp2(wo + 8, oy); } else row_zero(wo, 16); }
      else { if (st == FQ_ST_OK) pt_encode(ox, oy, wo); else row_zero(wo, 8); }
      if (status) status[row] = (unsigned char)st;
      if (AFFINE) { st8(out, 2 * row, wo); st8(out, 2 * row + 1, wo + 8); } else st8(out, row, wo);
    }
  }
}
static inline unsigned dh_finish_grid(size_t rows) {
  return (unsigned)((((rows + FQ_FIN_ROWS - 1) / FQ_FIN_ROWS) + FQ_DH_THREADS - 1) / FQ_DH_THREADS);
}
// scratch of a producer that only hands (X, Y, Z) and a status word to k_dh_finish (the fixed-base comb kernel)
static inline size_t fin_scratch_bytes(size_t rows) {
  size_t npad = (rows + 255) / 256 * 256;
  return npad * (6 * 16 + 4);
}
static inline DhScratch fin_scratch_view(void* base, size_t rows) {
  DhScratch sc;
  sc.npad = (rows + 255) / 256 * 256;
  sc.tab = nullptr; sc.plan = nullptr; sc.R = reinterpret_cast<uint4*>(base); sc.meta = reinterpret_cast<u32*>(sc.R + 6 * sc.npad);
  return sc;
}

template <bool ENDO> static cudaError_t dh_init() {
  cudaError_t e = cudaFuncSetAttribute(k_dh_ladder<ENDO, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, FQ_DH_SMEM);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(k_dh_ladder<ENDO, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, FQ_DH_SMEM);
}
// scratch: dh_scratch_bytes(min(n, FQ_DH_MAX_BATCH)) bytes on the device.  ev: optional 4 events recorded around the three
// kernels of the (last) batch, for per-kernel timing.
// strict: table selection by the strict scan instead of masked loads (dh.cuh)
template <bool ENDO> static cudaError_t dh_launch(int affine, int strict, const void* k, const void* pt, void* out, void* status, size_t n, void* scratch,
                                                  cudaStream_t s, cudaEvent_t* ev) {
  const size_t in_pt = affine ? 64 : 32, out_b = affine ? 64 : 32;
  for (size_t r0 = 0; r0 < n; r0 += FQ_DH_MAX_BATCH) {
    const size_t rows = n - r0 < FQ_DH_MAX_BATCH ? n - r0 : FQ_DH_MAX_BATCH;
    DhScratch sc = dh_scratch_view(scratch, rows);
    const unsigned g = (unsigned)(sc.npad / FQ_DH_THREADS);
    const unsigned gf = dh_finish_grid(rows);
    const char* kk = (const char*)k + 32 * r0; const char* pp = (const char*)pt + in_pt * r0;
    char* oo = (char*)out + out_b * r0; unsigned char* st = (unsigned char*)status + r0;
    if (ev) cudaEventRecord(ev[0], s);
    if (affine) k_dh_prep<true, ENDO><<<g, FQ_DH_THREADS, 0, s>>>(kk, pp, sc, rows);
    else k_dh_prep<false, ENDO><<<g, FQ_DH_THREADS, 0, s>>>(kk, pp, sc, rows);
    if (ev) cudaEventRecord(ev[1], s);
    if (strict) k_dh_ladder<ENDO, true><<<g, FQ_DH_THREADS, FQ_DH_SMEM, s>>>(sc);
    else k_dh_ladder<ENDO, false><<<g, FQ_DH_THREADS, FQ_DH_SMEM, s>>>(sc);
    if (ev) cudaEventRecord(ev[2], s);
    if (affine) k_dh_finish<true, true><<<gf, FQ_DH_THREADS, 0, s>>>(sc, oo, st, rows);
    else k_dh_finish<false, true><<<gf, FQ_DH_THREADS, 0, s>>>(sc, oo, st, rows);
    if (ev) cudaEventRecord(ev[3], s);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
  }
  return cudaSuccess;
}
