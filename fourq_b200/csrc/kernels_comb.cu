// kernels_comb.cu -- fixed-base kernels on per-digit tables (comb.cuh): fq_mul_base_comb, fq_dh_base_comb.
// One thread = one row; the 47.25 KiB table of the base point is copied once per CTA from global to shared memory and
// then read with warp-uniform addresses (broadcast), so the per-thread state is registers only.  k_comb leaves the result
// projective; k_dh_finish (kernels_dh.cuh) normalises sixteen rows per inversion and encodes.
#include "kernels_dh.cuh"
#include "comb.cuh"

#define FQ_COMB_THREADS 256

// tabs: [2][FQ_COMB_WORDS]; thread t builds digit t % 63 of base t / 63
__global__ void __launch_bounds__(64) k_comb_build(u32* tabs) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= 2 * FQ_COMB_DIGITS) return;
  int which = t / FQ_COMB_DIGITS, i = t % FQ_COMB_DIGITS;
  comb_build_digit(which, i, tabs + which * FQ_COMB_WORDS + i * FQ_COMB_DIGIT_WORDS);
}

// [k]B in R1 for every row; (X, Y, Z) go to scratch for k_dh_finish (one inversion per FQ_BATCHINV_ROWS rows, encode)
template <bool DH, bool STRICT> __global__ void __launch_bounds__(FQ_COMB_THREADS)
k_comb(const u32* __restrict__ tabs, const void* __restrict__ k, DhScratch sc, size_t n) {
  extern __shared__ uint4 stab4[];
  const uint4* src = reinterpret_cast<const uint4*>(tabs + (DH ? FQ_COMB_WORDS : 0));
  for (int j = threadIdx.x; j < FQ_COMB_WORDS / 4; j += FQ_COMB_THREADS) stab4[j] = src[j];
  __syncthreads();
  const u32* stab = reinterpret_cast<const u32*>(stab4);
  const size_t ntiles = (n + FQ_COMB_THREADS - 1) / FQ_COMB_THREADS;
  for (size_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    size_t row = tile * FQ_COMB_THREADS + threadIdx.x;
    if (row >= n) continue;
    u32 wk[8];
    ld8(k, row, wk);
    scal sk;
    FQ_UNROLL
    for (int i = 0; i < 8; i++) sk.v[i] = wk[i];
    ptR1 R = mul_comb<STRICT>(sk, stab);
    uint4* o = sc.R + row;
    stq(o, R.X.re); stq(o + sc.npad, R.X.im); stq(o + 2 * sc.npad, R.Y.re); stq(o + 3 * sc.npad, R.Y.im);
    stq(o + 4 * sc.npad, R.Z.re); stq(o + 5 * sc.npad, R.Z.im);
    sc.meta[row] = 0;
  }
}

cudaError_t fqk_comb_init(void** tabs_out, cudaStream_t s) {
  cudaError_t e;
  u32* tabs = nullptr;
  if ((e = cudaFuncSetAttribute(k_comb<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, FQ_COMB_WORDS * 4)) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(k_comb<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, FQ_COMB_WORDS * 4)) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(k_comb<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, FQ_COMB_WORDS * 4)) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(k_comb<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, FQ_COMB_WORDS * 4)) != cudaSuccess) return e;
  if ((e = cudaMalloc(&tabs, 2 * FQ_COMB_WORDS * sizeof(u32))) != cudaSuccess) return e;      // freed on every error path below
  k_comb_build<<<(2 * FQ_COMB_DIGITS + 63) / 64, 64, 0, s>>>(tabs);
  if ((e = cudaGetLastError()) != cudaSuccess) { cudaFree(tabs); return e; }
  if ((e = cudaStreamSynchronize(s)) != cudaSuccess) { cudaFree(tabs); return e; }
  *tabs_out = tabs;
  return cudaSuccess;
}

size_t fqk_comb_scratch_bytes(size_t n) { return fin_scratch_bytes(n < FQ_DH_MAX_BATCH ? n : FQ_DH_MAX_BATCH); }

cudaError_t fqk_comb(int dh, int strict, const void* tabs, const void* k, void* out, void* status, size_t n, void* scratch, int sms, cudaStream_t s) {
  for (size_t r0 = 0; r0 < n; r0 += FQ_DH_MAX_BATCH) {
    const size_t rows = n - r0 < FQ_DH_MAX_BATCH ? n - r0 : FQ_DH_MAX_BATCH;
    DhScratch sc = fin_scratch_view(scratch, rows);
    // persistent CTAs: the table copy (47 KiB) is paid once per CTA, each CTA then walks over tiles of 256 rows
    unsigned tiles = grid_for(rows, FQ_COMB_THREADS);
    // two CTAs are resident per SM (128 registers x 256 threads).  Big launches run two waves of CTAs (better balance at the
    // end); chunk-sized launches run one, so that the 47 KiB table copy of a CTA is shared by 4 tiles instead of 2
    unsigned cap = (unsigned)sms * (tiles >= (unsigned)sms * 16u ? 4u : 2u);
    unsigned g = tiles < cap ? tiles : cap;
    const char* kk = (const char*)k + 32 * r0; char* oo = (char*)out + 32 * r0;
    unsigned char* st = status ? (unsigned char*)status + r0 : nullptr;
    if (dh) {
      if (strict) k_comb<true, true><<<g, FQ_COMB_THREADS, FQ_COMB_WORDS * 4, s>>>((const u32*)tabs, kk, sc, rows);
      else k_comb<true, false><<<g, FQ_COMB_THREADS, FQ_COMB_WORDS * 4, s>>>((const u32*)tabs, kk, sc, rows);
      k_dh_finish<false, true><<<dh_finish_grid(rows), FQ_FIN_THREADS, 0, s>>>(sc, oo, st, rows);
    } else {
      if (strict) k_comb<false, true><<<g, FQ_COMB_THREADS, FQ_COMB_WORDS * 4, s>>>((const u32*)tabs, kk, sc, rows);
      else k_comb<false, false><<<g, FQ_COMB_THREADS, FQ_COMB_WORDS * 4, s>>>((const u32*)tabs, kk, sc, rows);
      k_dh_finish<false, false><<<dh_finish_grid(rows), FQ_FIN_THREADS, 0, s>>>(sc, oo, st, rows);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
  }
  return cudaSuccess;
}
