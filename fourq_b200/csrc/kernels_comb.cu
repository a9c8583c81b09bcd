// kernels_comb.cu -- fixed-base kernels on per-digit tables (comb.cuh): fq_mul_base_comb, fq_dh_base_comb.
// One thread = one row; the 47.25 KiB table of the base point is copied once per CTA from global to shared memory and
// then read with warp-uniform addresses (broadcast), so the per-thread state is registers only.
#include "kernels.h"
#include "kio.cuh"
#include "comb.cuh"

#define FQ_COMB_THREADS 256

// tabs: [2][FQ_COMB_WORDS]; thread t builds digit t % 63 of base t / 63
__global__ void __launch_bounds__(64) k_comb_build(u32* tabs) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= 2 * FQ_COMB_DIGITS) return;
  int which = t / FQ_COMB_DIGITS, i = t % FQ_COMB_DIGITS;
  comb_build_digit(which, i, tabs + which * FQ_COMB_WORDS + i * FQ_COMB_DIGIT_WORDS);
}

template <bool DH> __global__ void __launch_bounds__(FQ_COMB_THREADS)
k_comb(const u32* __restrict__ tabs, const void* __restrict__ k, void* __restrict__ out, unsigned char* __restrict__ status, size_t n) {
  extern __shared__ uint4 stab4[];
  const uint4* src = reinterpret_cast<const uint4*>(tabs + (DH ? FQ_COMB_WORDS : 0));
  for (int j = threadIdx.x; j < FQ_COMB_WORDS / 4; j += FQ_COMB_THREADS) stab4[j] = src[j];
  __syncthreads();
  const u32* stab = reinterpret_cast<const u32*>(stab4);
  const size_t ntiles = (n + FQ_COMB_THREADS - 1) / FQ_COMB_THREADS;
  for (size_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    size_t row = tile * FQ_COMB_THREADS + threadIdx.x;
    if (row >= n) continue;
    u32 wk[8], wo[8];
    ld8(k, row, wk);
    u32 st = row_comb<DH>(wk, stab, wo);
    if (status) status[row] = (unsigned char)st;
    st8(out, row, wo);
  }
}

cudaError_t fqk_comb_init(void** tabs_out, cudaStream_t s) {
  cudaError_t e;
  u32* tabs = nullptr;
  if ((e = cudaMalloc(&tabs, 2 * FQ_COMB_WORDS * sizeof(u32))) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(k_comb<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, FQ_COMB_WORDS * 4)) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(k_comb<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, FQ_COMB_WORDS * 4)) != cudaSuccess) return e;
  k_comb_build<<<(2 * FQ_COMB_DIGITS + 63) / 64, 64, 0, s>>>(tabs);
  if ((e = cudaGetLastError()) != cudaSuccess) { cudaFree(tabs); return e; }
  if ((e = cudaStreamSynchronize(s)) != cudaSuccess) { cudaFree(tabs); return e; }
  *tabs_out = tabs;
  return cudaSuccess;
}

cudaError_t fqk_comb(int dh, const void* tabs, const void* k, void* out, void* status, size_t n, int sms, cudaStream_t s) {
  if (n == 0) return cudaSuccess;
  // persistent CTAs: the table copy (47 KiB) is paid once per CTA, each CTA then walks over tiles of 256 rows
  unsigned tiles = grid_for(n, FQ_COMB_THREADS);
  unsigned cap = (unsigned)sms * 4u;
  unsigned g = tiles < cap ? tiles : cap;
  if (dh) k_comb<true><<<g, FQ_COMB_THREADS, FQ_COMB_WORDS * 4, s>>>((const u32*)tabs, k, out, (unsigned char*)status, n);
  else k_comb<false><<<g, FQ_COMB_THREADS, FQ_COMB_WORDS * 4, s>>>((const u32*)tabs, k, out, (unsigned char*)status, n);
  return cudaGetLastError();
}
