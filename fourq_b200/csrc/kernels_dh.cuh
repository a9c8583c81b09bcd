// kernels_dh.cuh -- variable-base Diffie-Hellman kernel (fq_dh, fq_dh_affine, fq_dh_endo, fq_dh_endo_affine).
// Included by kernels_dh_windowed.cu and kernels_dh_endo.cu (one translation unit per algorithm: they compile in parallel).
//
// Shape.  One thread = one row.  The per-thread table (7 x 128 B, dh.cuh) lives in shared memory and the loop needs
// ~240 registers, so an SM holds 256 threads either way: two CTAs of 128 threads, 112 KiB of shared memory each.
// The kernel is bound by instruction issue (profiles/: IPC 0.46-0.5 per scheduler with the FMA-heavy pipe 55-67 % and the
// ALU pipe 50-55 % busy; more resident warps do not raise it), so the code is organised to execute as few instructions
// per row as possible.  The once-per-row setup (decode, validate, [392]P, endomorphisms, table) calls out-of-line copies of
// the field/point routines (fp2.cuh, point.cuh): fully inlined it was 36k instructions (580 KiB) per kernel and 24 % of
// all stall samples were instruction-fetch stalls; now it is 14.5k and 2.4 %.  The 62/64-iteration loop stays inlined.
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

template <bool ENDO> static cudaError_t dh_init() {
  cudaError_t e;
  if ((e = cudaFuncSetAttribute(k_dh<false, ENDO>, cudaFuncAttributeMaxDynamicSharedMemorySize, FQ_DH_SMEM)) != cudaSuccess) return e;
  return cudaFuncSetAttribute(k_dh<true, ENDO>, cudaFuncAttributeMaxDynamicSharedMemorySize, FQ_DH_SMEM);
}
template <bool ENDO> static cudaError_t dh_launch(int affine, const void* k, const void* pt, void* out, void* status, size_t n, int sms, cudaStream_t s) {
  if (n == 0) return cudaSuccess;
  unsigned tiles = grid_for(n, FQ_DH_THREADS);
  unsigned cap = (unsigned)sms * 2048u;                 // grid-stride beyond that (keeps blockIdx in range for any n)
  unsigned g = tiles < cap ? tiles : cap;
  unsigned char* st = (unsigned char*)status;
  if (affine) k_dh<true, ENDO><<<g, FQ_DH_THREADS, FQ_DH_SMEM, s>>>(k, pt, out, st, n);
  else k_dh<false, ENDO><<<g, FQ_DH_THREADS, FQ_DH_SMEM, s>>>(k, pt, out, st, n);
  return cudaGetLastError();
}
