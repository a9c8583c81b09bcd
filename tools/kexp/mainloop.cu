// mainloop.cu -- developer experiment (not part of the product): the endomorphism main loop 64 x (DBL + select + ADD)
// alone, at different occupancies.  Table entries are read from shared memory (entry e -> slot e % NSLOT so that the
// LDS traffic is the real one even when fewer than 8 slots fit); results are not meaningful, only the timing is.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -o mainloop mainloop.cu && ./mainloop
#include <cstdio>
#include <cuda_runtime.h>
#include "../../fourq_b200/csrc/rows.cuh"

__device__ __forceinline__ void ld8(const void* base, size_t row, u32* w) {
  const uint4* p = reinterpret_cast<const uint4*>(base) + 2 * row;
  uint4 a = p[0], b = p[1];
  w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w; w[4] = b.x; w[5] = b.y; w[6] = b.z; w[7] = b.w;
}
template <int NSLOT> struct SelectSlots {
  TabView T;
  __device__ __forceinline__ ptR2 operator()(u32 idx) const {
    ptR2 S = tab_load(T, 7 % NSLOT);
#pragma unroll
    for (int e = 0; e < 7; e++) tab_take_r2(T, e % NSLOT, S, idx == (u32)e);
    return S;
  }
};
#ifndef KEXP_THREADS
#define KEXP_THREADS 128
#endif
#ifdef KEXP_MAXNREG
#define KEXP_BOUNDS(MINB) __maxnreg__(KEXP_MAXNREG)
#else
#define KEXP_BOUNDS(MINB) __launch_bounds__(KEXP_THREADS, MINB)
#endif
template <int MINB, int NSLOT> __global__ void KEXP_BOUNDS(MINB) k_main(const void* k, void* out, size_t n) {
  extern __shared__ uint4 smem[];
  size_t row = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  TabView T; T.base = smem + threadIdx.x; T.stride = blockDim.x;
  u32 wk[8]; ld8(k, row, wk);
  scal S;
  for (int i = 0; i < 8; i++) S.v[i] = wk[i];
  for (int e = 0; e < NSLOT; e++) for (int q = 0; q < 8; q++) tab_put(T, e, q, fp_set(wk[q], wk[(q + e) & 7], wk[(q + 3) & 7], wk[e & 7] & 0x7fffffffu));
  SelectSlots<NSLOT> sel; sel.T = T;
  ptR1 Q = pt_r2_to_r4(sel(wk[0] & 7));
#ifdef FQ_EXP_UNROLL2
#pragma unroll 2
#else
#pragma unroll 1
#endif
  for (int i = 63; i >= 0; i--) {
    pt_dbl(Q);
    u32 idx, neg;
    endo_next_digit(S, idx, neg);
    Q = pt_add(Q, pt_r2_cneg(neg, sel(idx)));
  }
  u32 wo[8];
  for (int i = 0; i < 4; i++) { wo[i] = Q.X.re.v[i] ^ Q.Y.re.v[i] ^ Q.Z.re.v[i]; wo[4 + i] = Q.X.im.v[i] ^ Q.Y.im.v[i] ^ Q.Z.im.v[i]; }
  uint4* p = reinterpret_cast<uint4*>(out) + 2 * row;
  p[0] = make_uint4(wo[0], wo[1], wo[2], wo[3]); p[1] = make_uint4(wo[4], wo[5], wo[6], wo[7]);
}

template <int MINB, int NSLOT> void run(const char* name, const void* k, void* out, size_t n) {
  int smem = NSLOT * 8 * 16 * KEXP_THREADS;
  cudaFuncSetAttribute(k_main<MINB, NSLOT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  int nb = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_main<MINB, NSLOT>, KEXP_THREADS, smem);
  cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, k_main<MINB, NSLOT>);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e9f;
  for (int it = 0; it < 4; it++) {
    cudaEventRecord(e0);
    k_main<MINB, NSLOT><<<(unsigned)(n / KEXP_THREADS), KEXP_THREADS, smem>>>(k, out, n);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
  }
  cudaError_t e = cudaGetLastError();
  double wides = 64.0 * (272 + 384) * n;
  printf("%-22s regs=%3d local=%4zu B ctas/SM=%d smem=%6d  %.3f ms  %.2f Mrows/s  %.2f T wide/s  %s\n", name, fa.numRegs, fa.localSizeBytes, nb, smem, best,
         n / best / 1e3, wides / best / 1e9, e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
  size_t n = 1 << 20;
  void *k, *out; cudaMalloc(&k, n * 32); cudaMalloc(&out, n * 32);
  cudaMemset(k, 0x5a, n * 32);
#ifdef KEXP_MAXNREG
  run<2, 7>("maxnreg 7slots", k, out, n);
#else
  run<2, 7>("minb2 7slots", k, out, n);
  run<2, 4>("minb2 4slots", k, out, n);
  run<3, 4>("minb3 4slots", k, out, n);
#endif
  return 0;
}
