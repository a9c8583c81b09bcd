// mainloop_affine.cu -- developer experiment (not part of the product): what the endomorphism main loop would cost if the per-row
// table were normalised to affine entries (x+y, y-x, 2dxy) of 96 bytes -- 64 x (DBL + strict select over 8 x 6 quads + mixed ADD with
// 7 multiplications) -- at 2 CTAs x 128 threads and at 5 CTAs x 64 threads per SM (the 672 B of shared memory per row that 7 affine
// entries need allow 320 rows per SM).  Results are not meaningful, only the timing is; the normalisation itself (a shared inversion
// over the 8 x R table points of a thread's rows) is NOT included.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -DFQ_STRICT_SELECT -o mainloop_affine mainloop_affine.cu && ./mainloop_affine
#include <cstdio>
#include <cuda_runtime.h>
#include "../../fourq_b200/csrc/endo.cuh"
#include "../../fourq_b200/csrc/comb.cuh"

__device__ __forceinline__ void ld8(const void* base, size_t row, u32* w) {
  const uint4* p = reinterpret_cast<const uint4*>(base) + 2 * row;
  uint4 a = p[0], b = p[1];
  w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w; w[4] = b.x; w[5] = b.y; w[6] = b.z; w[7] = b.w;
}
// per-thread table of 7 entries x 6 quads in shared memory, [entry][quad][thread]; entry 7 in registers
template <int THREADS> struct SelAffine {
  uint4* base; ptA3 T7;
  __device__ __forceinline__ ptA3 operator()(u32 idx) const {
    fp w[6] = {T7.N.re, T7.N.im, T7.D.re, T7.D.im, T7.F.re, T7.F.im};
#pragma unroll
    for (int e = 0; e < 7; e++) {
      const bool c = idx == (u32)e;
#pragma unroll
      for (int q = 0; q < 6; q++) quad_take<true>(base + (e * 6 + q) * THREADS, w[q], c);
    }
    ptA3 P; P.N = fp2_set(w[0], w[1]); P.D = fp2_set(w[2], w[3]); P.F = fp2_set(w[4], w[5]);
    return P;
  }
};
FQ_FN ptA3 a3_cneg(u32 m, const ptA3& P) {
  ptA3 R;
  R.N = fp2_select(m, P.D, P.N); R.D = fp2_select(m, P.N, P.D);
  R.F.re = fp_set(P.F.re.v[0] ^ m, P.F.re.v[1] ^ m, P.F.re.v[2] ^ m, P.F.re.v[3] ^ (m & FQ_P3));
  R.F.im = fp_set(P.F.im.v[0] ^ m, P.F.im.v[1] ^ m, P.F.im.v[2] ^ m, P.F.im.v[3] ^ (m & FQ_P3));
  return R;
}
template <int THREADS, int MINB> __global__ void __launch_bounds__(THREADS, MINB) k_main(const void* k, void* out, size_t n) {
  extern __shared__ uint4 smem[];
  size_t row = (size_t)blockIdx.x * THREADS + threadIdx.x;
  u32 wk[8]; ld8(k, row, wk);
  scal S;
  for (int i = 0; i < 8; i++) S.v[i] = wk[i];
  SelAffine<THREADS> sel; sel.base = smem + threadIdx.x;
  for (int e = 0; e < 7; e++) for (int q = 0; q < 6; q++) sel.base[(e * 6 + q) * THREADS] = make_uint4(wk[q], wk[(q + e) & 7], wk[(q + 3) & 7], wk[e & 7] & 0x7fffffffu);
  sel.T7.N = fp2_set(fp_set(wk[0], wk[1], wk[2], wk[3] & 0x7fffffffu), fp_set(wk[4], wk[5], wk[6], wk[7] & 0x7fffffffu));
  sel.T7.D = sel.T7.N; sel.T7.F = sel.T7.N;
  ptA3 first = sel(wk[0] & 7);
  ptR1 Q = pt_from_affine(first.N, first.D);
#pragma unroll 1
  for (int i = 63; i >= 0; i--) {
    pt_dbl(Q);
    u32 idx, neg;
    endo_next_digit(S, idx, neg);
    Q = pt_madd(Q, a3_cneg(neg, sel(idx)));
  }
  u32 wo[8];
  for (int i = 0; i < 4; i++) { wo[i] = Q.X.re.v[i] ^ Q.Y.re.v[i] ^ Q.Z.re.v[i]; wo[4 + i] = Q.X.im.v[i] ^ Q.Y.im.v[i] ^ Q.Z.im.v[i]; }
  uint4* p = reinterpret_cast<uint4*>(out) + 2 * row;
  p[0] = make_uint4(wo[0], wo[1], wo[2], wo[3]); p[1] = make_uint4(wo[4], wo[5], wo[6], wo[7]);
}
template <int THREADS, int MINB> void run(const char* name, const void* k, void* out, size_t n) {
  int smem = 7 * 6 * 16 * THREADS;
  cudaFuncSetAttribute(k_main<THREADS, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  int nb = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_main<THREADS, MINB>, THREADS, smem);
  cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, k_main<THREADS, MINB>);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e9f;
  for (int it = 0; it < 4; it++) {
    cudaEventRecord(e0);
    k_main<THREADS, MINB><<<(unsigned)(n / THREADS), THREADS, smem>>>(k, out, n);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
  }
  cudaError_t e = cudaGetLastError();
  printf("%-28s regs=%3d local=%4zu B ctas/SM=%d smem=%6d  %.3f ms  %.2f Mrows/s  %s\n", name, fa.numRegs, fa.localSizeBytes, nb, smem, best, n / best / 1e3,
         e == cudaSuccess ? "" : cudaGetErrorString(e));
}
int main() {
  size_t n = 1 << 20;
  void *k, *out; cudaMalloc(&k, n * 32); cudaMalloc(&out, n * 32);
  cudaMemset(k, 0x5a, n * 32);
  run<128, 2>("affine 128 thr x 2 CTAs", k, out, n);
  run<64, 5>("affine 64 thr x 5 CTAs", k, out, n);
  run<64, 4>("affine 64 thr x 4 CTAs", k, out, n);
  run<96, 3>("affine 96 thr x 3 CTAs", k, out, n);
  return 0;
}
