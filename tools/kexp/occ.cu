// occupancy experiment: DBL-only / ADD-only loops at forced CTAs per SM
#include <cstdio>
#include <cuda_runtime.h>
#include "../../fourq_b200/csrc/rows.cuh"
template <int MINB, int KIND> __global__ void __launch_bounds__(128, MINB) k_occ(const uint4* in, uint4* out, int iters) {
  extern __shared__ uint4 smem[];
  size_t row = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint4 a = in[2 * row], b = in[2 * row + 1];
  ptR1 Q;
  Q.X = fp2_set(fp_set(a.x, a.y, a.z, a.w & 0x7fffffffu), fp_set(b.x, b.y, b.z, b.w & 0x7fffffffu));
  Q.Y = fp2_set(fp_set(a.y, a.z, a.x, a.w & 0x7fffffffu), fp_set(b.z, b.y, b.x, b.w & 0x7fffffffu));
  Q.Z = fp2_set(fp_set(b.x, a.y, b.z, a.w & 0x7fffffffu), fp_set(a.x, b.y, a.z, b.w & 0x7fffffffu));
  Q.Ta = Q.X; Q.Tb = Q.Y;
  ptR2 S; S.N = Q.Y; S.D = Q.Z; S.E = Q.X; S.F = Q.Y;
  if (threadIdx.x == 999) smem[0] = a;
#pragma unroll 1
  for (int i = 0; i < iters; i++) {
    if (KIND == 0) pt_dbl(Q);
    else if (KIND == 1) { Q = pt_add(Q, S); S.F = Q.Ta; }
    else { pt_dbl(Q); Q = pt_add(Q, S); S.F = Q.Ta; }
  }
  out[2 * row] = make_uint4(Q.X.re.v[0] ^ Q.Y.re.v[0], Q.X.re.v[1] ^ Q.Z.re.v[1], Q.X.re.v[2] ^ Q.Ta.re.v[0], Q.X.re.v[3] ^ Q.Tb.re.v[1]);
  out[2 * row + 1] = make_uint4(Q.X.im.v[0] ^ Q.Y.im.v[0], Q.X.im.v[1] ^ Q.Z.im.v[1], Q.X.im.v[2] ^ Q.Ta.im.v[0], Q.X.im.v[3] ^ Q.Tb.im.v[1]);
}
template <int MINB, int KIND> void run(const char* name, const uint4* in, uint4* out, int ctas_per_sm, double wides_per_iter) {
  int iters = 256;
  int smem = (227 * 1024 / ctas_per_sm - 1024) & ~1023;
  cudaFuncSetAttribute(k_occ<MINB, KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  int nb = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_occ<MINB, KIND>, 128, smem);
  cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, k_occ<MINB, KIND>);
  size_t n = (size_t)148 * 128 * 12 * 4;   // divisible by 2,3,4,6 CTAs per SM
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e9f;
  for (int it = 0; it < 3; it++) {
    cudaEventRecord(e0);
    k_occ<MINB, KIND><<<(unsigned)(n / 128), 128, smem>>>(in, out, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
  }
  cudaError_t e = cudaGetLastError();
  printf("%-10s minb=%d want=%d got ctas/SM=%d regs=%3d local=%4zu  %.3f ms  %.3f T wide/s (%.1f%% of 9.27)  %s\n", name, MINB, ctas_per_sm, nb, fa.numRegs, fa.localSizeBytes, best,
         wides_per_iter * iters * n / best / 1e9, wides_per_iter * iters * n / best / 1e9 / 9.27 * 100, e == cudaSuccess ? "" : cudaGetErrorString(e));
}
int main() {
  size_t n = (size_t)148 * 128 * 48;
  uint4 *in, *out; cudaMalloc(&in, n * 32); cudaMalloc(&out, n * 32); cudaMemset(in, 0x5a, n * 32);
  run<1, 0>("dbl", in, out, 1, 272); run<2, 0>("dbl", in, out, 2, 272); run<3, 0>("dbl", in, out, 3, 272); run<4, 0>("dbl", in, out, 4, 272); run<4, 0>("dbl", in, out, 3, 272); run<4,0>("dbl", in, out, 2, 272);
  run<1, 1>("add", in, out, 1, 384); run<2, 1>("add", in, out, 2, 384); run<3, 1>("add", in, out, 3, 384); run<4, 1>("add", in, out, 4, 384); run<4, 1>("add", in, out, 2, 384);
  run<2, 2>("dbl+add", in, out, 2, 656); run<3, 2>("dbl+add", in, out, 3, 656); run<4, 2>("dbl+add", in, out, 4, 656);
  return 0;
}
