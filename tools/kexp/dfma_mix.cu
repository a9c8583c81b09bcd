// dfma_mix.cu -- developer experiment: does the FP64 pipe (DFMA) run beside the integer multiplier (IMAD.WIDE)?
// Per inner iteration each thread issues W IMAD.WIDE (= W mad.lo/madc.hi pairs, two carry chains as in the limb arithmetic)
// and D DFMA (independent chains).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef uint32_t u32;

template <int W, int D> __global__ void __launch_bounds__(128) kern(u32* out, u32 s0, double d0, int trips) {
  u32 a = threadIdx.x * 3 + s0, v0 = blockIdx.x + 5, v1 = blockIdx.x * 7 + 1;
  u32 e[4] = {1, 2, 3, 4}, o[4] = {5, 6, 7, 8};
  double f[8];
  for (int i = 0; i < 8; i++) f[i] = d0 + i + threadIdx.x;
  double m = d0 * 1.0000001, c = d0 * 0.5;
  for (int t = 0; t < trips; t++) {
#pragma unroll
    for (int u = 0; u < 8; u++) {
#pragma unroll
      for (int w = 0; w < W; w += 4)
        asm volatile("mad.lo.cc.u32 %0,%8,%9,%0; madc.hi.cc.u32 %1,%8,%9,%1; madc.lo.cc.u32 %2,%8,%10,%2; madc.hi.u32 %3,%8,%10,%3;"
                     "mad.lo.cc.u32 %4,%8,%10,%4; madc.hi.cc.u32 %5,%8,%10,%5; madc.lo.cc.u32 %6,%8,%9,%6; madc.hi.u32 %7,%8,%9,%7;"
                     : "+r"(e[0]), "+r"(e[1]), "+r"(e[2]), "+r"(e[3]), "+r"(o[0]), "+r"(o[1]), "+r"(o[2]), "+r"(o[3]) : "r"(a), "r"(v0), "r"(v1));
#pragma unroll
      for (int d = 0; d < D; d++) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(f[d & 7]) : "d"(m), "d"(c));
    }
    a ^= e[0];
  }
  u32 s = e[0] ^ e[1] ^ e[2] ^ e[3] ^ o[0] ^ o[1] ^ o[2] ^ o[3];
  double fs = 0; for (int i = 0; i < 8; i++) fs += f[i];
  if (s == 0x12345678u || fs == 1.2345) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int W, int D> void run(u32* d, int sms, int warps_per_sm) {
  int trips = 4096;
  dim3 grid(sms * warps_per_sm / 4), block(128);
  kern<W, D><<<grid, block>>>(d, 1, 1.5, 16);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e9f;
  for (int r = 0; r < 3; r++) {
    cudaEventRecord(e0); kern<W, D><<<grid, block>>>(d, 1, 1.5, trips); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
  }
  double thr_trips = (double)grid.x * 128 * trips * 8;                       // thread-level inner iterations
  double cyc = best * 1e-3 * 1.965e9 / (thr_trips / 32 / (sms * 4));         // cycles per inner iteration per scheduler
  printf("W=%2d wide D=%2d dfma  warps/SM=%2d  %.3f ms  cycles per iteration per SMSP %.2f  wide/clk/SM %.1f  dfma/clk/SM %.1f\n",
         W, D, warps_per_sm, best, cyc, thr_trips * W / (best * 1e-3 * 1.965e9) / sms, thr_trips * D / (best * 1e-3 * 1.965e9) / sms);
}

int main() {
  u32* d; cudaMalloc(&d, 1 << 24);
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0); int sms = p.multiProcessorCount;
  for (int w = 8; w <= 32; w *= 2) {
    run<4, 0>(d, sms, w); run<0, 8>(d, sms, w); run<0, 16>(d, sms, w);
    run<4, 2>(d, sms, w); run<4, 4>(d, sms, w); run<4, 8>(d, sms, w); run<4, 16>(d, sms, w);
  }
  return 0;
}
