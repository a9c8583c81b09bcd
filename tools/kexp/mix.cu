// mix.cu -- developer experiment: throughput of IMAD.WIDE row patterns mixed with ALU instructions of different operand
// counts, to find what bounds the limb arithmetic (pipes, issue, or register-operand bandwidth).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef uint32_t u32;

// one "row": 2 chains of {wide.cc, wide.x.cc, addc} = 4 IMAD.WIDE + 2 IADD3.X(1 read)
#define ROW(A, V0, V1, V2, V3) \
  asm volatile("mad.lo.cc.u32 %0,%10,%11,%0; madc.hi.cc.u32 %1,%10,%11,%1; madc.lo.cc.u32 %2,%10,%13,%2; madc.hi.cc.u32 %3,%10,%13,%3; addc.u32 %4,%4,0;" \
               "mad.lo.cc.u32 %5,%10,%12,%5; madc.hi.cc.u32 %6,%10,%12,%6; madc.lo.cc.u32 %7,%10,%14,%7; madc.hi.cc.u32 %8,%10,%14,%8; addc.u32 %9,%9,0;" \
               : "+r"(e[0]), "+r"(e[1]), "+r"(e[2]), "+r"(e[3]), "+r"(e[4]), "+r"(o[0]), "+r"(o[1]), "+r"(o[2]), "+r"(o[3]), "+r"(o[4]) \
               : "r"(A), "r"(V0), "r"(V1), "r"(V2), "r"(V3))

// K extra ALU instructions per row of kind KIND: 0 = IADD3.X 2-read carry chain (add.cc/addc.cc), 1 = LOP3 3-read, 2 = SHF 2-read, 3 = IADD3 1-read (x+imm)
template <int KIND, int K> __device__ __forceinline__ void extra(u32 (&x)[8]) {
  if (KIND == 0) {
#pragma unroll
    for (int i = 0; i < K; i += 4) asm volatile("add.cc.u32 %0,%0,%4; addc.cc.u32 %1,%1,%5; addc.cc.u32 %2,%2,%6; addc.u32 %3,%3,%7;" : "+r"(x[0]), "+r"(x[1]), "+r"(x[2]), "+r"(x[3]) : "r"(x[4]), "r"(x[5]), "r"(x[6]), "r"(x[7]));
  } else if (KIND == 1) {
#pragma unroll
    for (int i = 0; i < K; i++) asm volatile("lop3.b32 %0,%0,%1,%2,0xE8;" : "+r"(x[i & 3]) : "r"(x[4 + (i & 3)]), "r"(x[(i + 1) & 3]));
  } else if (KIND == 2) {
#pragma unroll
    for (int i = 0; i < K; i++) asm volatile("shf.l.wrap.b32 %0,%0,%1,7;" : "+r"(x[i & 3]) : "r"(x[4 + (i & 3)]));
  } else if (KIND == 4) {   // 32-bit IMAD (FMA pipe)
#pragma unroll
    for (int i = 0; i < K; i++) asm volatile("mad.lo.u32 %0,%1,%2,%0;" : "+r"(x[i & 3]) : "r"(x[4 + (i & 3)]), "r"(x[(i + 1) & 3]));
  } else if (KIND == 5) {   // half IMAD (FMA pipe), half LOP3 (ALU pipe), alternating
#pragma unroll
    for (int i = 0; i < K; i += 2) {
      asm volatile("mad.lo.u32 %0,%1,%2,%0;" : "+r"(x[i & 3]) : "r"(x[4 + (i & 3)]), "r"(x[(i + 1) & 3]));
      asm volatile("lop3.b32 %0,%0,%1,%2,0xE8;" : "+r"(x[(i + 2) & 3]) : "r"(x[4 + ((i + 2) & 3)]), "r"(x[(i + 3) & 3]));
    }
  } else if (KIND == 3) {
#pragma unroll
    for (int i = 0; i < K; i += 4) asm volatile("add.cc.u32 %0,%0,7; addc.cc.u32 %1,%1,0; addc.cc.u32 %2,%2,0; addc.u32 %3,%3,0;" : "+r"(x[0]), "+r"(x[1]), "+r"(x[2]), "+r"(x[3]));
  }
}

template <int KIND, int K, bool ROWS = true> __global__ void __launch_bounds__(128) kern(u32* out, u32 s0, u32 s1, int trips) {
  u32 a[4], v[8], e[5], o[5], x[8];
  for (int i = 0; i < 4; i++) a[i] = threadIdx.x * 3 + i + s0;
  for (int i = 0; i < 8; i++) { v[i] = blockIdx.x + i * 5 + s1; x[i] = threadIdx.x + i; }
  for (int i = 0; i < 5; i++) { e[i] = i; o[i] = i + 9; }
  for (int t = 0; t < trips; t++) {
#pragma unroll
    for (int u = 0; u < 4; u++) {
      if (ROWS) ROW(a[0], v[0], v[1], v[2], v[3]); extra<KIND, K>(x);
      if (ROWS) ROW(a[1], v[4], v[0], v[1], v[2]); extra<KIND, K>(x);
      if (ROWS) ROW(a[2], v[5], v[6], v[0], v[1]); extra<KIND, K>(x);
      if (ROWS) ROW(a[3], v[7], v[5], v[6], v[0]); extra<KIND, K>(x);
      a[u] ^= e[0]; // loop-carried
    }
  }
  u32 s = 0;
  for (int i = 0; i < 5; i++) s ^= e[i] ^ o[i];
  for (int i = 0; i < 8; i++) s ^= x[i];
  if (s == 0x12345678u) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int KIND, int K, bool ROWS = true> void run(const char* name, u32* d, int sms, int warps_per_sm) {
  int trips = 2048;
  dim3 grid(sms * warps_per_sm / 4), block(128);
  kern<KIND, K, ROWS><<<grid, block>>>(d, 1, 2, 16);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e9f;
  for (int r = 0; r < 3; r++) {
    cudaEventRecord(e0); kern<KIND, K, ROWS><<<grid, block>>>(d, 1, 2, trips); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
  }
  double rows = (double)grid.x * 128 * trips * 16;      // thread-rows
  double cyc_per_row_per_smsp = best * 1e-3 * 1.965e9 / (rows / 32 / (sms * 4));
  printf("%-34s warps/SM=%2d  %.3f ms  cycles per row (4 wide + 2 addc + %d extra) per SMSP: %.2f   wide/clk/SM %.1f\n", name, warps_per_sm, best, K, cyc_per_row_per_smsp,
         rows * 4 / (best * 1e-3 * 1.965e9) / sms);
}

int main() {
  u32* d; cudaMalloc(&d, 1 << 24);
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0); int sms = p.multiProcessorCount;
  for (int w = 8; w <= 32; w *= 4) {
    run<4, 4>("+4 IMAD32", d, sms, w);
    run<4, 8>("+8 IMAD32", d, sms, w);
    run<5, 4>("+2 IMAD32 +2 LOP3", d, sms, w);
    run<5, 8>("+4 IMAD32 +4 LOP3", d, sms, w);
    run<5, 16>("+8 IMAD32 +8 LOP3", d, sms, w);
    run<1, 8, false>("NO ROWS: 8 LOP3", d, sms, w);
    run<4, 8, false>("NO ROWS: 8 IMAD32", d, sms, w);
    run<5, 8, false>("NO ROWS: 4 IMAD32 + 4 LOP3", d, sms, w);
    run<0, 8, false>("NO ROWS: 8 IADD3.X 2-read", d, sms, w);
    run<0, 0>("row only", d, sms, w);
    run<3, 4>("+4 IADD3.X 1-read", d, sms, w);
    run<3, 8>("+8 IADD3.X 1-read", d, sms, w);
    run<0, 4>("+4 IADD3.X 2-read", d, sms, w);
    run<0, 8>("+8 IADD3.X 2-read", d, sms, w);
    run<0, 12>("+12 IADD3.X 2-read", d, sms, w);
    run<1, 4>("+4 LOP3 3-read", d, sms, w);
    run<1, 8>("+8 LOP3 3-read", d, sms, w);
    run<2, 4>("+4 SHF 2-read", d, sms, w);
    run<2, 8>("+8 SHF 2-read", d, sms, w);
  }
  return 0;
}
