// Integer-pipe microbenchmark for B200 (sm_100a): issue rates of IMAD / IMAD.WIDE(.X) / IADD3 / LOP3 / SHF and of
// their mixes.  Output feeds DESIGN.md (which formulation of the limb arithmetic is cheapest) and the IMAD roofline
// denominator.  Every variant runs ILP independent dependency chains per thread, UNROLL ops per chain per loop trip.
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/pipe_microbench tools/pipe_microbench.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
typedef uint32_t u32;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1);} } while (0)

constexpr int ILP = 8;
constexpr int UNROLL = 16;

#define WIDE(LO, HI, X, B) asm volatile("mad.lo.cc.u32 %0,%2,%3,%0; madc.hi.u32 %1,%2,%3,%1;" : "+r"(LO), "+r"(HI) : "r"(X), "r"(B))
// All multiplicands are loop-carried registers so neither NVVM nor ptxas can hoist or strength-reduce them.
template <int V> __device__ __forceinline__ void body(u32 (&lo)[ILP], u32 (&hi)[ILP], u32 a, u32 b) {
#pragma unroll
  for (int u = 0; u < UNROLL; u++) {
#pragma unroll
    for (int i = 0; i < ILP; i++) {
      const int j = (i + 4) % ILP;
      if (V == 0) {        // IMAD.WIDE.U32, 64-bit accumulate, no carry: {lo,hi} += lo*b
        asm volatile("{.reg .u64 t; mov.b64 t,{%0,%1}; mad.wide.u32 t,%2,%3,t; mov.b64 {%0,%1},t;}" : "+r"(lo[i]), "+r"(hi[i]) : "r"(lo[i]), "r"(b));
      } else if (V == 1) { // IMAD.WIDE.U32 with carry-out (mad.lo.cc + madc.hi fused), multiplicand from another chain
        asm volatile("mad.lo.cc.u32 %0,%2,%3,%0; madc.hi.u32 %1,%2,%3,%1;" : "+r"(lo[i]), "+r"(hi[i]) : "r"(lo[j]), "r"(b));
      } else if (V == 2) { // IMAD (lo 32 bits)
        asm volatile("mad.lo.u32 %0,%0,%1,%2;" : "+r"(lo[i]) : "r"(b), "r"(hi[i]));
      } else if (V == 3) { // IMAD.HI
        asm volatile("mad.hi.u32 %0,%0,%1,%2;" : "+r"(lo[i]) : "r"(b), "r"(hi[i]));
      } else if (V == 4) { // IADD3 (dependent, fibonacci-like so it cannot be folded)
        asm volatile("add.u32 %0,%0,%1;" : "+r"(lo[i]) : "r"(hi[i]));
        asm volatile("add.u32 %0,%0,%1;" : "+r"(hi[i]) : "r"(lo[i]));
      } else if (V == 5) { // LOP3
        asm volatile("lop3.b32 %0,%0,%1,%2,0xE8;" : "+r"(lo[i]) : "r"(hi[i]), "r"(b));
        asm volatile("lop3.b32 %0,%0,%1,%2,0xE8;" : "+r"(hi[i]) : "r"(lo[i]), "r"(a));
      } else if (V == 6) { // SHF funnel
        asm volatile("shf.l.wrap.b32 %0,%0,%1,7;" : "+r"(lo[i]) : "r"(hi[i]));
        asm volatile("shf.l.wrap.b32 %0,%0,%1,5;" : "+r"(hi[i]) : "r"(lo[i]));
      } else if (V == 7) { // IMAD.WIDE + IADD3 1:1 (add on a different chain)
        WIDE(lo[i], hi[i], lo[(i + 1) % ILP], b);
        asm volatile("add.u32 %0,%0,%1;" : "+r"(hi[j]) : "r"(lo[j]));
      } else if (V == 8) { // IMAD.WIDE + 2 IADD3
        WIDE(lo[i], hi[i], lo[(i + 1) % ILP], b);
        asm volatile("add.u32 %0,%0,%1;" : "+r"(hi[j]) : "r"(lo[j]));
        asm volatile("add.u32 %0,%0,%1;" : "+r"(hi[(i + 2) % ILP]) : "r"(lo[(i + 6) % ILP]));
      } else if (V == 9) { // IMAD.WIDE + LOP3 1:1
        WIDE(lo[i], hi[i], lo[(i + 1) % ILP], b);
        asm volatile("lop3.b32 %0,%0,%1,%2,0xE8;" : "+r"(hi[j]) : "r"(lo[j]), "r"(a));
      } else if (V == 10) { // multiplier row shape: IMAD.WIDE(carry out) -> IMAD.WIDE.X(carry in/out) -> addc catch
        asm volatile("mad.lo.cc.u32 %0,%4,%5,%0; madc.hi.cc.u32 %1,%4,%5,%1; madc.lo.cc.u32 %2,%4,%6,%2; madc.hi.cc.u32 %3,%4,%6,%3; addc.u32 %7,%7,0;"
                     : "+r"(lo[i]), "+r"(hi[i]), "+r"(lo[j]), "+r"(hi[j]) : "r"(lo[(i + 2) % ILP]), "r"(b), "r"(a), "r"(hi[(i + 6) % ILP]));
      } else if (V == 11) { // IADD3 carry chain (add.cc/addc.cc x4)
        asm volatile("add.cc.u32 %0,%0,%2; addc.cc.u32 %1,%1,%3; addc.cc.u32 %2,%2,%0; addc.u32 %3,%3,%1;" : "+r"(lo[i]), "+r"(hi[i]), "+r"(lo[j]), "+r"(hi[j]));
      } else if (V == 12) { // FFMA for reference (fp32 pipe rate)
        float f = __uint_as_float(lo[i]);
        asm volatile("fma.rn.f32 %0,%0,%1,%2;" : "+f"(f) : "f"(__uint_as_float(b)), "f"(__uint_as_float(hi[i])));
        lo[i] = __float_as_uint(f);
      } else if (V == 13) { // IMAD + IMAD.HI pair (the non-wide way to get a 64-bit product)
        u32 x = lo[i];
        asm volatile("mad.lo.u32 %0,%1,%2,%0;" : "+r"(lo[i]) : "r"(x), "r"(b));
        asm volatile("mad.hi.u32 %0,%1,%2,%0;" : "+r"(hi[i]) : "r"(x), "r"(b));
      }
    }
  }
}

template <int V> __global__ void __launch_bounds__(256) kern(u32* out, u32 a, u32 b, int trips) {
  u32 lo[ILP], hi[ILP];
#pragma unroll
  for (int i = 0; i < ILP; i++) { lo[i] = threadIdx.x + i; hi[i] = blockIdx.x + i * 3; }
  for (int t = 0; t < trips; t++) body<V>(lo, hi, a, b);
  u32 s = 0;
#pragma unroll
  for (int i = 0; i < ILP; i++) s ^= lo[i] ^ hi[i];
  if (s == 0x12345678u) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

struct Var { const char* name; double ops_per_inner; };

template <int V> double run(const Var& v, int blocks_per_sm, int trips, int sms, float* ms_out) {
  u32* d; CK(cudaMalloc(&d, 1 << 24));
  dim3 grid(sms * blocks_per_sm), block(256);
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  kern<V><<<grid, block>>>(d, 0x9e3779b9u, 0x7f4a7c15u, trips / 8);   // warm-up
  CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < 3; r++) {
    CK(cudaEventRecord(e0));
    kern<V><<<grid, block>>>(d, 0x9e3779b9u, 0x7f4a7c15u, trips);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
  }
  CK(cudaGetLastError());
  double inner = (double)grid.x * 256 * (double)trips * UNROLL * ILP;
  *ms_out = best;
  CK(cudaFree(d));
  return inner * v.ops_per_inner / (best * 1e-3);
}

int main(int argc, char** argv) {
  int dev = 0; CK(cudaSetDevice(dev));
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, dev));
  int sms = p.multiProcessorCount; int clk_khz = 0; CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, dev));
  printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_khz_max\": %d}\n", p.name, sms, clk_khz);
  int trips = argc > 1 ? atoi(argv[1]) : 4096;
  Var vars[] = {{"mad.wide acc (ptxas splits: imad.wide+2 iadd3; products)", 1}, {"imad_wide_cc(lo.cc+hi)", 1}, {"imad_lo", 1}, {"imad_hi", 1}, {"iadd3", 2}, {"lop3", 2},
                {"shf", 2}, {"imad_wide+iadd3 (instr)", 2}, {"imad_wide+2iadd3 (instr)", 3}, {"imad_wide+lop3 (instr)", 2},
                {"row: 2 imad_wide.x + addc (instr)", 3}, {"iadd3 carry chain x4 (instr)", 4}, {"ffma", 1}, {"imad_lo+imad_hi (instr)", 2}};
  for (int bps = 1; bps <= 4; bps *= 2) {
    for (int v = 0; v < 14; v++) {
      float ms = 0; double r = 0;
      switch (v) {
        case 0: r = run<0>(vars[v], bps, trips, sms, &ms); break;   case 1: r = run<1>(vars[v], bps, trips, sms, &ms); break;
        case 2: r = run<2>(vars[v], bps, trips, sms, &ms); break;   case 3: r = run<3>(vars[v], bps, trips, sms, &ms); break;
        case 4: r = run<4>(vars[v], bps, trips, sms, &ms); break;   case 5: r = run<5>(vars[v], bps, trips, sms, &ms); break;
        case 6: r = run<6>(vars[v], bps, trips, sms, &ms); break;   case 7: r = run<7>(vars[v], bps, trips, sms, &ms); break;
        case 8: r = run<8>(vars[v], bps, trips, sms, &ms); break;   case 9: r = run<9>(vars[v], bps, trips, sms, &ms); break;
        case 10: r = run<10>(vars[v], bps, trips, sms, &ms); break; case 11: r = run<11>(vars[v], bps, trips, sms, &ms); break;
        case 12: r = run<12>(vars[v], bps, trips, sms, &ms); break; case 13: r = run<13>(vars[v], bps, trips, sms, &ms); break;
      }
      // per-SM per-clock rate at the max clock (lower bound on the true per-clock rate if the clock sagged)
      double per_sm_clk = r / sms / (clk_khz * 1e3);
      printf("{\"variant\": \"%s\", \"warps_per_sm\": %d, \"ms\": %.3f, \"Tops\": %.3f, \"per_sm_per_maxclk\": %.2f}\n",
             vars[v].name, bps * 8, ms, r / 1e12, per_sm_clk);
      fflush(stdout);
    }
  }
  return 0;
}
