// kio.cuh -- row loads/stores shared by the kernels.  A thread owns one 32-byte row and moves it with ONE 256-bit access
// (LDG.E.256 / STG.E.256, sm_100+): a warp request covers 1 KiB of contiguous memory and every 32-byte sector it touches is used
// in full by that one request.  64-byte affine rows are two such accesses.  Rows must be 32-byte aligned (cudaMalloc'd buffers
// and the engine's staging buffers are).
#pragma once
#include "rows.cuh"

__device__ __forceinline__ void ld8(const void* base, size_t row, u32* w) {
  const char* p = reinterpret_cast<const char*>(base) + 32 * row;
  asm volatile("ld.global.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7]) : "l"(p));
}
__device__ __forceinline__ void st8(void* base, size_t row, const u32* w) {
  char* p = reinterpret_cast<char*>(base) + 32 * row;
  asm volatile("st.global.v8.u32 [%8], {%0,%1,%2,%3,%4,%5,%6,%7};"
               :: "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7]), "l"(p) : "memory");
}
static inline unsigned grid_for(size_t n, unsigned threads) { return (unsigned)((n + threads - 1) / threads); }
