// kio.cuh -- 128-bit row loads/stores shared by the kernels: a thread owns one row; a warp touches 1 KiB contiguous.
#pragma once
#include "rows.cuh"

__device__ __forceinline__ void ld8(const void* base, size_t row, u32* w) {
  const uint4* p = reinterpret_cast<const uint4*>(base) + 2 * row;
  uint4 a = p[0], b = p[1];
  w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w; w[4] = b.x; w[5] = b.y; w[6] = b.z; w[7] = b.w;
}
__device__ __forceinline__ void st8(void* base, size_t row, const u32* w) {
  uint4* p = reinterpret_cast<uint4*>(base) + 2 * row;
  p[0] = make_uint4(w[0], w[1], w[2], w[3]); p[1] = make_uint4(w[4], w[5], w[6], w[7]);
}
static inline unsigned grid_for(size_t n, unsigned threads) { return (unsigned)((n + threads - 1) / threads); }
