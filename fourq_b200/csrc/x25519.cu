// x25519.cu -- batched X25519 kernel (config 5 of BASELINE.json: the compare.py counterpart).
#include "kernels.h"
#include "x25519.cuh"

__global__ void __launch_bounds__(128) k_x25519(const void* k, const void* u, void* out, size_t n) {
  size_t row = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= n) return;
  const uint4* kp = reinterpret_cast<const uint4*>(k) + 2 * row;
  const uint4* up = reinterpret_cast<const uint4*>(u) + 2 * row;
  uint4 k0 = kp[0], k1 = kp[1], u0 = up[0], u1 = up[1];
  u32 kw[8] = {k0.x, k0.y, k0.z, k0.w, k1.x, k1.y, k1.z, k1.w};
  u32 uw[8] = {u0.x, u0.y, u0.z, u0.w, u1.x, u1.y, u1.z, u1.w};
  u32 ow[8];
  row_x25519(kw, uw, ow);
  uint4* op = reinterpret_cast<uint4*>(out) + 2 * row;
  op[0] = make_uint4(ow[0], ow[1], ow[2], ow[3]); op[1] = make_uint4(ow[4], ow[5], ow[6], ow[7]);
}

cudaError_t fqk_x25519(const void* k, const void* u, void* out, size_t n, cudaStream_t s) {
  if (n == 0) return cudaSuccess;
  k_x25519<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(k, u, out, n);
  return cudaGetLastError();
}
