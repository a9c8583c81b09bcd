// x25519.cu -- batched X25519 kernels (config 5 of BASELINE.json: the compare.py counterpart).
// Two kernels, like the Curve4Q DH path: k_x25519 runs the 255-step ladder of a row and leaves x2, z2 in scratch;
// k_x25519_finish computes x2 / z2 for FQ_X_FIN_ROWS rows of a thread with ONE z^(p-2) chain (Montgomery's trick).  The
// reference's z2^(p-2) maps z2 = 0 to 0 (curve25519.py:80): zeros are replaced by 1 in the shared product and their output
// forced to 0.
#include "kernels.h"
#include "x25519.cuh"

#define FQ_X_FIN_ROWS 4

static __device__ __forceinline__ void ld8x(const void* base, size_t row, u32* w) {
  const uint4* p = reinterpret_cast<const uint4*>(base) + 2 * row;
  uint4 a = p[0], b = p[1];
  w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w; w[4] = b.x; w[5] = b.y; w[6] = b.z; w[7] = b.w;
}
// scratch: x2 then z2, each as two quads per row, component-major ([4][npad] uint4)
static __device__ __forceinline__ void st_f25(uint4* p, size_t npad, const f25& a) {
  p[0] = make_uint4(a.v[0], a.v[1], a.v[2], a.v[3]); p[npad] = make_uint4(a.v[4], a.v[5], a.v[6], a.v[7]);
}
static __device__ __forceinline__ f25 ld_f25(const uint4* p, size_t npad) {
  uint4 a = p[0], b = p[npad];
  f25 r; r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w; r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
  return r;
}

__global__ void __launch_bounds__(128) k_x25519(const void* __restrict__ k, const void* __restrict__ u, uint4* __restrict__ scratch, size_t npad, size_t n) {
  size_t row = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= n) return;
  u32 kw[8], uw[8];
  ld8x(k, row, kw); ld8x(u, row, uw);
  f25 x2, z2;
  x25519_ladder(kw, uw, x2, z2);
  st_f25(scratch + row, npad, x2); st_f25(scratch + 2 * npad + row, npad, z2);
}

__global__ void __launch_bounds__(128) k_x25519_finish(const uint4* __restrict__ scratch, size_t npad, void* __restrict__ out, size_t n) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  f25 pre[FQ_X_FIN_ROWS];
  u32 zero[FQ_X_FIN_ROWS];
  f25 acc = f25_small(1);
#pragma unroll
  for (int j = 0; j < FQ_X_FIN_ROWS; j++) {
    const size_t row = t + j * stride;
    f25 z = f25_small(1);
    if (row < n) z = ld_f25(scratch + 2 * npad + row, npad);
    f25 zc = f25_canon(z);
    u32 nz = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) nz |= zc.v[i];
    zero[j] = nz == 0 ? 0xffffffffu : 0u;
#pragma unroll
    for (int i = 0; i < 8; i++) z.v[i] = (z.v[i] & ~zero[j]) | ((i == 0 ? 1u : 0u) & zero[j]);
    acc = (j == 0) ? z : f25_mul(acc, z);
    pre[j] = acc;
  }
  f25 inv = f25_inv(acc);
#pragma unroll
  for (int j = FQ_X_FIN_ROWS - 1; j >= 0; j--) {
    const size_t row = t + j * stride;
    f25 z = f25_small(1), x = f25_small(0);
    if (row < n) { z = ld_f25(scratch + 2 * npad + row, npad); x = ld_f25(scratch + row, npad); }
#pragma unroll
    for (int i = 0; i < 8; i++) z.v[i] = (z.v[i] & ~zero[j]) | ((i == 0 ? 1u : 0u) & zero[j]);
    f25 zi = (j == 0) ? inv : f25_mul(inv, pre[j > 0 ? j - 1 : 0]);
    if (j > 0) inv = f25_mul(inv, z);
    f25 r = f25_canon(f25_mul(x, zi));                                        // curve25519.py:80, 35-39
    if (row < n) {
      uint4* op = reinterpret_cast<uint4*>(out) + 2 * row;
      op[0] = make_uint4(r.v[0] & ~zero[j], r.v[1] & ~zero[j], r.v[2] & ~zero[j], r.v[3] & ~zero[j]);
      op[1] = make_uint4(r.v[4] & ~zero[j], r.v[5] & ~zero[j], r.v[6] & ~zero[j], r.v[7] & ~zero[j]);
    }
  }
}

size_t fqk_x25519_scratch_bytes(size_t n) { return (n + 127) / 128 * 128 * 64; }

cudaError_t fqk_x25519(const void* k, const void* u, void* out, size_t n, void* scratch, cudaStream_t s) {
  if (n == 0) return cudaSuccess;
  const size_t npad = (n + 127) / 128 * 128;
  k_x25519<<<(unsigned)(npad / 128), 128, 0, s>>>(k, u, (uint4*)scratch, npad, n);
  const size_t groups = (n + FQ_X_FIN_ROWS - 1) / FQ_X_FIN_ROWS;
  k_x25519_finish<<<(unsigned)((groups + 127) / 128), 128, 0, s>>>((const uint4*)scratch, npad, out, n);
  return cudaGetLastError();
}
