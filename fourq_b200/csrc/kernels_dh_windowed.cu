// kernels_dh_windowed.cu -- the DH kernels with MUL_windowed (curve4q.py:188-235); see kernels_dh.cuh
#include "kernels_dh.cuh"
cudaError_t fqk_dh_windowed_init() { return dh_init<false>(); }
cudaError_t fqk_dh_windowed(int affine, int strict, const void* k, const void* pt, void* out, void* status, size_t n, void* scratch, cudaStream_t s, cudaEvent_t* ev) {
  return dh_launch<false>(affine, strict, k, pt, out, status, n, scratch, s, ev);
}
