// kernels_dh_endo.cu -- the DH kernels with MUL_endo (curve4q.py:405-442); see kernels_dh.cuh
#include "kernels_dh.cuh"
cudaError_t fqk_dh_endo_init() { return dh_init<true>(); }
cudaError_t fqk_dh_endo(int affine, int strict, const void* k, const void* pt, void* out, void* status, size_t n, void* scratch, cudaStream_t s, cudaEvent_t* ev) {
  return dh_launch<true>(affine, strict, k, pt, out, status, n, scratch, s, ev);
}
