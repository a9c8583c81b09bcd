// kernels.h -- launch wrappers of kernels.cu (device pointers, one stream).  Internal to the library.
#pragma once
#include <cstddef>
#ifdef FQ_MOCK_CUDA
#include "mock_cuda_runtime.h"
#else
#include <cuda_runtime.h>
#endif

enum { FQK_MUL = 0, FQK_SQR = 1, FQK_INV = 2, FQK_ADD = 3, FQK_SUB = 4, FQK_NEG = 5, FQK_CONJ = 6, FQK_INVSQRT = 7 };

cudaError_t fqk_device_init(cudaStream_t s);   // once per device: fixed-base tables -> __constant__, smem opt-in
cudaError_t fqk_fp2_op(int op, const void* a, const void* b, void* out, size_t n, cudaStream_t s);
cudaError_t fqk_fp_op(int op, const void* a, const void* b, void* out, size_t n, cudaStream_t s);   // GF(p), 16-byte rows, op = FQ_FP_* of the header
// GFp.select / GFp2.select (fields.py:59-64, :236-238): halves = 1 (16-byte rows) or 2 (32-byte rows), c = one condition byte per row
cudaError_t fqk_select(int halves, const void* c, const void* x, const void* y, void* out, size_t n, cudaStream_t s);
cudaError_t fqk_decode(int spec, const void* enc, void* xy, void* status, size_t n, cudaStream_t s);   // spec: draft's t == 0 branch instead of the reference's exception
cudaError_t fqk_encode(const void* xy, void* enc, size_t n, cudaStream_t s);
cudaError_t fqk_on_curve(const void* xy, void* ok, size_t n, cudaStream_t s);
// `strict` (everywhere below): table selection by the strict scan instead of masked loads (dh.cuh, fq_set_select_mode)
// variable-base DH = three kernels (prepare, ladder, finish; kernels_dh.cuh) that hand the per-row table and the projective
// result over through `scratch`, a device buffer of at least fqk_dh_scratch_bytes(n) bytes owned by the caller.
// ev: optional array of 4 events recorded before/between/after the three kernels (per-kernel timing).
size_t fqk_dh_scratch_bytes(size_t n);
cudaError_t fqk_dh(int affine, int endo, int strict, const void* k, const void* pt, void* out, void* status, size_t n, void* scratch, cudaStream_t s, cudaEvent_t* ev);
// per-algorithm translation units (kernels_dh_windowed.cu, kernels_dh_endo.cu)
cudaError_t fqk_dh_windowed_init();
cudaError_t fqk_dh_endo_init();
cudaError_t fqk_dh_windowed(int affine, int strict, const void* k, const void* pt, void* out, void* status, size_t n, void* scratch, cudaStream_t s, cudaEvent_t* ev);
cudaError_t fqk_dh_endo(int affine, int strict, const void* k, const void* pt, void* out, void* status, size_t n, void* scratch, cudaStream_t s, cudaEvent_t* ev);
cudaError_t fqk_fixed_base(int dh, int endo, int strict, const void* k, void* out, void* status, size_t n, void* scratch, cudaStream_t s);   // scratch: fqk_comb_scratch_bytes(n)
// fixed-base per-digit tables (kernels_comb.cu): tabs is the device buffer returned by fqk_comb_init
cudaError_t fqk_comb_init(void** tabs_out, cudaStream_t s);
size_t fqk_comb_scratch_bytes(size_t n);     // device scratch the caller passes to fqk_comb
cudaError_t fqk_comb(int dh, int strict, const void* tabs, const void* k, void* out, void* status, size_t n, void* scratch, int sms, cudaStream_t s);
size_t fqk_x25519_scratch_bytes(size_t n);     // device scratch the caller passes to fqk_x25519 (x2, z2 of every row)
cudaError_t fqk_x25519(const void* k, const void* u, void* out, size_t n, void* scratch, cudaStream_t s);
cudaError_t fqk_f25_op(int op, const void* a, const void* b, void* out, size_t n, cudaStream_t s);   // GF(2^255-19), 32-byte rows, op = FQ_FP_MUL.._SUB of the header
cudaError_t fqk_imad_peak(int variant, void* scratch, int blocks, int trips, cudaStream_t s);
