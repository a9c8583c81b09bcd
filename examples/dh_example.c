/* dh_example.c -- using the C ABI of libfourq_b200.so from plain C.
 *
 *   gcc -O2 -I include -o dh_example examples/dh_example.c -L fourq_b200 -lfourq_b200 -Wl,-rpath,$PWD/fourq_b200
 *
 * Alice and Bob derive a shared secret as the draft describes (draft-ladd-cfrg-4q.md:707-714): A = Compress([a]G),
 * B = Compress([b]G), K = DH(a, B) = DH(b, A); the same thing the reference does with MUL_windowed / DH_windowed
 * (impl/curve4q.py:582-584, 446-465), for a whole batch per call.  Without a CUDA device every call returns
 * FQ_ERR_NO_DEVICE: the engine has no CPU path.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "fourq_b200.h"

int main(void) {
  const size_t n = 4096;
  uint8_t *a, *b, *A, *B, *kab, *kba, *st;
  if (fq_device_count() < 1) { fprintf(stderr, "no GPU: %s\n", fq_last_error()); return 2; }
  /* pinned host memory makes the copies asynchronous; any host memory is accepted */
  if (fq_host_alloc((void**)&a, n * 32) || fq_host_alloc((void**)&b, n * 32) || fq_host_alloc((void**)&A, n * 32) ||
      fq_host_alloc((void**)&B, n * 32) || fq_host_alloc((void**)&kab, n * 32) || fq_host_alloc((void**)&kba, n * 32) ||
      fq_host_alloc((void**)&st, n)) { fprintf(stderr, "%s\n", fq_last_error()); return 1; }
  srand(1);
  for (size_t i = 0; i < n * 32; i++) { a[i] = (uint8_t)rand(); b[i] = (uint8_t)rand(); }      /* not a CSPRNG: example only */
  if (fq_mul_base_comb(a, A, n, 1) || fq_mul_base_comb(b, B, n, 1)) { fprintf(stderr, "%s\n", fq_last_error()); return 1; }
  if (fq_dh_endo(a, B, kab, st, n, 1)) { fprintf(stderr, "%s\n", fq_last_error()); return 1; }
  for (size_t i = 0; i < n; i++) if (st[i] != FQ_ST_OK) { fprintf(stderr, "row %zu failed with status %d\n", i, st[i]); return 1; }
  if (fq_dh_endo(b, A, kba, st, n, 1)) { fprintf(stderr, "%s\n", fq_last_error()); return 1; }
  if (memcmp(kab, kba, n * 32) != 0) { fprintf(stderr, "shared secrets differ\n"); return 1; }
  for (size_t i = 0; i < n; i++) kab[32 * i + 31] &= 0x7f;      /* the shared secret is the y coordinate: clear the sign bit of x */
  printf("%zu shared secrets agree; first: ", n);
  for (int i = 0; i < 32; i++) printf("%02x", kab[i]);
  printf("\nkernel time of the last call: %.3f ms\n", fq_last_kernel_ms());
  fq_host_free(a); fq_host_free(b); fq_host_free(A); fq_host_free(B); fq_host_free(kab); fq_host_free(kba); fq_host_free(st);
  return 0;
}
