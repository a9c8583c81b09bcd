// batchinv.cuh -- one field inversion shared by the R rows a thread owns (Montgomery's trick), and the three users of it:
// fq_fp2_inv (GFp2.inv, fields.py:194-199), the affine normalisation at the end of every scalar multiplication
// (R1toAffine, curve4q.py:103-106, with DH_core's neutral check :459 and encode :41-46) and X25519's x2 / z2
// (curve25519.py:78-80).
//
//     prefix pass   acc_j = z_0 z_1 ... z_j                     1 multiplication per row, acc_j parked in memory
//     one chain     inv   = acc_{R-1}^(p-2)                     126 S + 12 M in GF(p) (254 S + 11 M in GF(2^255-19))
//     back pass     1/z_j = inv * acc_{j-1};  inv *= z_j        2 multiplications per row
//
// The prefix products are parked in the row's own slot of the OUTPUT buffer (a prefix is exactly one output row: 32 B)
// and overwritten by the result in the back pass, which reads slot j-1 before it writes slot j -- no scratch beyond the
// buffers the caller already owns.  The thread's rows are t, t + stride, t + 2 stride, ... so that warps move
// contiguous lines.  Zeros never enter the shared product: the reference maps inv(0) to 0 (fields.py:104-106 with x = 0,
// curve25519.py:80), so a zero is replaced by one on the way in and its result forced to zero on the way out.
// Included by the kernels and by tests/hostsim (the CPU instruction-level simulation of the same code).
#pragma once
#include "rows.cuh"

#define FQ_BATCHINV_ROWS 16            // rows per thread that share one inversion

// IO concept (all methods take the thread-local row index j = 0 .. R-1; rows past the end of the batch read as one and
// store nothing):
//   elem z(int j, u32& zero)      the value to invert, zero replaced by one (zero = all ones if it was)
//   void park(int j, elem acc)    remember the prefix product z_0 .. z_j
//   elem parked(int j)
//   void emit(int j, elem zi, u32 zero)    consume 1 / z_j
template <class OPS, class IO> FQ_FN void batch_invert(IO& io, int R) {
  typedef typename OPS::elem elem;
  // Both passes are software-pipelined: the loads of the next row are issued before the multiplications of this one (the rows of
  // a thread are `stride` rows apart, so every load is a trip to L2 or HBM that would otherwise sit on the dependency chain).
  u32 zero, zero_next = 0;
  elem z_next = io.z(0, zero_next);
  elem acc = OPS::one();
  FQ_NOUNROLL
  for (int j = 0; j < R; j++) {
    elem z = z_next;
    if (j + 1 < R) z_next = io.z(j + 1, zero_next);
    acc = (j == 0) ? z : OPS::mul(acc, z);
    if (j + 1 < R) io.park(j, acc);
  }
  elem inv = OPS::inv(acc);
  z_next = io.z(R - 1, zero_next);
  elem p_next = R > 1 ? io.parked(R - 2) : OPS::one();
  FQ_NOUNROLL
  for (int j = R - 1; j >= 0; j--) {
    elem z = z_next, p = p_next;
    zero = zero_next;
    if (j > 0) { z_next = io.z(j - 1, zero_next); if (j > 1) p_next = io.parked(j - 2); }
    elem zi = inv;
    if (j > 0) { zi = OPS::mul(inv, p); inv = OPS::mul(inv, z); }
    io.emit(j, zi, zero);
  }
}

// multiplications inlined: a real call would make the prefetched loads wait at the call boundary
struct Fp2Ops {
  typedef fp2 elem;
  static FQ_MFN fp2 one() { return fp2_one(); }
  static FQ_MFN fp2 mul(const fp2& a, const fp2& b) { return fp2_mul(a, b); }
  static FQ_MFN fp2 inv(const fp2& a) { return fp2_inv(a); }
};

FQ_FN fp ldq4(const uint4* p) { uint4 w = *p; return fp_set(w.x, w.y, w.z, w.w); }
FQ_FN void stq4(uint4* p, const fp& a) { *p = make_uint4(a.v[0], a.v[1], a.v[2], a.v[3]); }
FQ_FN u32 fp2_zero_mask(const fp2& a) { return (fp_is_zero(a.re) & fp_is_zero(a.im)) ? 0xffffffffu : 0u; }

// ---------------------------------------------------------------- fq_fp2_inv: rows of 32 bytes in, 32 bytes out
struct Fp2InvIO {
  const uint4* a; uint4* out; size_t n, t, stride;
  FQ_MFN fp2 z(int j, u32& zero) const {
    const size_t row = t + (size_t)j * stride;
    fp2 x = fp2_one();
    if (row < n) x = fp2_set(fp_from_u128(ldq4(a + 2 * row)), fp_from_u128(ldq4(a + 2 * row + 1)));    // any 128-bit halves (fields.py reduces)
    zero = fp2_zero_mask(x);
    return fp2_select(zero, fp2_one(), x);
  }
  FQ_MFN void park(int j, const fp2& acc) const {
    const size_t row = t + (size_t)j * stride;
    if (row < n) { stq4(out + 2 * row, acc.re); stq4(out + 2 * row + 1, acc.im); }
  }
  FQ_MFN fp2 parked(int j) const {
    const size_t row = t + (size_t)j * stride;
    return row < n ? fp2_set(ldq4(out + 2 * row), ldq4(out + 2 * row + 1)) : fp2_one();
  }
  FQ_MFN void emit(int j, const fp2& zi, u32 zero) const {
    const size_t row = t + (size_t)j * stride;
    if (row >= n) return;
    fp2 r = fp2_canon(fp2_select(zero, fp2_zero(), zi));
    stq4(out + 2 * row, r.re); stq4(out + 2 * row + 1, r.im);
  }
};

// ---------------------------------------------------------------- R1toAffine + neutral check + encode for projective rows
// R: (X, Y, Z) of every row as six quads, component-major ([6][npad]); meta[row] >> 8 = the status so far (rows that failed
// validation carry arbitrary coordinates: their output is zero-filled).  out: 32-byte encoded rows, or 64-byte x | y rows
// (AFFINE).  CHECK_NEUTRAL = false (MUL_*: no failure path, the neutral point is encoded like any other; status may be null).
template <bool AFFINE, bool CHECK_NEUTRAL> struct FinishIO {
  const uint4* R; const u32* meta; size_t npad; uint4* out; unsigned char* status; size_t n, t, stride;
  FQ_MFN fp2 z(int j, u32& zero) const {
    const size_t row = t + (size_t)j * stride;
    fp2 v = fp2_one();
    if (row < n) v = fp2_set(ldq4(R + 4 * npad + row), ldq4(R + 5 * npad + row));
    zero = fp2_zero_mask(v);              // Z is never 0 on the curve (complete formulas); only rows that already failed can hold one
    return fp2_select(zero, fp2_one(), v);
  }
  FQ_MFN uint4* slot(size_t row) const { return out + (AFFINE ? 4 : 2) * row; }
  FQ_MFN void park(int j, const fp2& acc) const {
    const size_t row = t + (size_t)j * stride;
    if (row < n) { stq4(slot(row), acc.re); stq4(slot(row) + 1, acc.im); }
  }
  FQ_MFN fp2 parked(int j) const {
    const size_t row = t + (size_t)j * stride;
    return row < n ? fp2_set(ldq4(slot(row)), ldq4(slot(row) + 1)) : fp2_one();
  }
  FQ_MFN void emit(int j, const fp2& zi, u32) const {
    const size_t row = t + (size_t)j * stride;
    if (row >= n) return;
    const fp2 X = fp2_set(ldq4(R + row), ldq4(R + npad + row)), Y = fp2_set(ldq4(R + 2 * npad + row), ldq4(R + 3 * npad + row));
    const fp2b Zi = fp2_prep(zi);
    const fp2 ox = fp2_canon(fp2_mul_prep(X, Zi)), oy = fp2_canon(fp2_mul_prep(Y, Zi));                   // curve4q.py:103-106
    u32 st = meta[row] >> 8;
    const bool neutral = fp2_eq_canon(ox, fp2_zero()) & fp2_eq_canon(oy, fp2_one());                // curve4q.py:459
    if (CHECK_NEUTRAL && st == FQ_ST_OK && neutral) st = FQ_ST_NEUTRAL;
    u32 wo[AFFINE ? 16 : 8];
    if (AFFINE) { if (st == FQ_ST_OK) { row_store_fp2(wo, ox); row_store_fp2(wo + 8, oy); } else row_zero(wo, 16); }
    else { if (st == FQ_ST_OK) pt_encode(ox, oy, wo); else row_zero(wo, 8); }
    if (status) status[row] = (unsigned char)st;
    FQ_UNROLL
    for (int q = 0; q < (AFFINE ? 4 : 2); q++) slot(row)[q] = make_uint4(wo[4 * q], wo[4 * q + 1], wo[4 * q + 2], wo[4 * q + 3]);
  }
};
