// scalar.cuh -- scalar reduction mod N and the signed fixed-window recoding of MUL_windowed (curve4q.py:216-226).
//
// The reference computes r = m mod N, adds N if r is even, then 63 digits d_i = (r mod 32) - 16, r = (r - d_i)/16.
// Because r stays odd, (r - d_i)/16 = (r >> 4) | 1: there is no borrow propagation and digit i is a pure bit field,
//     sign_i = bit (4i+4) of r   (1 = positive),     |d_i| = 2*ind_i + 1,  ind_i = bits (4i+1..4i+3) ^ (sign_i ? 0 : 7),
// and d_62 = 1 always (r < 2N < 2^247).  The digits are kept packed in a 256-bit shift register (r << 7) whose top
// nibble is the next digit: no secret-dependent branch or address anywhere.
#pragma once
#include "arith.cuh"

struct scal { u32 v[8]; };

FQ_FN u32 curve_n(int i) {     // curve4q.py:12, little-endian limbs
  const u32 n[8] = {0xc7768ce7u, 0x2fb2540eu, 0xfe0f7999u, 0xdfbd004du, 0x9cbc14e5u, 0xf0539782u, 0x4e5e0a72u, 0x0029cbc1u};
  return n[i];
}

// k mod N, then +N if even  ->  odd r < 2N < 2^247
FQ_FN scal scal_reduce_odd(const scal& k) {
  // q = floor((k >> 224) * floor(2^277 / N) / 2^53) is floor(k/N) or one less (k < 2^256, N > 2^245)
  u32 q = (u32)(((u64)k.v[7] * 0xc4000000ull) >> 53);
  scal r;
  u64 carry = 0;
  u32 qn[8];
  FQ_UNROLL
  for (int i = 0; i < 8; i++) { u64 t = (u64)q * curve_n(i) + carry; qn[i] = (u32)t; carry = t >> 32; }
  r.v[0] = sub_cc(k.v[0], qn[0]);
  FQ_UNROLL
  for (int i = 1; i < 7; i++) r.v[i] = subc_cc(k.v[i], qn[i]);
  r.v[7] = subc(k.v[7], qn[7]);
  // r in [0, 2N): subtract N if r >= N
  scal d;
  d.v[0] = sub_cc(r.v[0], curve_n(0));
  FQ_UNROLL
  for (int i = 1; i < 8; i++) d.v[i] = subc_cc(r.v[i], curve_n(i));
  u32 borrow = subc(0, 0);                 // 0xffffffff if r < N
  FQ_UNROLL
  for (int i = 0; i < 8; i++) r.v[i] = (r.v[i] & borrow) | (d.v[i] & ~borrow);
  // curve4q.py:218-219
  u32 even = (r.v[0] & 1) - 1;             // 0xffffffff if even
  r.v[0] = add_cc(r.v[0], curve_n(0) & even);
  FQ_UNROLL
  for (int i = 1; i < 7; i++) r.v[i] = addc_cc(r.v[i], curve_n(i) & even);
  r.v[7] = addc(r.v[7], curve_n(7) & even);
  return r;
}

// digit register: S = r << 7, so that digit 61 (bits 245..248 of r) is the top nibble
FQ_FN scal scal_digits_init(const scal& r) {
  scal s;
  s.v[0] = r.v[0] << 7;
  FQ_UNROLL
  for (int i = 1; i < 8; i++) s.v[i] = shl_pair(r.v[i - 1], r.v[i], 7);
  return s;
}
// pops the next digit (from i = 61 downwards): idx in 0..7 (table entry [2 idx + 1]P), neg = all ones if negative
FQ_FN void scal_next_digit(scal& s, u32& idx, u32& neg) {
  u32 nib = s.v[7] >> 28;
  neg = (nib >> 3) - 1;                    // sign bit 1 = positive
  idx = (nib ^ neg) & 7;
  FQ_UNROLL
  for (int i = 7; i > 0; i--) s.v[i] = shl_pair(s.v[i - 1], s.v[i], 4);
  s.v[0] <<= 4;
}
