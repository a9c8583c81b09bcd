// opcount.cu -- developer tool: SASS instruction count per primitive, measured as count(REP=3) - count(REP=2) of a
// dependent chain, so loads/stores and setup cancel.  python tools/opcount.py compiles this and prints the table.
#include "../fourq_b200/csrc/rows.cuh"

__device__ __forceinline__ fp2 ld_fp2(const u32* p) { return fp2_set(fp_set(p[0], p[1], p[2], p[3]), fp_set(p[4], p[5], p[6], p[7])); }
__device__ __forceinline__ void st_fp2(u32* p, const fp2& a) { for (int i = 0; i < 4; i++) { p[i] = a.re.v[i]; p[4 + i] = a.im.v[i]; } }

#define KERNEL(NAME, BODY)                                                                   \
  template <int REP> __global__ void NAME(u32* io) {                                         \
    u32* p = io + threadIdx.x * 64;                                                          \
    fp2 x = ld_fp2(p), y = ld_fp2(p + 8), z = ld_fp2(p + 16);                                 \
    ptR1 Q; Q.X = x; Q.Y = y; Q.Z = z; Q.Ta = ld_fp2(p + 24); Q.Tb = ld_fp2(p + 32);          \
    ptR2 S; S.N = y; S.D = z; S.E = x; S.F = ld_fp2(p + 40);                                  \
    _Pragma("unroll") for (int r = 0; r < REP; r++) { BODY; }                                \
    st_fp2(p, x); st_fp2(p + 8, y); st_fp2(p + 16, z); st_fp2(p + 24, Q.X); st_fp2(p + 32, Q.Y); st_fp2(p + 40, Q.Z); \
    st_fp2(p + 48, Q.Ta); st_fp2(p + 56, Q.Tb);                                              \
  }                                                                                          \
  template __global__ void NAME<2>(u32*);                                                    \
  template __global__ void NAME<3>(u32*);

KERNEL(op_fp2_mul, x = fp2_mul(x, y))
KERNEL(op_fp2_mul_prepped, { fp2b Y = fp2_prep(y); x = fp2_mul_prep(x, Y); z = fp2_mul_prep(z, Y); y = fp2_add(x, z); } )
KERNEL(op_fp2_sqr, x = fp2_sqr(x))
KERNEL(op_fp2_add, { x = fp2_add(x, y); y = fp2_add(y, x); })
KERNEL(op_fp2_sub, { x = fp2_sub(x, y); y = fp2_sub(y, x); })
KERNEL(op_fp_mul, x.re = fp_mul(x.re, y.re))
KERNEL(op_fp_sqr, x.re = fp_sqr(x.re))
KERNEL(op_pt_dbl, pt_dbl(Q))
KERNEL(op_pt_add, { Q = pt_add(Q, S); S.F = Q.Ta; })
KERNEL(op_cneg, { S = pt_r2_cneg(x.re.v[0], S); x.re.v[0] += S.N.re.v[1]; })
