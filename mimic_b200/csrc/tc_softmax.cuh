// Mask-free softmax fast path shared by the tcgen05 forward kernels: row max and exp2/pack/store of one
// 128-column fp32 score row held in registers.  Written for instruction-level parallelism — one softmax
// warp per SM sub-partition has nobody to hide latency behind, so every reduction runs as 8 (max) or
// 4 (sum) independent chains instead of one 64-long dependent chain.
#pragma once
#include "tc_ptx.cuh"

namespace mimic {
namespace tc {

__device__ __forceinline__ float row_max128(const float (&s)[128]) {
  float m[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) m[k] = fmaxf(s[16 * k], s[16 * k + 1]);
#pragma unroll
  for (int c = 2; c < 16; c += 2)
#pragma unroll
    for (int k = 0; k < 8; ++k) m[k] = fmaxf(m[k], fmaxf(s[16 * k + c], s[16 * k + c + 1]));
  return fmaxf(fmaxf(fmaxf(m[0], m[1]), fmaxf(m[2], m[3])), fmaxf(fmaxf(m[4], m[5]), fmaxf(m[6], m[7])));
}

// P = exp2(s * scale2 + neg_m) for all 128 columns -> bf16 pairs -> TMEM columns [t_P, t_P + 64); returns sum(P)
__device__ __forceinline__ float exp_pack_store128(const float (&s)[128], float scale2, float neg_m, uint32_t t_P) {
  const float2 sc = make_float2(scale2, scale2), ng = make_float2(neg_m, neg_m);
  float2 sum[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) sum[k] = make_float2(0.f, 0.f);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 a[16];
#pragma unroll
    for (int c2 = 0; c2 < 16; ++c2) a[c2] = __ffma2_rn(make_float2(s[32 * i + 2 * c2], s[32 * i + 2 * c2 + 1]), sc, ng);
    uint32_t pk[16];
#pragma unroll
    for (int c2 = 0; c2 < 16; ++c2) {
      const float2 e = make_float2(ptx::ex2(a[c2].x), ptx::ex2(a[c2].y));
      sum[c2 & 3] = __fadd2_rn(sum[c2 & 3], e);
      pk[c2] = ptx::pack_bf16(e.x, e.y);
    }
    ptx::tmem_st16(t_P + 16 * i, pk);
  }
  const float2 t = __fadd2_rn(__fadd2_rn(sum[0], sum[1]), __fadd2_rn(sum[2], sum[3]));
  return t.x + t.y;
}

// 64-column variants (kernel with 64-key tiles, two CTAs per SM)
__device__ __forceinline__ float row_max64(const float (&s)[64]) {
  float m[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) m[k] = fmaxf(s[8 * k], s[8 * k + 1]);
#pragma unroll
  for (int c = 2; c < 8; c += 2)
#pragma unroll
    for (int k = 0; k < 8; ++k) m[k] = fmaxf(m[k], fmaxf(s[8 * k + c], s[8 * k + c + 1]));
  return fmaxf(fmaxf(fmaxf(m[0], m[1]), fmaxf(m[2], m[3])), fmaxf(fmaxf(m[4], m[5]), fmaxf(m[6], m[7])));
}
__device__ __forceinline__ float exp_pack_store64(const float (&s)[64], float scale2, float neg_m, uint32_t t_P) {
  const float2 sc = make_float2(scale2, scale2), ng = make_float2(neg_m, neg_m);
  float2 sum[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) sum[k] = make_float2(0.f, 0.f);
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    float2 a[16];
#pragma unroll
    for (int c2 = 0; c2 < 16; ++c2) a[c2] = __ffma2_rn(make_float2(s[32 * i + 2 * c2], s[32 * i + 2 * c2 + 1]), sc, ng);
    uint32_t pk[16];
#pragma unroll
    for (int c2 = 0; c2 < 16; ++c2) {
      const float2 e = make_float2(ptx::ex2(a[c2].x), ptx::ex2(a[c2].y));
      sum[c2 & 3] = __fadd2_rn(sum[c2 & 3], e);
      pk[c2] = ptx::pack_bf16(e.x, e.y);
    }
    ptx::tmem_st16(t_P + 16 * i, pk);
  }
  const float2 t = __fadd2_rn(__fadd2_rn(sum[0], sum[1]), __fadd2_rn(sum[2], sum[3]));
  return t.x + t.y;
}

// P = exp2(s * scale2 + neg_m) for 32 columns -> 16 bf16 pairs -> TMEM columns [t_P, t_P + 16); returns sum(P)
__device__ __forceinline__ float exp_pack_store32(const float (&s)[32], float scale2, float neg_m, uint32_t t_P) {
  const float2 sc = make_float2(scale2, scale2), ng = make_float2(neg_m, neg_m);
  float2 sum[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) sum[k] = make_float2(0.f, 0.f);
  uint32_t pk[16];
#pragma unroll
  for (int c2 = 0; c2 < 16; ++c2) {
    const float2 a = __ffma2_rn(make_float2(s[2 * c2], s[2 * c2 + 1]), sc, ng);
    const float2 e = make_float2(ptx::ex2(a.x), ptx::ex2(a.y));
    sum[c2 & 3] = __fadd2_rn(sum[c2 & 3], e);
    pk[c2] = ptx::pack_bf16(e.x, e.y);
  }
  ptx::tmem_st16(t_P, pk);
  const float2 t = __fadd2_rn(__fadd2_rn(sum[0], sum[1]), __fadd2_rn(sum[2], sum[3]));
  return t.x + t.y;
}

// sum of exp2(s * scale2 - m_new) over 32 columns, and their max (for tiles that are entirely "excluded")
__device__ __forceinline__ float max32(const float (&s)[32]) {
  float m[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) m[k] = fmaxf(s[8 * k], s[8 * k + 1]);
#pragma unroll
  for (int c = 2; c < 8; c += 2)
#pragma unroll
    for (int k = 0; k < 4; ++k) m[k] = fmaxf(m[k], fmaxf(s[8 * k + c], s[8 * k + c + 1]));
  return fmaxf(fmaxf(m[0], m[1]), fmaxf(m[2], m[3]));
}
// The same sum with the pairs selected by MASK (bit c2 = columns 2 c2, 2 c2 + 1) exponentiated on the FMA pipe instead of
// the MUFU: 2^a = 2^n 2^r with n = round(a) (magic-number add), r in [-0.5, 0.5], 2^r by a degree-3 polynomial (relative
// error 7.5e-5), the exponent added with an integer shift-add.  For the exponent-SUM-only steps above the diagonal of an
// active handle: they write no P, so their warps have issue slots to spare while the MUFU (one warp instruction per 8
// cycles per sub-partition, shared by the two tiles' warps) is what bounds them.
template <uint32_t MASK>
__device__ __forceinline__ float expsum32_poly(const float (&s)[32], float scale2, float neg_m) {
  const float2 sc = make_float2(scale2, scale2), ng = make_float2(neg_m, neg_m);
  const float2 magic = make_float2(12582912.f, 12582912.f), nmagic = make_float2(-12582912.f, -12582912.f);
  const float2 c3 = make_float2(0.05517027899622917f, 0.05517027899622917f), c2v = make_float2(0.2426076978445053f, 0.2426076978445053f);
  const float2 c1 = make_float2(0.693260908126831f, 0.693260908126831f), c0 = make_float2(0.9999282956123352f, 0.9999282956123352f);
  float2 sum[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) sum[k] = make_float2(0.f, 0.f);
#pragma unroll
  for (int c2 = 0; c2 < 16; ++c2) {
    float2 a = __ffma2_rn(make_float2(s[2 * c2], s[2 * c2 + 1]), sc, ng);
    float2 e;
    if ((MASK >> c2) & 1u) {
      a.x = fmaxf(a.x, -126.f);
      a.y = fmaxf(a.y, -126.f);
      const float2 t = __fadd2_rn(a, magic);                 // integer part in the low mantissa bits
      const float2 n = __fadd2_rn(t, nmagic);
      const float2 r = __fadd2_rn(a, make_float2(-n.x, -n.y));
      float2 q = __ffma2_rn(r, c3, c2v);
      q = __ffma2_rn(q, r, c1);
      q = __ffma2_rn(q, r, c0);
      e.x = __int_as_float(__float_as_int(q.x) + (__float_as_int(t.x) << 23));
      e.y = __int_as_float(__float_as_int(q.y) + (__float_as_int(t.y) << 23));
    } else {
      e = make_float2(ptx::ex2(a.x), ptx::ex2(a.y));
    }
    sum[c2 & 3] = __fadd2_rn(sum[c2 & 3], e);
  }
  const float2 tt = __fadd2_rn(__fadd2_rn(sum[0], sum[1]), __fadd2_rn(sum[2], sum[3]));
  return tt.x + tt.y;
}

__device__ __forceinline__ float expsum32(const float (&s)[32], float scale2, float neg_m) {
  const float2 sc = make_float2(scale2, scale2), ng = make_float2(neg_m, neg_m);
  float2 sum[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) sum[k] = make_float2(0.f, 0.f);
#pragma unroll
  for (int c2 = 0; c2 < 16; ++c2) {
    const float2 a = __ffma2_rn(make_float2(s[2 * c2], s[2 * c2 + 1]), sc, ng);
    sum[c2 & 3] = __fadd2_rn(sum[c2 & 3], make_float2(ptx::ex2(a.x), ptx::ex2(a.y)));
  }
  const float2 t = __fadd2_rn(__fadd2_rn(sum[0], sum[1]), __fadd2_rn(sum[2], sum[3]));
  return t.x + t.y;
}

}  // namespace tc
}  // namespace mimic
