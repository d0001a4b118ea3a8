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
