// Forward, 64-key tiles, two CTAs per SM, EIGHT softmax warps per CTA ("wide softmax").  OPT-IN (MIMIC_FWD_KERNEL=5), an
// experiment kept for the record: it is correct (60 randomised shapes against the SIMT kernels, scripts/fuzz_attention.py) and
// 8-16 % SLOWER than the four-warp kernels (scripts/ab_fwd_wide.py: context pass 703 vs 839 TFLOP/s, T=4096 802 vs 899
// inactive / 802 vs 875 active) — halving every serial phase of the softmax thread does not shorten the step, so that chain is
// not what bounds the forward.  The default path does not reach this file.
//
// Why: in attn_tc_fwd64.cu one thread owns a whole 64-column score row, and the step of the key loop is the serial chain
// of that thread (TMEM load ~400 cycles, row maximum ~290, exp2 / pack / store ~750, fences ~150) — ~2000 cycles per
// 64 keys against 649 cycles of MMA work.  Here two threads share a row: warps 2-5 take columns 0-31 of every score tile,
// warps 6-9 columns 32-63 (warp w and w + 4 sit on the same TMEM lane quarter), so every serial phase halves.  What the two
// halves must agree on is the reference maximum of the row: each writes its 32-column maximum to shared memory, a 64-thread
// named barrier per lane quarter, each reads the other's.  That barrier also orders half 1's P store (TMEM columns 16-31 of
// the tile) after half 0's score load of the same columns.  Row sums, the excluded-key statistics and z1 stay per thread and
// are merged once, in the epilogue.  O columns 0-63 / 64-127 are rescaled and written out by half 0 / half 1.
//
//   TMEM (256 cols): S0 [0,64) S1 [64,128) O [128,256);  P_j = first 32 cols of S[j&1].
//   smem: Q 32 KB | K ring 2 x 16 KB | V ring 2 x 16 KB | barriers, w, shift, exchange buffers, key bits.
#include <cuda.h>

#include "attn_common.cuh"
#include "tc_host.cuh"
#include "tc_ptx.cuh"
#include "tc_softmax.cuh"

namespace mimic {
namespace tc {

using namespace ptx;

namespace f64w {
constexpr int BM = 128, BN = 64, HD = 128, NKV = 2;
constexpr int Q_HALF = 128 * 128;            // [128 rows][128 B]
constexpr int Q_TILE = 2 * Q_HALF;           // 32 KB
constexpr int KV_HALF = BN * 128;            // [64 keys][128 B] = 8 KB
constexpr int KV_TILE = 2 * KV_HALF;         // 16 KB
constexpr int THREADS = 320;                 // TMA warp, MMA warp, 8 softmax warps
constexpr int SOFT_THREADS = 256;
constexpr float LOG2E = 1.4426950408889634f, LN2 = 0.6931471805599453f;
constexpr float RESCALE_THRESHOLD = 8.f;
constexpr int OFF_Q = 0, OFF_K = Q_TILE, OFF_V = OFF_K + NKV * KV_TILE, OFF_BAR = OFF_V + NKV * KV_TILE;
constexpr int OFF_TMEM = OFF_BAR + 24 * 8, OFF_W = OFF_TMEM + 16, OFF_SHIFT = OFF_W + 512;
constexpr int OFF_XCH = OFF_SHIFT + 512;     // row maxima: [2 step parities][2 halves][128 rows] fp32 = 2 KB
constexpr int OFF_FIN = OFF_XCH + 2048;      // epilogue merge: [2 halves][128 rows] float4 (l_c, m_x, l_x, z1) = 4 KB
constexpr int OFF_KBITS = OFF_FIN + 4096;
enum { Q_FULL = 0, K_FULL = 1, K_EMPTY = 3, V_FULL = 5, V_EMPTY = 7, S_FULL = 9, SOFT_DONE = 11, PV_DONE = 13, O_FULL = 14 };
constexpr uint32_t IDESC_QK = idesc_bf16(BM, BN, 0, 0);
constexpr uint32_t IDESC_PV = idesc_bf16(BM, HD, 0, 1);
constexpr uint32_t TMEM_COLS = 256, TM_O = 128;
inline size_t smem_bytes(int S) { return OFF_KBITS + (size_t)((S + BN - 1) / BN) * 8; }

__device__ __forceinline__ uint32_t low_bits(int n) { return n >= 32 ? 0xFFFFFFFFu : (n <= 0 ? 0u : ((1u << n) - 1u)); }

// P = exp2(s * scale2 + neg_m) of 32 columns -> 16 packed TMEM columns at t_P, in two 16-column chunks (the register
// budget is 102 per thread with 640 threads on the SM); returns sum(P)
__device__ __forceinline__ float exp_pack_store32w(const float (&s)[32], float scale2, float neg_m, uint32_t t_P) {
  const float2 sc = make_float2(scale2, scale2), ng = make_float2(neg_m, neg_m);
  float2 sum[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    uint32_t pk[8];
#pragma unroll
    for (int c2 = 0; c2 < 8; ++c2) {
      const float2 a = __ffma2_rn(make_float2(s[16 * i + 2 * c2], s[16 * i + 2 * c2 + 1]), sc, ng);
      const float2 e = make_float2(ex2(a.x), ex2(a.y));
      sum[c2 & 1] = __fadd2_rn(sum[c2 & 1], e);
      pk[c2] = pack_bf16(e.x, e.y);
    }
    tmem_st8(t_P + 8 * i, pk);
  }
  const float2 t = __fadd2_rn(sum[0], sum[1]);
  return t.x + t.y;
}
// 64-thread barrier of the two warps that share a TMEM lane quarter (ids 2..5; constant ids so that only six hardware
// barriers are reserved)
__device__ __forceinline__ void pair_sync(int quarter) {
  switch (quarter) {
    case 0: named_bar_sync(2, 64); break;
    case 1: named_bar_sync(3, 64); break;
    case 2: named_bar_sync(4, 64); break;
    default: named_bar_sync(5, 64); break;
  }
}
}  // namespace f64w

__global__ void __launch_bounds__(f64w::THREADS, 2)
attn_fwd_tc64w_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                      const __grid_constant__ CUtensorMap tm_v, const __grid_constant__ CUtensorMap tm_o, AttnParams p) {
  using namespace f64w;
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) {
    if (threadIdx.x == 0) printf("[mimic tcgen05] dynamic shared memory is not 1024-byte aligned\n");
    __trap();
  }
  uint8_t* sQ = smem + OFF_Q;
  uint8_t* sK = smem + OFF_K;
  uint8_t* sV = smem + OFF_V;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + OFF_TMEM);
  float* w_s = reinterpret_cast<float*>(smem + OFF_W);
  float* shift_s = reinterpret_cast<float*>(smem + OFF_SHIFT);
  float* xch = reinterpret_cast<float*>(smem + OFF_XCH);
  float4* fin = reinterpret_cast<float4*>(smem + OFF_FIN);
  uint32_t* kbits = reinterpret_cast<uint32_t*>(smem + OFF_KBITS);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qt = gridDim.x - 1 - blockIdx.x;          // heaviest (latest) query tiles first
  const int h = blockIdx.y, b = blockIdx.z;
  const int kvh = h / (p.H / p.Hkv);
  const int T = p.T, S = p.S, off = p.S - p.T;
  const int t0 = qt * BM;
  const int n_tiles = (S + BN - 1) / BN;
  int jd = n_tiles - 1;
  if (p.causal) {
    const int smax = min(t0 + BM - 1, T - 1) + off;
    jd = smax < 0 ? -1 : min(n_tiles - 1, smax / BN);
  }
  const int n = p.active ? n_tiles : jd + 1;

  if (warp == 0) {
    if (lane == 0) {
      mbar_init(&bars[Q_FULL], 1);
      for (int i = 0; i < 2; ++i) {
        mbar_init(&bars[K_FULL + i], 1);
        mbar_init(&bars[K_EMPTY + i], 1);
        mbar_init(&bars[V_FULL + i], 1);
        mbar_init(&bars[V_EMPTY + i], 1);
        mbar_init(&bars[S_FULL + i], 1);
        mbar_init(&bars[SOFT_DONE + i], SOFT_THREADS);
      }
      mbar_init(&bars[PV_DONE], 1);
      mbar_init(&bars[O_FULL], 1);
      fence_barrier_init();
      // first ring fill before the CTA-wide sync (consumers only touch the barriers after it)
      mbar_expect_tx(&bars[Q_FULL], Q_TILE);
      tma_load_4d(sQ, &tm_q, &bars[Q_FULL], 0, t0, h, b);
      tma_load_4d(sQ + Q_HALF, &tm_q, &bars[Q_FULL], 64, t0, h, b);
      for (int j = 0; j < NKV && j < n; ++j) {
        mbar_expect_tx(&bars[K_FULL + j], KV_TILE);
        tma_load_4d(sK + j * KV_TILE, &tm_k, &bars[K_FULL + j], 0, j * BN, kvh, b);
        tma_load_4d(sK + j * KV_TILE + KV_HALF, &tm_k, &bars[K_FULL + j], 64, j * BN, kvh, b);
        if (j <= jd) {
          mbar_expect_tx(&bars[V_FULL + j], KV_TILE);
          tma_load_4d(sV + j * KV_TILE, &tm_v, &bars[V_FULL + j], 0, j * BN, kvh, b);
          tma_load_4d(sV + j * KV_TILE + KV_HALF, &tm_v, &bars[V_FULL + j], 64, j * BN, kvh, b);
        }
      }
    }
    __syncwarp();
    tmem_alloc(tmem_slot, TMEM_COLS);
  } else if (warp >= 2) {
    const int st = threadIdx.x - 64;
    if (p.active && st < HD) {
      w_s[st] = p.w[h * HD + st];
      shift_s[st] = p.shift[h * HD + st];
    }
    // key-validity bits, one 32-key word per (warp, u): up to 16 independent byte loads in flight per lane
    const uint8_t* km = p.key_mask ? p.key_mask + (int64_t)b * S : nullptr;
    for (int w0 = warp - 2; w0 < n_tiles * 2; w0 += 8 * 16) {
      uint8_t ok[16];
#pragma unroll
      for (int u = 0; u < 16; ++u) {
        const int s = (w0 + 8 * u) * 32 + lane;
        ok[u] = (s < S) ? (km == nullptr ? (uint8_t)1 : km[s]) : (uint8_t)0;
      }
#pragma unroll
      for (int u = 0; u < 16; ++u) {
        const int wi = w0 + 8 * u;
        const uint32_t word = __ballot_sync(0xffffffffu, ok[u] != 0);
        if (lane == 0 && wi < n_tiles * 2) kbits[wi] = word;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_O = tmem_base + TM_O;

  if (warp == 0) {
    // ===================== TMA producer: K one tile ahead of V (the order the stages free up in) =====================
    if (lane == 0) {
      for (int j = NKV; j <= n; ++j) {
        if (j < n) {
          const int st = j & 1;
          mbar_wait(&bars[K_EMPTY + st], ((j >> 1) & 1) ^ 1, 100 + st);
          mbar_expect_tx(&bars[K_FULL + st], KV_TILE);
          tma_load_4d(sK + st * KV_TILE, &tm_k, &bars[K_FULL + st], 0, j * BN, kvh, b);
          tma_load_4d(sK + st * KV_TILE + KV_HALF, &tm_k, &bars[K_FULL + st], 64, j * BN, kvh, b);
        }
        const int jv = j - 1;
        if (jv >= NKV && jv <= jd) {
          const int sv = jv & 1;
          mbar_wait(&bars[V_EMPTY + sv], ((jv >> 1) & 1) ^ 1, 110 + sv);
          mbar_expect_tx(&bars[V_FULL + sv], KV_TILE);
          tma_load_4d(sV + sv * KV_TILE, &tm_v, &bars[V_FULL + sv], 0, jv * BN, kvh, b);
          tma_load_4d(sV + sv * KV_TILE + KV_HALF, &tm_v, &bars[V_FULL + sv], 64, jv * BN, kvh, b);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (identical to attn_tc_fwd64.cu) =====================
    if (n > 0) {
      constexpr uint32_t HI = desc_hi_sw128(1024);
      const uint32_t q_lo = desc_lo(smem_u32(sQ), 16), k_lo0 = desc_lo(smem_u32(sK), 16);
      const uint32_t v_lo0 = desc_lo(smem_u32(sV), KV_HALF);            // V as MN-major B: d halves 8 KB apart
      mbar_wait(&bars[Q_FULL], 0, 200);
      auto issue_qk = [&](int j) {
        const int st = j & 1;
        mbar_wait(&bars[K_FULL + st], (j >> 1) & 1, 210 + st);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t k_lo = k_lo0 + st * (KV_TILE >> 4);
          const uint32_t d_s = tmem_base + st * BN;
#pragma unroll
          for (int k = 0; k < HD / 16; ++k) {
            const uint32_t qoff = ((k >> 2) * Q_HALF + (k & 3) * 32) >> 4, koff = ((k >> 2) * KV_HALF + (k & 3) * 32) >> 4;
            if (k == 0) mma_ss_c<false>(d_s, desc64(HI, q_lo + qoff), desc64(HI, k_lo + koff), IDESC_QK);
            else mma_ss_c<true>(d_s, desc64(HI, q_lo + qoff), desc64(HI, k_lo + koff), IDESC_QK);
          }
          tc_commit(&bars[K_EMPTY + st]);
          tc_commit(&bars[S_FULL + st]);
        }
        __syncwarp();
      };
      issue_qk(0);
      if (n > 1) issue_qk(1);
      for (int j = 0; j < n; ++j) {
        const int st = j & 1;
        mbar_wait(&bars[SOFT_DONE + st], (j >> 1) & 1, 220 + st);
        tc_fence_after();
        if (j <= jd) {
          mbar_wait(&bars[V_FULL + st], (j >> 1) & 1, 230 + st);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t v_lo = v_lo0 + st * (KV_TILE >> 4);
            const uint32_t a_p = tmem_base + st * BN;
            if (j > 0) {
#pragma unroll
              for (int k = 0; k < BN / 16; ++k) mma_ts_c<true>(tmem_O, a_p + k * 8, desc64(HI, v_lo + k * (2048 >> 4)), IDESC_PV);
            } else {
#pragma unroll
              for (int k = 0; k < BN / 16; ++k) {
                if (k == 0) mma_ts_c<false>(tmem_O, a_p, desc64(HI, v_lo), IDESC_PV);
                else mma_ts_c<true>(tmem_O, a_p + k * 8, desc64(HI, v_lo + k * (2048 >> 4)), IDESC_PV);
              }
            }
            tc_commit(&bars[V_EMPTY + st]);
            tc_commit(&bars[PV_DONE]);
          }
          __syncwarp();
        }
        if (j + 2 < n) issue_qk(j + 2);
      }
      if (elect_one()) tc_commit(&bars[O_FULL]);
      __syncwarp();
    }
  } else {
    // ===================== softmax / correction / epilogue: TWO threads per query row =====================
    const int quarter = warp & 3;                        // TMEM lane quarter this warp may touch
    const int half = (warp - 2) >> 2;                    // 0: columns 0-31 of every score tile, 1: columns 32-63
    const int r = quarter * 32 + lane;
    const int t = t0 + r;
    const bool row_ok = t < T;
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    const float scale2 = p.scale * LOG2E;

    float z1 = 0.f;
    mbar_wait(&bars[Q_FULL], 0, 300);
    if (p.active && half == 0) {
#pragma unroll
      for (int hf = 0; hf < 2; ++hf)
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const uint4 u = *reinterpret_cast<const uint4*>(sQ + hf * Q_HALF + r * 128 + ((c ^ (r & 7)) << 4));
          const __nv_bfloat162* hh = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 f = __bfloat1622float2(hh[e]);
            const int col = hf * 64 + c * 8 + e * 2;
            z1 = fmaf(f.x, w_s[col], z1);
            z1 = fmaf(f.y, w_s[col + 1], z1);
          }
        }
      z1 += p.bias[h];
    }

    // m_c: reference maximum of the row (identical in both halves); l_c, m_x, l_x: this thread's 32 columns only
    float m_c = -INFINITY, l_c = 0.f, m_x = -INFINITY, l_x = 0.f;   // log2 domain
    for (int j = 0; j < n; ++j) {
      const int st = j & 1;
      const int key0 = j * BN + 32 * half;
      const uint32_t inr = low_bits(S - key0);
      const uint32_t cz = p.causal ? low_bits(t + off - key0 + 1) : 0xFFFFFFFFu;
      const uint32_t vw = row_ok ? (kbits[j * 2 + half] & cz) : 0u;
      const uint32_t xw = (p.active && row_ok) ? (inr & ~vw) : 0u;
      const bool all_vis = __all_sync(0xffffffffu, vw == 0xFFFFFFFFu);
      const bool any_exc = __any_sync(0xffffffffu, xw != 0u);
      mbar_wait(&bars[S_FULL + st], (j >> 1) & 1, 310 + st);
      tc_fence_after();
      const uint32_t taddr = tmem_base + lane_off + st * BN;      // start of the tile; this thread's scores at + 32 * half
      float s[32];
      tmem_ld32(taddr + 32 * half, s);
      tc_wait_ld();
      if (j <= jd) {
        float mv = -INFINITY;
        if (all_vis) {
          mv = max32(s);
        } else {
#pragma unroll
          for (int c = 0; c < 32; ++c)
            if ((vw >> c) & 1u) mv = fmaxf(mv, s[c]);
        }
        // the row's maximum over both halves; the barrier also orders the partner's score load before this thread's P store
        float* x = xch + (j & 1) * 256;
        x[half * 128 + r] = mv;
        pair_sync(quarter);
        mv = fmaxf(mv, x[(half ^ 1) * 128 + r]) * scale2;
        const float m_new = fmaxf(m_c, mv);
        const bool need = (m_c != -INFINITY) && (m_new > m_c + RESCALE_THRESHOLD);
        if (m_c == -INFINITY) m_c = m_new;
        if (__any_sync(0xffffffffu, need)) {            // lazy rescale: this half's 64 columns of O and its partial row sum
          const float alpha = need ? ex2(m_c - m_new) : 1.f;
          if (j > 0) {
            mbar_wait(&bars[PV_DONE], (j - 1) & 1, 320);
            tc_fence_after();
#pragma unroll
            for (int i = 0; i < 2; ++i) {
              float o[32];
              tmem_ld32(tmem_O + lane_off + 64 * half + 32 * i, o);
              tc_wait_ld();
              uint32_t ou[32];
#pragma unroll
              for (int c = 0; c < 32; ++c) ou[c] = __float_as_uint(o[c] * alpha);
              tmem_st32(tmem_O + lane_off + 64 * half + 32 * i, ou);
            }
          }
          if (need) {
            l_c *= alpha;
            m_c = m_new;
          }
        }
        const float neg_m = (m_c == -INFINITY) ? 0.f : -m_c;
        const uint32_t t_P = taddr + 16 * half;          // 32 keys = 16 packed columns
        if (all_vis) {
          l_c += exp_pack_store32w(s, scale2, neg_m, t_P);
        } else {
          float rowsum = 0.f;
          uint32_t pk[16];
#pragma unroll
          for (int c2 = 0; c2 < 16; ++c2) {
            float p0 = ex2(fmaf(s[2 * c2], scale2, neg_m)), p1 = ex2(fmaf(s[2 * c2 + 1], scale2, neg_m));
            if (!((vw >> (2 * c2)) & 1u)) p0 = 0.f;
            if (!((vw >> (2 * c2 + 1)) & 1u)) p1 = 0.f;
            rowsum += p0 + p1;
            pk[c2] = pack_bf16(p0, p1);
          }
          tmem_st16(t_P, pk);
          l_c += rowsum;
        }
        if (any_exc) {                                 // excluded keys inside a tile that also has visible ones
          float mx = -INFINITY;
#pragma unroll
          for (int c = 0; c < 32; ++c)
            if ((xw >> c) & 1u) mx = fmaxf(mx, s[c]);
          mx *= scale2;
          const float mx_new = fmaxf(m_x, mx);
          float es = 0.f;
#pragma unroll
          for (int c = 0; c < 32; ++c)
            if ((xw >> c) & 1u) es += ex2(fmaf(s[c], scale2, -mx_new));
          l_x = l_x * ((m_x == -INFINITY) ? 0.f : ex2(m_x - mx_new)) + es;
          m_x = mx_new;
        }
        tc_wait_st();
      } else if (__all_sync(0xffffffffu, xw == 0xFFFFFFFFu)) {
        // tile above the causal diagonal, every key of this half excluded: row-sum of exp only
        const float mx_new = fmaxf(m_x, max32(s) * scale2);
        l_x = l_x * ((m_x == -INFINITY) ? 0.f : ex2(m_x - mx_new)) + expsum32(s, scale2, -mx_new);
        m_x = mx_new;
      } else if (any_exc) {
        float mx = -INFINITY;
#pragma unroll
        for (int c = 0; c < 32; ++c)
          if ((xw >> c) & 1u) mx = fmaxf(mx, s[c]);
        mx *= scale2;
        const float mx_new = fmaxf(m_x, mx);
        float es = 0.f;
#pragma unroll
        for (int c = 0; c < 32; ++c) {
          const float e = ex2(fmaf(s[c], scale2, -mx_new));
          es += ((xw >> c) & 1u) ? e : 0.f;
        }
        l_x = l_x * ((m_x == -INFINITY) ? 0.f : ex2(m_x - mx_new)) + es;
        m_x = mx_new;
      }
      tc_fence_before();
      mbar_arrive(&bars[SOFT_DONE + st]);
    }

    // ---- epilogue: merge the two halves of every row, then each half writes its 64 output columns ----
    fin[half * 128 + r] = make_float4(l_c, m_x, l_x, z1);
    pair_sync(quarter);
    {
      const float4 o4 = fin[(half ^ 1) * 128 + r];
      l_c += o4.x;                                       // same reference maximum in both halves
      const float mx = fmaxf(m_x, o4.y);
      l_x = (mx == -INFINITY) ? 0.f : l_x * ((m_x == -INFINITY) ? 0.f : ex2(m_x - mx)) +
                                         o4.z * ((o4.y == -INFINITY) ? 0.f : ex2(o4.y - mx));
      m_x = mx;
      if (half == 1) z1 = o4.w;
    }
    const float inv = l_c > 0.f ? 1.f / l_c : 0.f;
    const float lse_c = l_c > 0.f ? (m_c + log2f(l_c)) * LN2 : -INFINITY;
    float mu = 0.f, lse_full = 0.f;
    if (p.active) {
      const float lse_x = l_x > 0.f ? (m_x + log2f(l_x)) * LN2 : -INFINITY;
      const float hi = fmaxf(lse_c, lse_x), lo = fminf(lse_c, lse_x);
      lse_full = (lo == -INFINITY) ? hi : hi + log1pf(__expf(lo - hi));
      mu = 1.f / (1.f + __expf(lse_full - z1));
    }
    if (jd >= 0) {
      mbar_wait(&bars[O_FULL], 0, 330);
      tc_fence_after();
    }
    // the output tile goes through the dead Q tile: half 0 fills the first 64-column box, half 1 the second
    uint8_t* srow = sQ + half * Q_HALF + r * 128;          // 128B-swizzled [128 rows][64 cols] box, as the TMA map expects
#pragma unroll
    for (int i2 = 0; i2 < 2; ++i2) {
      float o[32];
      if (jd >= 0) {
        tmem_ld32(tmem_O + lane_off + 64 * half + 32 * i2, o);
        tc_wait_ld();
      } else {
#pragma unroll
        for (int c = 0; c < 32; ++c) o[c] = 0.f;
      }
#pragma unroll
      for (int c8 = 0; c8 < 4; ++c8) {
        uint32_t pk[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int c = 64 * half + 32 * i2 + 8 * c8 + 2 * e;
          const float f0 = o[8 * c8 + 2 * e] * inv + (p.active ? mu * shift_s[c] : 0.f);
          const float f1 = o[8 * c8 + 2 * e + 1] * inv + (p.active ? mu * shift_s[c + 1] : 0.f);
          pk[e] = pack_bf16(f0, f1);
        }
        const int chunk = 4 * i2 + c8;                    // 16-byte chunk of this row's 128-byte line
        *reinterpret_cast<uint4*>(srow + ((chunk ^ (r & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      }
    }
    fence_proxy_async();                                 // generic-proxy smem writes -> visible to the TMA (async proxy)
    named_bar_sync(1, SOFT_THREADS);
    if (threadIdx.x == 64) {
      tma_store_4d(&tm_o, sQ, 0, t0, h, b);
      tma_store_4d(&tm_o, sQ + Q_HALF, 64, t0, h, b);
      bulk_commit();
      bulk_wait_read0();                                 // shared memory may be released (CTA exit) once the reads are done
    }
    if (row_ok && half == 0) {
      const int64_t ridx = ((int64_t)b * p.H + h) * T + t;
      p.lse_c[ridx] = lse_c;
      if (p.active) {
        p.lse_full[ridx] = lse_full;
        p.mu[ridx] = mu;
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

int fwd64w(const AttnParams& p, cudaStream_t st) {
  using namespace f64w;
  CUtensorMap tq, tk, tv, to;
  int rc;
  if ((rc = make_map(&tq, p.q, p.B, p.H, p.T, p.q_sb, p.q_sh, p.q_st, BM))) return rc;
  if ((rc = make_map(&to, p.out, p.B, p.H, p.T, p.o_sb, p.o_sh, p.o_st, BM))) return rc;
  if ((rc = make_map(&tk, p.k, p.B, p.Hkv, p.S, p.k_sb, p.k_sh, p.k_ss, BN))) return rc;
  if ((rc = make_map(&tv, p.v, p.B, p.Hkv, p.S, p.v_sb, p.v_sh, p.v_ss, BN))) return rc;
  const size_t smem = smem_bytes(p.S);
  static size_t configured = 0;
  if (smem > configured) {
    MIMIC_CUDA_OK(cudaFuncSetAttribute(attn_fwd_tc64w_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  dim3 grid((p.T + BM - 1) / BM, p.H, p.B);
  attn_fwd_tc64w_kernel<<<grid, THREADS, smem, st>>>(tq, tk, tv, to, p);
  MIMIC_LAUNCH_OK("attn_fwd_tc64w_kernel");
  return MIMIC_OK;
}

}  // namespace tc
}  // namespace mimic
