// Shifted self-attention on Blackwell tensor cores: tcgen05.mma with TMEM accumulators, operands
// staged by TMA (128B swizzle), warp-specialised (TMA producer / MMA issuer / 4 softmax warps).
// bf16 I/O, head_dim 128, fp32 statistics.  Math: SURVEY.md Appendix A; reference call sites
// src/shift_encoder.py:352-359 (SDPA) + 682-732 (do_shift).
//
// Forward, one CTA per (128-query tile, head, batch):
//   S_j = Q K_j^T          tcgen05.mma SS (Q, K_j K-major in smem)        -> TMEM S[j&1]  (128 cols fp32)
//   softmax warps          tcgen05.ld S row -> two DISJOINT online statistics:
//                            visible keys  (m_c, l_c): P = exp2(s - m_c)  -> bf16 P back into TMEM (aliases S)
//                            excluded keys (m_x, l_x): future / masked keys, only their exp-sum
//   O += P_j V_j           tcgen05.mma TS (P from TMEM, V_j MN-major in smem) -> TMEM O (128 cols fp32)
//   tiles above the causal diagonal are QK^T + row-sum only (no P, no PV) and are skipped entirely
//   when the handle is inactive.
//   epilogue               out = O / l_c + sigmoid(z1 - lse_full) * shift,  lse_full = logaddexp(lse_c, lse_x)
// TMEM map (512 cols): S0 [0,128) S1 [128,256) O [256,384); P_j = first 64 cols of S[j&1].
#include <cuda.h>

#include <cstdlib>

#include "attn_common.cuh"
#include "tc_host.cuh"
#include "tc_ptx.cuh"
#include "tc_softmax.cuh"

namespace mimic {
namespace tc {

using namespace ptx;

constexpr int BM = 128, BN = 128, HD = 128;
constexpr int HALF_BYTES = 128 * 64 * 2;   // one 64-column half of a 128-row tile: [128][128 B], 128B-swizzled
constexpr int TILE_BYTES = 2 * HALF_BYTES;
constexpr int FWD_THREADS = 192;
constexpr float LOG2E = 1.4426950408889634f, LN2 = 0.6931471805599453f;
constexpr float RESCALE_THRESHOLD = 8.f;    // lazy O rescale: only when the running max grows by > 2^8

// K / V rings have NKV stages each (3 when shared memory allows: the TMA round trip is longer than one tile of MMA work)
enum FwdBar { Q_FULL = 0, K_FULL = 1, K_EMPTY = 4, V_FULL = 7, V_EMPTY = 10, S_FULL = 13, SOFT_DONE = 15, PV_DONE = 17, O_FULL = 18, N_FWD_BARS = 19 };

constexpr uint32_t IDESC_QK = idesc_bf16(BM, BN, 0, 0);   // A = Q (K-major), B = K (K-major)
constexpr uint32_t IDESC_PV = idesc_bf16(BM, HD, 0, 1);   // A = P (TMEM),    B = V (MN-major: [keys][d])

__device__ __forceinline__ uint32_t low_bits(int n) { return n >= 32 ? 0xFFFFFFFFu : (n <= 0 ? 0u : ((1u << n) - 1u)); }

template <int NKV>
struct FwdSmem {
  static constexpr int OFF_Q = 0, OFF_K = TILE_BYTES, OFF_V = (1 + NKV) * TILE_BYTES, OFF_BAR = (1 + 2 * NKV) * TILE_BYTES;
  static constexpr int OFF_TMEM = OFF_BAR + 24 * 8, OFF_W = OFF_TMEM + 16, OFF_SHIFT = OFF_W + 512, OFF_KBITS = OFF_SHIFT + 512;
  // the dynamic shared window starts 1024-byte aligned (no static __shared__ in this kernel; checked at entry)
  static size_t bytes(int S) { return OFF_KBITS + (size_t)((S + BN - 1) / BN) * 4 * 4; }
};
constexpr size_t MAX_DYN_SMEM = 232448;

template <int NKV>
__global__ void __launch_bounds__(FWD_THREADS, 1)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                   const __grid_constant__ CUtensorMap tm_v, AttnParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  using FwdSmem = FwdSmem<NKV>;
  if ((smem_u32(smem) & 1023u) != 0) {
    if (threadIdx.x == 0) printf("[mimic tcgen05] dynamic shared memory is not 1024-byte aligned\n");
    __trap();
  }
  uint8_t* sQ = smem + FwdSmem::OFF_Q;
  uint8_t* sK = smem + FwdSmem::OFF_K;
  uint8_t* sV = smem + FwdSmem::OFF_V;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + FwdSmem::OFF_BAR);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + FwdSmem::OFF_TMEM);
  float* w_s = reinterpret_cast<float*>(smem + FwdSmem::OFF_W);
  float* shift_s = reinterpret_cast<float*>(smem + FwdSmem::OFF_SHIFT);
  uint32_t* kbits = reinterpret_cast<uint32_t*>(smem + FwdSmem::OFF_KBITS);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qt = gridDim.x - 1 - blockIdx.x;          // heaviest (latest) query tiles first
  const int h = blockIdx.y, b = blockIdx.z;
  const int kvh = h / (p.H / p.Hkv);
  const int T = p.T, S = p.S, off = p.S - p.T;
  const int t0 = qt * BM;
  const int n_tiles = (S + BN - 1) / BN;
  int jd = n_tiles - 1;                               // last key tile with a visible entry
  if (p.causal) {
    const int smax = min(t0 + BM - 1, T - 1) + off;
    jd = smax < 0 ? -1 : min(n_tiles - 1, smax / BN);
  }
  const int n = p.active ? n_tiles : jd + 1;          // key tiles this CTA visits

  if (warp == 0) {
    if (lane == 0) {
      mbar_init(&bars[Q_FULL], 1);
      for (int i = 0; i < NKV; ++i) {
        mbar_init(&bars[K_FULL + i], 1);
        mbar_init(&bars[K_EMPTY + i], 1);
        mbar_init(&bars[V_FULL + i], 1);
        mbar_init(&bars[V_EMPTY + i], 1);
      }
      for (int i = 0; i < 2; ++i) {
        mbar_init(&bars[S_FULL + i], 1);
        mbar_init(&bars[SOFT_DONE + i], 128);
      }
      mbar_init(&bars[PV_DONE], 1);
      mbar_init(&bars[O_FULL], 1);
      fence_barrier_init();
      prefetch_tmap(&tm_q);
      prefetch_tmap(&tm_k);
      prefetch_tmap(&tm_v);
    }
    __syncwarp();
    tmem_alloc(tmem_slot, 512);
  } else if (warp >= 2) {
    const int st = threadIdx.x - 64;                  // 0..127
    if (p.active) {
      w_s[st] = p.w[h * HD + st];
      shift_s[st] = p.shift[h * HD + st];
    }
    // key-validity bit words for this batch row: bit = key in range and not masked
    const uint8_t* km = p.key_mask ? p.key_mask + (int64_t)b * S : nullptr;
    for (int wi = warp - 2; wi < n_tiles * 4; wi += 4) {
      const int s = wi * 32 + lane;
      const bool ok = s < S && (km == nullptr || km[s] != 0);
      const uint32_t word = __ballot_sync(0xffffffffu, ok);
      if (lane == 0) kbits[wi] = word;
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_O = tmem_base + 256;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      mbar_expect_tx(&bars[Q_FULL], TILE_BYTES);
      tma_load_4d(sQ, &tm_q, &bars[Q_FULL], 0, t0, h, b);
      tma_load_4d(sQ + HALF_BYTES, &tm_q, &bars[Q_FULL], 64, t0, h, b);
      for (int j = 0; j < n; ++j) {
        const int st = j % NKV;
        const uint32_t ph = (j / NKV) & 1;
        mbar_wait(&bars[K_EMPTY + st], ph ^ 1, 100 + st);
        mbar_expect_tx(&bars[K_FULL + st], TILE_BYTES);
        tma_load_4d(sK + st * TILE_BYTES, &tm_k, &bars[K_FULL + st], 0, j * BN, kvh, b);
        tma_load_4d(sK + st * TILE_BYTES + HALF_BYTES, &tm_k, &bars[K_FULL + st], 64, j * BN, kvh, b);
        if (j <= jd) {
          mbar_wait(&bars[V_EMPTY + st], ph ^ 1, 110 + st);
          mbar_expect_tx(&bars[V_FULL + st], TILE_BYTES);
          tma_load_4d(sV + st * TILE_BYTES, &tm_v, &bars[V_FULL + st], 0, j * BN, kvh, b);
          tma_load_4d(sV + st * TILE_BYTES + HALF_BYTES, &tm_v, &bars[V_FULL + st], 64, j * BN, kvh, b);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // All 32 lanes run the control flow (warp-uniform waits and address arithmetic); one elected lane issues.
    if (n > 0) {
      constexpr uint32_t HI = desc_hi_sw128(1024);
      const uint32_t q_lo = desc_lo(smem_u32(sQ), 16), k_lo0 = desc_lo(smem_u32(sK), 16);
      const uint32_t v_lo0 = desc_lo(smem_u32(sV), HALF_BYTES);          // V as MN-major B: d halves 16 KB apart
      mbar_wait(&bars[Q_FULL], 0, 200);
      auto issue_qk = [&](int j) {
        const int st = j & 1, ks = j % NKV;
        mbar_wait(&bars[K_FULL + ks], (j / NKV) & 1, 210 + ks);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t k_lo = k_lo0 + ks * (TILE_BYTES >> 4);
          const uint32_t d_s = tmem_base + st * BN;
          for (int rep = 0; rep <= (p.debug_flags & 255); ++rep) {   // debug_flags > 0: re-issue QK^T (timing experiments)
#pragma unroll
          for (int k = 0; k < HD / 16; ++k) {
            const uint32_t koff = ((k >> 2) * HALF_BYTES + (k & 3) * 32) >> 4;
            if (k == 0) mma_ss_c<false>(d_s, desc64(HI, q_lo + koff), desc64(HI, k_lo + koff), IDESC_QK);
            else mma_ss_c<true>(d_s, desc64(HI, q_lo + koff), desc64(HI, k_lo + koff), IDESC_QK);
          }
          }
          tc_commit(&bars[K_EMPTY + ks]);
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
          const int vs = j % NKV;
          mbar_wait(&bars[V_FULL + vs], (j / NKV) & 1, 230 + vs);
          tc_fence_after();
          if (elect_one()) {
            // A: P_j in TMEM, 16 bf16 of K per step = 8 columns; B: V_j [keys][d], 16 keys = 2048 B per step
            const uint32_t v_lo = v_lo0 + vs * (TILE_BYTES >> 4);
            const uint32_t a_p = tmem_base + st * BN;
            for (int rep = 0; rep < ((p.debug_flags >> 8) & 255); ++rep) {   // debug: re-issue PV (timing experiments only)
#pragma unroll
              for (int k = 0; k < BN / 16; ++k) mma_ts_c<true>(tmem_O, a_p + k * 8, desc64(HI, v_lo + k * (2048 >> 4)), IDESC_PV);
            }
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
            tc_commit(&bars[V_EMPTY + vs]);
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
    // ===================== softmax / correction / epilogue: one query row per thread =====================
    const int quarter = warp & 3;                     // TMEM lane quarter this warp may touch
    const int r = quarter * 32 + lane;
    const int t = t0 + r;
    const bool row_ok = t < T;
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    const float scale2 = p.scale * LOG2E;

    float z1 = 0.f;
    if (n > 0 || true) {
      mbar_wait(&bars[Q_FULL], 0, 300);
      if (p.active) {
        // q row r from the swizzled tile: 16-byte chunk c of a 128-byte row sits at chunk (c ^ (r & 7))
#pragma unroll
        for (int hf = 0; hf < 2; ++hf)
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const uint4 u = *reinterpret_cast<const uint4*>(sQ + hf * HALF_BYTES + r * 128 + ((c ^ (r & 7)) << 4));
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
    }

    float m_c = -INFINITY, l_c = 0.f, m_x = -INFINITY, l_x = 0.f;   // log2 domain
    for (int j = 0; j < n; ++j) {
      const int st = j & 1;
      mbar_wait(&bars[S_FULL + st], (j >> 1) & 1, 310 + st);
      tc_fence_after();
      const uint32_t taddr = tmem_base + lane_off + st * BN;
      const int key0 = j * BN;
      uint32_t vw[4], xw[4];                          // visible / excluded bit words of this row in this tile
      bool all_vis = row_ok, any_exc = false;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const uint32_t inr = low_bits(S - key0 - 32 * i);
        const uint32_t cz = p.causal ? low_bits(t + off - key0 - 32 * i + 1) : 0xFFFFFFFFu;
        vw[i] = row_ok ? (kbits[j * 4 + i] & cz) : 0u;
        xw[i] = (p.active && row_ok) ? (inr & ~vw[i]) : 0u;
        all_vis = all_vis && (vw[i] == 0xFFFFFFFFu);
        any_exc = any_exc || (xw[i] != 0u);
      }
      const bool all_vis_w = __all_sync(0xffffffffu, all_vis);    // warp-uniform: mask-free fast path
      if (j <= jd) {
        float s[128];
        tmem_ld32(taddr, *reinterpret_cast<float(*)[32]>(s));
        tmem_ld32(taddr + 32, *reinterpret_cast<float(*)[32]>(s + 32));
        tmem_ld32(taddr + 64, *reinterpret_cast<float(*)[32]>(s + 64));
        tmem_ld32(taddr + 96, *reinterpret_cast<float(*)[32]>(s + 96));
        tc_wait_ld();
        for (int rep = 0; rep < (p.debug_flags >> 16); ++rep) {    // debug: extra TMEM reads (timing experiments only)
          tmem_ld32(taddr, *reinterpret_cast<float(*)[32]>(s));
          tmem_ld32(taddr + 32, *reinterpret_cast<float(*)[32]>(s + 32));
          tmem_ld32(taddr + 64, *reinterpret_cast<float(*)[32]>(s + 64));
          tmem_ld32(taddr + 96, *reinterpret_cast<float(*)[32]>(s + 96));
          tc_wait_ld();
        }
        float mv = -INFINITY;
        if (all_vis_w) {
          mv = row_max128(s);
        } else {
#pragma unroll
          for (int c = 0; c < 128; ++c)
            if ((vw[c >> 5] >> (c & 31)) & 1u) mv = fmaxf(mv, s[c]);
        }
        mv *= scale2;
        const float m_new = fmaxf(m_c, mv);
        // lazy rescale of O and l_c (only when the max grew by more than 2^8; warp-collective TMEM ops)
        const bool need = (m_c != -INFINITY) && (m_new > m_c + RESCALE_THRESHOLD);
        if (m_c == -INFINITY) m_c = m_new;
        if (__any_sync(0xffffffffu, need)) {
          const float alpha = need ? ex2(m_c - m_new) : 1.f;
          if (j > 0) {
            mbar_wait(&bars[PV_DONE], (j - 1) & 1, 320);
            tc_fence_after();
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              float o[32];
              tmem_ld32(tmem_O + lane_off + 32 * i, o);
              tc_wait_ld();
              uint32_t ou[32];
#pragma unroll
              for (int c = 0; c < 32; ++c) ou[c] = __float_as_uint(o[c] * alpha);
              tmem_st32(tmem_O + lane_off + 32 * i, ou);
            }
          }
          if (need) {
            l_c *= alpha;
            m_c = m_new;
          }
        }
        // P = exp2(s*scale2 - m_c) on visible keys -> bf16 pairs -> TMEM (first 64 columns of this S buffer)
        const float neg_m = (m_c == -INFINITY) ? 0.f : -m_c;
        if (all_vis_w) {
          l_c += exp_pack_store128(s, scale2, neg_m, taddr);
        } else {
          float rowsum = 0.f;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            uint32_t pk[16];
#pragma unroll
            for (int c2 = 0; c2 < 16; ++c2) {
              const int c = 32 * i + 2 * c2;
              float p0 = ex2(fmaf(s[c], scale2, neg_m)), p1 = ex2(fmaf(s[c + 1], scale2, neg_m));
              if (!((vw[i] >> (2 * c2)) & 1u)) p0 = 0.f;
              if (!((vw[i] >> (2 * c2 + 1)) & 1u)) p1 = 0.f;
              rowsum += p0 + p1;
              pk[c2] = pack_bf16(p0, p1);
            }
            tmem_st16(taddr + 16 * i, pk);
          }
          l_c += rowsum;
        }
        if (any_exc) {                                 // excluded keys inside a tile that also has visible ones
          float mx = -INFINITY;
#pragma unroll
          for (int c = 0; c < 128; ++c)
            if ((xw[c >> 5] >> (c & 31)) & 1u) mx = fmaxf(mx, s[c]);
          mx *= scale2;
          const float mx_new = fmaxf(m_x, mx);
          float es = 0.f;
#pragma unroll
          for (int c = 0; c < 128; ++c)
            if ((xw[c >> 5] >> (c & 31)) & 1u) es += ex2(fmaf(s[c], scale2, -mx_new));
          l_x = l_x * ((m_x == -INFINITY) ? 0.f : ex2(m_x - mx_new)) + es;
          m_x = mx_new;
        }
        tc_wait_st();
      } else {
        // tile above the causal diagonal: every in-range key is excluded; row-sum of exp only
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float s[32];
          tmem_ld32(taddr + 32 * i, s);
          tc_wait_ld();
          if (__all_sync(0xffffffffu, xw[i] == 0xFFFFFFFFu)) {
            const float mx_new = fmaxf(m_x, max32(s) * scale2);
            l_x = l_x * ((m_x == -INFINITY) ? 0.f : ex2(m_x - mx_new)) + expsum32(s, scale2, -mx_new);
            m_x = mx_new;
          } else if (xw[i] != 0u) {
            float mx = -INFINITY;
#pragma unroll
            for (int c = 0; c < 32; ++c)
              if ((xw[i] >> c) & 1u) mx = fmaxf(mx, s[c]);
            mx *= scale2;
            const float mx_new = fmaxf(m_x, mx);
            float es = 0.f;
#pragma unroll
            for (int c = 0; c < 32; ++c) {
              const float e = ex2(fmaf(s[c], scale2, -mx_new));
              es += ((xw[i] >> c) & 1u) ? e : 0.f;
            }
            l_x = l_x * ((m_x == -INFINITY) ? 0.f : ex2(m_x - mx_new)) + es;
            m_x = mx_new;
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&bars[SOFT_DONE + st]);
    }

    // ---- epilogue ----
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
    __nv_bfloat16* orow = reinterpret_cast<__nv_bfloat16*>(p.out) + (int64_t)b * p.o_sb + (int64_t)t * p.o_st + (int64_t)h * p.o_sh;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float o[32];
      if (jd >= 0) {
        tmem_ld32(tmem_O + lane_off + 32 * i, o);
        tc_wait_ld();
      } else {
#pragma unroll
        for (int c = 0; c < 32; ++c) o[c] = 0.f;
      }
      if (row_ok) {
#pragma unroll
        for (int c8 = 0; c8 < 4; ++c8) {
          float f[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const int c = 32 * i + 8 * c8 + e;
            f[e] = o[8 * c8 + e] * inv + (p.active ? mu * shift_s[c] : 0.f);
          }
          store8(orow + 32 * i + 8 * c8, f);
        }
      }
    }
    if (row_ok) {
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
    tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------------
// host side
bool supports(const AttnParams& p, bool backward) {
  (void)backward;
  return p.dtype == MIMIC_BF16 && p.d == HD && p.S <= (1 << 18) && p.T >= 1;
}

int fwd64(const AttnParams& p, cudaStream_t st);  // attn_tc_fwd64.cu: 64-key tiles, two CTAs per SM
int fwd64p(const AttnParams& p, cudaStream_t st); // attn_tc_fwd64p.cu: the same pipeline in persistent CTAs (static work lists)

int fwdp(const AttnParams& p, cudaStream_t st);   // attn_tc_fwdp.cu: persistent, two 128-row query tiles ping-ponged over 128-key tiles
bool fwdp_supports(const AttnParams& p, bool forced);

// Forward kernel choice.  MIMIC_FWD_KERNEL (read ONCE per process; A/B scripts set it before the first call) forces a family:
// "1" single-tile, 128-key tiles / "3" 64-key tiles, two CTAs per SM (persistent when it pays) / "4" that one persistent
// always / "5" never persistent / "6" persistent ping-pong.  Default: 6 from 1024 (active) / 3072 (inactive) query rows on, 4 for a padded context pass, else 3.  (The one-item-per-CTA ping-pong and
// the eight-softmax-warp experiments of round 1 were measured slower everywhere and are gone: DESIGN.md §4.)
static int g_forced_fwd = -1;            // mimic_debug_force_fwd_kernel(); -1 = MIMIC_FWD_KERNEL (read once) or 0
void force_fwd_kernel(int k) { g_forced_fwd = k; }
static int forced_fwd_kernel() {
  static const int env = [] {
    const char* e = getenv("MIMIC_FWD_KERNEL");
    return e ? atoi(e) : 0;
  }();
  return g_forced_fwd >= 0 ? g_forced_fwd : env;
}
static bool env_flag(const char* name) { return getenv(name) != nullptr; }

int fwd(const AttnParams& p, cudaStream_t st) {
  static const bool no_persistent = env_flag("MIMIC_NO_PERSISTENT");
  const int forced = forced_fwd_kernel();
  // persistent 64-key variant applies: inactive handle (context pass), several tiles per CTA slot, AND the K/V of every
  // (kv head, batch) stream fit the L2 budget together — its CTAs in flight touch all streams at once
  const long long items64 = (long long)((p.T + 127) / 128) * p.H * p.B;
  const long long kv_bytes = 2LL * p.B * p.Hkv * p.S * p.d * 2;
  const bool p64_applies = !p.active && items64 >= 4LL * 2 * num_sms() && kv_bytes <= l2_budget_bytes() &&
                           (size_t)((p.S + 63) / 64) * 8 < 16000 && !no_persistent;
  // Ping-pong kernel from 2048 query rows on — except the padded context pass, where the persistent 64-key kernel is
  // ahead (measured at the Idefics2 shape, padding mask: 727 vs 654 TFLOP/s; without a mask 727 vs 755): the tile that
  // holds the padding takes the two-pass mixed path in every item of that batch row.
  const bool prefer_p64 = p.key_mask != nullptr && p64_applies;
  if (forced == 6 ? fwdp_supports(p, true) : (forced == 0 && !prefer_p64 && fwdp_supports(p, false))) return fwdp(p, st);
  const int kern = (forced == 0 || forced == 6) ? 3 : forced;
  if ((kern == 3 || kern == 4 || kern == 5) && (size_t)((p.S + 63) / 64) * 8 < 16000) {
    const bool persistent = kern == 4 || p64_applies;
    if (persistent && !no_persistent && kern != 5) return fwd64p(p, st);
    return fwd64(p, st);
  }
  AttnParams pd = p;
  static const int dbg_flags = [] { const char* e = getenv("MIMIC_DEBUG_FLAGS"); return e ? atoi(e) : 0; }();
  pd.debug_flags = dbg_flags;
  CUtensorMap tq, tk, tv;
  int rc;
  if ((rc = make_map(&tq, p.q, p.B, p.H, p.T, p.q_sb, p.q_sh, p.q_st, 128))) return rc;
  if ((rc = make_map(&tk, p.k, p.B, p.Hkv, p.S, p.k_sb, p.k_sh, p.k_ss, 128))) return rc;
  if ((rc = make_map(&tv, p.v, p.B, p.Hkv, p.S, p.v_sb, p.v_sh, p.v_ss, 128))) return rc;
  dim3 grid((p.T + BM - 1) / BM, p.H, p.B);
  const size_t smem3 = FwdSmem<3>::bytes(p.S), smem2 = FwdSmem<2>::bytes(p.S);
  static size_t configured3 = 0, configured2 = 0;
  if (smem3 <= MAX_DYN_SMEM) {
    if (smem3 > configured3) {
      MIMIC_CUDA_OK(cudaFuncSetAttribute(attn_fwd_tc_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem3));
      configured3 = smem3;
    }
    attn_fwd_tc_kernel<3><<<grid, FWD_THREADS, smem3, st>>>(tq, tk, tv, pd);
  } else {
    if (smem2 > configured2) {
      MIMIC_CUDA_OK(cudaFuncSetAttribute(attn_fwd_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
      configured2 = smem2;
    }
    attn_fwd_tc_kernel<2><<<grid, FWD_THREADS, smem2, st>>>(tq, tk, tv, pd);
  }
  MIMIC_LAUNCH_OK("attn_fwd_tc_kernel");
  return MIMIC_OK;
}

}  // namespace tc
}  // namespace mimic
