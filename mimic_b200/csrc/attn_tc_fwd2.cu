// Shifted self-attention forward, two query tiles per CTA ("ping-pong"): two softmax warpgroups
// each own one 128-row query tile, so the tensor pipe always has the other tile's MMAs to run while
// one warpgroup is in its softmax, and every K/V tile loaded by TMA feeds two QK^T and two PV MMAs.
// Same math, masks, statistics and outputs as attn_fwd_tc_kernel (attn_tc.cu); used when T > 128.
//
//   warps 0-3  softmax/correction/epilogue of query tile 0   (TMEM S0 [0,128)   O0 [256,384))
//   warps 4-7  softmax/correction/epilogue of query tile 1   (TMEM S1 [128,256) O1 [384,512))
//   warp  8    TMA producer (Q x2 once, K/V ring of 2 stages)      warp 9  MMA issuer
// MMA issue order:  QK0(0) QK1(0) | for j: [P0(j)] PV0(j) QK0(j+1)  [P1(j)] PV1(j) QK1(j+1)
// (P_x(j) aliases S_x, so QK_x(j+1) is issued after PV_x(j); tcgen05 ops of one thread run in order).
// Register budget: 384 threads -> 168 registers/thread; a full 128-column fp32 row lives in registers
// between the max and the exp passes (ptxas spills a few values; setmaxnreg re-balancing is future work).
#include <cuda.h>

#include "attn_common.cuh"
#include "tc_host.cuh"
#include "tc_ptx.cuh"
#include "tc_softmax.cuh"

namespace mimic {
namespace tc {

using namespace ptx;

namespace f2 {
constexpr int BM = 128, BN = 128, HD = 128, NQT = 2, NKV = 2;
constexpr int HALF_BYTES = 128 * 64 * 2;
constexpr int TILE_BYTES = 2 * HALF_BYTES;
constexpr int THREADS = 384;
constexpr float LOG2E = 1.4426950408889634f, LN2 = 0.6931471805599453f;
constexpr float RESCALE_THRESHOLD = 8.f;
constexpr int OFF_Q = 0, OFF_K = NQT * TILE_BYTES, OFF_V = OFF_K + NKV * TILE_BYTES, OFF_BAR = OFF_V + NKV * TILE_BYTES;
constexpr int OFF_TMEM = OFF_BAR + 24 * 8, OFF_W = OFF_TMEM + 16, OFF_SHIFT = OFF_W + 512, OFF_KBITS = OFF_SHIFT + 512;
// barriers
enum { Q_FULL = 0, K_FULL = 1, K_EMPTY = 3, V_FULL = 5, V_EMPTY = 7, S_FULL = 9, SOFT_DONE = 11, PV_DONE = 13, O_FULL = 15, N_BARS = 17 };
constexpr uint32_t IDESC_QK = idesc_bf16(BM, BN, 0, 0);
constexpr uint32_t IDESC_PV = idesc_bf16(BM, HD, 0, 1);
inline size_t smem_bytes(int S) { return OFF_KBITS + (size_t)((S + BN - 1) / BN) * 16; }
}  // namespace f2

__device__ __forceinline__ uint32_t low_bits2(int n) { return n >= 32 ? 0xFFFFFFFFu : (n <= 0 ? 0u : ((1u << n) - 1u)); }

template <int N>
__device__ __forceinline__ void setmaxnreg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N>
__device__ __forceinline__ void setmaxnreg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }

__global__ void __launch_bounds__(f2::THREADS, 1)
attn_fwd_tc2_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                    const __grid_constant__ CUtensorMap tm_v, AttnParams p) {
  using namespace f2;
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
  uint32_t* kbits = reinterpret_cast<uint32_t*>(smem + OFF_KBITS);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qp = gridDim.x - 1 - blockIdx.x;          // query-tile pair, heaviest first
  const int h = blockIdx.y, b = blockIdx.z;
  const int kvh = h / (p.H / p.Hkv);
  const int T = p.T, S = p.S, off = p.S - p.T;
  const int n_tiles = (S + BN - 1) / BN;
  // per query tile x: first row, last key tile with a visible entry (jd), number of key tiles visited (nx)
  int t0x[NQT], jdx[NQT], nx[NQT];
#pragma unroll
  for (int x = 0; x < NQT; ++x) {
    t0x[x] = (qp * NQT + x) * BM;
    if (t0x[x] >= T) {
      jdx[x] = -1;
      nx[x] = 0;
    } else {
      jdx[x] = n_tiles - 1;
      if (p.causal) {
        const int smax = min(t0x[x] + BM - 1, T - 1) + off;
        jdx[x] = smax < 0 ? -1 : min(n_tiles - 1, smax / BN);
      }
      nx[x] = p.active ? n_tiles : jdx[x] + 1;
    }
  }
  const int n = max(nx[0], nx[1]);
  const int jdmax = max(jdx[0], jdx[1]);

  if (warp == 8) {
    if (lane == 0) {
      mbar_init(&bars[Q_FULL], 1);
      for (int i = 0; i < 2; ++i) {
        mbar_init(&bars[K_FULL + i], 1);
        mbar_init(&bars[K_EMPTY + i], 1);
        mbar_init(&bars[V_FULL + i], 1);
        mbar_init(&bars[V_EMPTY + i], 1);
        mbar_init(&bars[S_FULL + i], 1);
        mbar_init(&bars[SOFT_DONE + i], 128);
        mbar_init(&bars[PV_DONE + i], 1);
        mbar_init(&bars[O_FULL + i], 1);
      }
      fence_barrier_init();
      prefetch_tmap(&tm_q);
      prefetch_tmap(&tm_k);
      prefetch_tmap(&tm_v);
    }
    __syncwarp();
    tmem_alloc(tmem_slot, 512);
  } else if (warp < 8) {
    const int st = threadIdx.x;                       // 0..255
    if (p.active && st < HD) {
      w_s[st] = p.w[h * HD + st];
      shift_s[st] = p.shift[h * HD + st];
    }
    const uint8_t* km = p.key_mask ? p.key_mask + (int64_t)b * S : nullptr;
    for (int wi = warp; wi < n_tiles * 4; wi += 8) {
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

  if (warp >= 8) {
    if (warp == 8) {
      // ===================== TMA producer =====================
      if (lane == 0) {
        mbar_expect_tx(&bars[Q_FULL], NQT * TILE_BYTES);
#pragma unroll
        for (int x = 0; x < NQT; ++x) {
          tma_load_4d(sQ + x * TILE_BYTES, &tm_q, &bars[Q_FULL], 0, t0x[x], h, b);       // rows >= T arrive as zeros
          tma_load_4d(sQ + x * TILE_BYTES + HALF_BYTES, &tm_q, &bars[Q_FULL], 64, t0x[x], h, b);
        }
        for (int j = 0; j < n; ++j) {
          const int st = j & 1;
          const uint32_t ph = (j >> 1) & 1;
          mbar_wait(&bars[K_EMPTY + st], ph ^ 1, 100 + st);
          mbar_expect_tx(&bars[K_FULL + st], TILE_BYTES);
          tma_load_4d(sK + st * TILE_BYTES, &tm_k, &bars[K_FULL + st], 0, j * BN, kvh, b);
          tma_load_4d(sK + st * TILE_BYTES + HALF_BYTES, &tm_k, &bars[K_FULL + st], 64, j * BN, kvh, b);
          if (j <= jdmax) {
            mbar_wait(&bars[V_EMPTY + st], ph ^ 1, 110 + st);
            mbar_expect_tx(&bars[V_FULL + st], TILE_BYTES);
            tma_load_4d(sV + st * TILE_BYTES, &tm_v, &bars[V_FULL + st], 0, j * BN, kvh, b);
            tma_load_4d(sV + st * TILE_BYTES + HALF_BYTES, &tm_v, &bars[V_FULL + st], 64, j * BN, kvh, b);
          }
        }
      }
    } else if (warp == 9 && n > 0) {
      // ===================== MMA issuer (warp-uniform control flow, one elected lane issues) =====================
      constexpr uint32_t HI = desc_hi_sw128(1024);
      const uint32_t q_lo0 = desc_lo(smem_u32(sQ), 16), k_lo0 = desc_lo(smem_u32(sK), 16);
      const uint32_t v_lo0 = desc_lo(smem_u32(sV), HALF_BYTES);
      mbar_wait(&bars[Q_FULL], 0, 200);
      // QK_x(j): S_x = Q_x K_j^T.  `last` = no later MMA reads this K stage -> release it.
      auto issue_qk = [&](int x, int j, bool last) {
        const int ks = j & 1;
        mbar_wait(&bars[K_FULL + ks], (j >> 1) & 1, 210 + ks);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t q_lo = q_lo0 + x * (TILE_BYTES >> 4), k_lo = k_lo0 + ks * (TILE_BYTES >> 4);
          const uint32_t d_s = tmem_base + x * BN;
#pragma unroll
          for (int k = 0; k < HD / 16; ++k) {
            const uint32_t koff = ((k >> 2) * HALF_BYTES + (k & 3) * 32) >> 4;
            if (k == 0) mma_ss_c<false>(d_s, desc64(HI, q_lo + koff), desc64(HI, k_lo + koff), IDESC_QK);
            else mma_ss_c<true>(d_s, desc64(HI, q_lo + koff), desc64(HI, k_lo + koff), IDESC_QK);
          }
          if (last) tc_commit(&bars[K_EMPTY + ks]);
          tc_commit(&bars[S_FULL + x]);
        }
        __syncwarp();
      };
      auto last_qk_user = [&](int x, int j) {          // is x the last query tile whose QK reads key tile j?
        for (int y = x + 1; y < NQT; ++y)
          if (j < nx[y]) return false;
        return true;
      };
#pragma unroll
      for (int x = 0; x < NQT; ++x)
        if (nx[x] > 0) issue_qk(x, 0, last_qk_user(x, 0));
      for (int j = 0; j < n; ++j) {
        const int vs = j & 1;
#pragma unroll
        for (int x = 0; x < NQT; ++x) {
          if (j >= nx[x]) continue;
          mbar_wait(&bars[SOFT_DONE + x], j & 1, 220 + x);
          tc_fence_after();
          if (j <= jdx[x]) {
            mbar_wait(&bars[V_FULL + vs], (j >> 1) & 1, 230 + vs);
            tc_fence_after();
            bool last_pv = true;
            for (int y = x + 1; y < NQT; ++y)
              if (j <= jdx[y] && j < nx[y]) last_pv = false;
            if (elect_one()) {
              const uint32_t v_lo = v_lo0 + vs * (TILE_BYTES >> 4);
              const uint32_t a_p = tmem_base + x * BN, d_o = tmem_base + 256 + x * BN;
              if (j > 0) {
#pragma unroll
                for (int k = 0; k < BN / 16; ++k) mma_ts_c<true>(d_o, a_p + k * 8, desc64(HI, v_lo + k * (2048 >> 4)), IDESC_PV);
              } else {
#pragma unroll
                for (int k = 0; k < BN / 16; ++k) {
                  if (k == 0) mma_ts_c<false>(d_o, a_p, desc64(HI, v_lo), IDESC_PV);
                  else mma_ts_c<true>(d_o, a_p + k * 8, desc64(HI, v_lo + k * (2048 >> 4)), IDESC_PV);
                }
              }
              if (last_pv) tc_commit(&bars[V_EMPTY + vs]);
              tc_commit(&bars[PV_DONE + x]);
            }
            __syncwarp();
          }
          if (j + 1 < nx[x]) issue_qk(x, j + 1, last_qk_user(x, j + 1));
        }
      }
      if (elect_one()) {
        tc_commit(&bars[O_FULL + 0]);
        tc_commit(&bars[O_FULL + 1]);
      }
      __syncwarp();
    }
  } else {
    // ===================== softmax / correction / epilogue warpgroup x: one query row per thread =====================
    const int x = warp >> 2;
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;
    const int t0 = t0x[x], t = t0 + r;
    const bool row_ok = t < T;
    const int jd = jdx[x], nloc = nx[x];
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    const uint32_t t_S = tmem_base + lane_off + x * BN, t_O = tmem_base + lane_off + 256 + x * BN;
    const float scale2 = p.scale * LOG2E;
    const float2 scale2v = make_float2(scale2, scale2);
    const uint8_t* sQx = sQ + x * TILE_BYTES;

    float z1 = 0.f;
    if (nloc > 0 || t0 < T) {
      mbar_wait(&bars[Q_FULL], 0, 300);
      if (p.active) {
#pragma unroll
        for (int hf = 0; hf < 2; ++hf)
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const uint4 u = *reinterpret_cast<const uint4*>(sQx + hf * HALF_BYTES + r * 128 + ((c ^ (r & 7)) << 4));
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
    for (int j = 0; j < nloc; ++j) {
      mbar_wait(&bars[S_FULL + x], j & 1, 310 + x);
      tc_fence_after();
      const int key0 = j * BN;
      uint32_t vw[4], xw[4];
      bool all_vis_t = row_ok, any_exc_t = false;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const uint32_t inr = low_bits2(S - key0 - 32 * i);
        const uint32_t cz = p.causal ? low_bits2(t + off - key0 - 32 * i + 1) : 0xFFFFFFFFu;
        vw[i] = row_ok ? (kbits[j * 4 + i] & cz) : 0u;
        xw[i] = (p.active && row_ok) ? (inr & ~vw[i]) : 0u;
        all_vis_t = all_vis_t && (vw[i] == 0xFFFFFFFFu);
        any_exc_t = any_exc_t || (xw[i] != 0u);
      }
      const bool all_vis = __all_sync(0xffffffffu, all_vis_t);     // warp-uniform: mask-free fast path
      const bool any_exc = __any_sync(0xffffffffu, any_exc_t);
      if (j <= jd) {
        float s[128];
        tmem_ld32(t_S, *reinterpret_cast<float(*)[32]>(s));
        tmem_ld32(t_S + 32, *reinterpret_cast<float(*)[32]>(s + 32));
        tmem_ld32(t_S + 64, *reinterpret_cast<float(*)[32]>(s + 64));
        tmem_ld32(t_S + 96, *reinterpret_cast<float(*)[32]>(s + 96));
        tc_wait_ld();
        float mv = -INFINITY;
        if (all_vis) {
          mv = row_max128(s);
        } else {
#pragma unroll
          for (int c = 0; c < 128; ++c)
            if ((vw[c >> 5] >> (c & 31)) & 1u) mv = fmaxf(mv, s[c]);
        }
        mv *= scale2;
        const float m_new = fmaxf(m_c, mv);
        const bool need = (m_c != -INFINITY) && (m_new > m_c + RESCALE_THRESHOLD);
        if (m_c == -INFINITY) m_c = m_new;
        if (__any_sync(0xffffffffu, need)) {            // lazy rescale of O and l_c (warp-collective TMEM ops)
          const float alpha = need ? ex2(m_c - m_new) : 1.f;
          if (j > 0) {
            mbar_wait(&bars[PV_DONE + x], (j - 1) & 1, 320 + x);
            tc_fence_after();
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              float o[32];
              tmem_ld32(t_O + 32 * i, o);
              tc_wait_ld();
              uint32_t ou[32];
#pragma unroll
              for (int c = 0; c < 32; ++c) ou[c] = __float_as_uint(o[c] * alpha);
              tmem_st32(t_O + 32 * i, ou);
            }
          }
          if (need) {
            l_c *= alpha;
            m_c = m_new;
          }
        }
        const float neg_m = (m_c == -INFINITY) ? 0.f : -m_c;
        if (all_vis) {
          l_c += exp_pack_store128(s, scale2, neg_m, t_S);      // packed fp32x2 FMA/ADD, 4 independent sum chains
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
            tmem_st16(t_S + 16 * i, pk);
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
        const bool all_exc = __all_sync(0xffffffffu, xw[0] == 0xFFFFFFFFu && xw[1] == 0xFFFFFFFFu &&
                                                         xw[2] == 0xFFFFFFFFu && xw[3] == 0xFFFFFFFFu);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float s[32];
          tmem_ld32(t_S + 32 * i, s);
          tc_wait_ld();
          if (all_exc) {
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
      mbar_arrive(&bars[SOFT_DONE + x]);
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
      mbar_wait(&bars[O_FULL + x], 0, 330 + x);
      tc_fence_after();
    }
    __nv_bfloat16* orow = reinterpret_cast<__nv_bfloat16*>(p.out) + (int64_t)b * p.o_sb + (int64_t)t * p.o_st + (int64_t)h * p.o_sh;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float o[32];
      if (jd >= 0) {
        tmem_ld32(t_O + 32 * i, o);
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
  if (warp == 8) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

int fwd2(const AttnParams& p, cudaStream_t st) {
  using namespace f2;
  CUtensorMap tq, tk, tv;
  int rc;
  if ((rc = make_map(&tq, p.q, p.B, p.H, p.T, p.q_sb, p.q_sh, p.q_st, 128))) return rc;
  if ((rc = make_map(&tk, p.k, p.B, p.Hkv, p.S, p.k_sb, p.k_sh, p.k_ss, 128))) return rc;
  if ((rc = make_map(&tv, p.v, p.B, p.Hkv, p.S, p.v_sb, p.v_sh, p.v_ss, 128))) return rc;
  const size_t smem = smem_bytes(p.S);
  static size_t configured = 0;
  if (smem > configured) {
    MIMIC_CUDA_OK(cudaFuncSetAttribute(attn_fwd_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  dim3 grid((p.T + NQT * BM - 1) / (NQT * BM), p.H, p.B);
  attn_fwd_tc2_kernel<<<grid, THREADS, smem, st>>>(tq, tk, tv, p);
  MIMIC_LAUNCH_OK("attn_fwd_tc2_kernel");
  return MIMIC_OK;
}

}  // namespace tc
}  // namespace mimic
