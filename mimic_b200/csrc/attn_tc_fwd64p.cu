// Shifted self-attention forward, 64-key tiles, TWO PERSISTENT CTAs PER SM (static work lists).
// Same math / masks / outputs / pipeline as attn_fwd_tc64_kernel (attn_tc_fwd64.cu).  The grid is 2 x #SMs CTAs; CTA c
// walks the work items c, c + G, c + 2G, ... of a list ordered by descending work (query tiles from the last to the
// first, (head, batch) fastest), i.e. a round-robin of a longest-first list.  All mbarrier phases and the K/V rings run
// on across items, so the next item's Q / K / V loads and first QK^T MMAs are in flight while the softmax warps still
// normalise and store the previous item, and the per-CTA costs measured on the one-tile-per-CTA kernel
// (scripts/trace_fwd.py: ~5500 cycles of block scheduling between CTAs, ~2100 of barrier init / TMEM alloc / key-bit
// table / sync) are paid once per CTA instead of once per tile.
// Used when the K/V of all (kv head, batch) streams fit L2 together (GQA, short contexts): every CTA in flight works on
// the same query-tile rank of a different (head, batch), so the streams are all live at once.  Measured +9 % at the
// Idefics2 context-pass shape and -17 % at T = 8192 with 64 MHA streams (268 MB of K/V > L2) — fwd() picks per shape.
#include <cuda.h>

#include "attn_common.cuh"
#include "tc_host.cuh"
#include "tc_ptx.cuh"
#include "tc_softmax.cuh"

namespace mimic {
namespace tc {

using namespace ptx;

namespace f64p {
constexpr int BM = 128, BN = 64, HD = 128, NKV = 2;
constexpr int Q_HALF = 128 * 128;            // [128 rows][128 B]
constexpr int Q_TILE = 2 * Q_HALF;           // 32 KB
constexpr int KV_HALF = BN * 128;            // [64 keys][128 B] = 8 KB
constexpr int KV_TILE = 2 * KV_HALF;         // 16 KB
constexpr int THREADS = 192;
constexpr float LOG2E = 1.4426950408889634f, LN2 = 0.6931471805599453f;
constexpr float RESCALE_THRESHOLD = 8.f;
constexpr int OFF_Q = 0, OFF_K = Q_TILE, OFF_V = OFF_K + NKV * KV_TILE, OFF_BAR = OFF_V + NKV * KV_TILE;
constexpr int OFF_TMEM = OFF_BAR + 24 * 8, OFF_W = OFF_TMEM + 16, OFF_SHIFT = OFF_W + 512, OFF_KBITS = OFF_SHIFT + 512;
enum { Q_FULL = 0, K_FULL = 1, K_EMPTY = 3, V_FULL = 5, V_EMPTY = 7, S_FULL = 9, SOFT_DONE = 11, PV_DONE = 13, O_FULL = 14, Q_EMPTY = 15 };
constexpr uint32_t IDESC_QK = idesc_bf16(BM, BN, 0, 0);
constexpr uint32_t IDESC_PV = idesc_bf16(BM, HD, 0, 1);
constexpr uint32_t TMEM_COLS = 256, TM_O = 128;
// key-validity bit tables: one per batch row when they all fit 12 KB (built once per CTA), else one rebuilt per item
inline int kbit_rows(int B, int S) { return (size_t)B * ((S + BN - 1) / BN) * 8 <= 12288 ? B : 1; }
inline size_t smem_bytes(int B, int S) { return OFF_KBITS + (size_t)kbit_rows(B, S) * ((S + BN - 1) / BN) * 8; }
}  // namespace f64p

__device__ __forceinline__ uint32_t low_bits3p(int n) { return n >= 32 ? 0xFFFFFFFFu : (n <= 0 ? 0u : ((1u << n) - 1u)); }

__global__ void __launch_bounds__(f64p::THREADS, 2)
attn_fwd_tc64p_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                     const __grid_constant__ CUtensorMap tm_v, AttnParams p) {
  using namespace f64p;
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
  uint32_t* kbits_all = reinterpret_cast<uint32_t*>(smem + OFF_KBITS);
  const bool kbits_per_row = (size_t)p.B * ((p.S + BN - 1) / BN) * 8 <= 12288;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int T = p.T, S = p.S, off = p.S - p.T;
  const int n_tiles = (S + BN - 1) / BN;
  const int n_qt = (T + BM - 1) / BM;
  const int HB = p.H * p.B;
  const int n_items = n_qt * HB;
  const int G = p.H / p.Hkv;
  // item -> (query tile, head, batch): query-tile rank slowest (descending work under the causal mask), heads fastest
  // (a GQA group shares its K/V in L2): the round-robin hands every CTA a descending sequence of work
  struct Item {
    int qt, h, b, kvh, t0, jd, n;
  };
  auto decode = [&](int it) {
    Item m;
    const int qrank = it / HB, hb = it - qrank * HB;
    m.qt = n_qt - 1 - qrank;
    m.b = hb / p.H;
    m.h = hb - m.b * p.H;
    m.kvh = m.h / G;
    m.t0 = m.qt * BM;
    m.jd = n_tiles - 1;
    if (p.causal) {
      const int smax = min(m.t0 + BM - 1, T - 1) + off;
      m.jd = smax < 0 ? -1 : min(n_tiles - 1, smax / BN);
    }
    m.n = p.active ? n_tiles : m.jd + 1;
    return m;
  };

  if (warp == 0) {
    if (lane == 0) {
      mbar_init(&bars[Q_FULL], 1);
      for (int i = 0; i < 2; ++i) {
        mbar_init(&bars[K_FULL + i], 1);
        mbar_init(&bars[K_EMPTY + i], 1);
        mbar_init(&bars[V_FULL + i], 1);
        mbar_init(&bars[V_EMPTY + i], 1);
        mbar_init(&bars[S_FULL + i], 1);
        mbar_init(&bars[SOFT_DONE + i], 128);
      }
      mbar_init(&bars[PV_DONE], 1);
      mbar_init(&bars[O_FULL], 1);
      mbar_init(&bars[Q_EMPTY], 1 + 128);          // last QK^T MMA of the item (commit) + every softmax thread (z1 read)
      fence_barrier_init();
      prefetch_tmap(&tm_q);
      prefetch_tmap(&tm_k);
      prefetch_tmap(&tm_v);
    }
    __syncwarp();
    tmem_alloc(tmem_slot, TMEM_COLS);
  } else if (warp >= 2 && kbits_per_row) {
    // every batch row's key bits, once per CTA: 16 independent byte loads in flight per lane
    const int words = n_tiles * 2;
    for (int bb = 0; bb < p.B; ++bb) {
      const uint8_t* km = p.key_mask ? p.key_mask + (int64_t)bb * S : nullptr;
      for (int w0 = warp - 2; w0 < words; w0 += 4 * 16) {
        uint8_t ok[16];
#pragma unroll
        for (int u = 0; u < 16; ++u) {
          const int s = (w0 + 4 * u) * 32 + lane;
          ok[u] = (s < S) ? (km == nullptr ? (uint8_t)1 : km[s]) : (uint8_t)0;
        }
#pragma unroll
        for (int u = 0; u < 16; ++u) {
          const int wi = w0 + 4 * u;
          const uint32_t word = __ballot_sync(0xffffffffu, ok[u] != 0);
          if (lane == 0 && wi < words) kbits_all[bb * words + wi] = word;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_O = tmem_base + TM_O;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      uint32_t gk = 0, gv = 0;                           // K / V ring counters, running across items
      int k_item = 0;
      for (int it = blockIdx.x; it < n_items; it += gridDim.x, ++k_item) {
        mbar_wait(&bars[Q_EMPTY], (k_item & 1) ^ 1, 90);   // previous item's QK^T MMAs and z1 reads are done with sQ
        const Item m = decode(it);
        mbar_expect_tx(&bars[Q_FULL], Q_TILE);
        tma_load_4d(sQ, &tm_q, &bars[Q_FULL], 0, m.t0, m.h, m.b);
        tma_load_4d(sQ + Q_HALF, &tm_q, &bars[Q_FULL], 64, m.t0, m.h, m.b);
        for (int j = 0; j < m.n; ++j) {
          const int st = gk & 1;
          mbar_wait(&bars[K_EMPTY + st], ((gk >> 1) & 1) ^ 1, 100 + st);
          mbar_expect_tx(&bars[K_FULL + st], KV_TILE);
          tma_load_4d(sK + st * KV_TILE, &tm_k, &bars[K_FULL + st], 0, j * BN, m.kvh, m.b);
          tma_load_4d(sK + st * KV_TILE + KV_HALF, &tm_k, &bars[K_FULL + st], 64, j * BN, m.kvh, m.b);
          ++gk;
          if (j <= m.jd) {
            const int sv = gv & 1;
            mbar_wait(&bars[V_EMPTY + sv], ((gv >> 1) & 1) ^ 1, 110 + sv);
            mbar_expect_tx(&bars[V_FULL + sv], KV_TILE);
            tma_load_4d(sV + sv * KV_TILE, &tm_v, &bars[V_FULL + sv], 0, j * BN, m.kvh, m.b);
            tma_load_4d(sV + sv * KV_TILE + KV_HALF, &tm_v, &bars[V_FULL + sv], 64, j * BN, m.kvh, m.b);
            ++gv;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (warp-uniform control flow, one elected lane issues) =====================
    constexpr uint32_t HI = desc_hi_sw128(1024);
    const uint32_t q_lo = desc_lo(smem_u32(sQ), 16), k_lo0 = desc_lo(smem_u32(sK), 16);
    const uint32_t v_lo0 = desc_lo(smem_u32(sV), KV_HALF);            // V as MN-major B: d halves 8 KB apart
    uint32_t gq = 0, gs = 0, gv = 0;                     // QK^T issued / softmax steps consumed / PV issued, across items
    int k_item = 0;
    auto issue_qk = [&](bool last) {                     // S[gq & 1] = Q K_gq^T; the item's last one also releases sQ
      const int st = gq & 1;
      mbar_wait(&bars[K_FULL + st], (gq >> 1) & 1, 210 + st);
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
        if (last) tc_commit(&bars[Q_EMPTY]);
      }
      __syncwarp();
      ++gq;
    };
    for (int it = blockIdx.x; it < n_items; it += gridDim.x, ++k_item) {
      const Item m = decode(it);
      const int n = m.n, jd = m.jd;
      if (n == 0) {                                      // no key in range: nothing to issue, but sQ must still be released
        if (elect_one()) mbar_arrive(&bars[Q_EMPTY]);
        __syncwarp();
        continue;
      }
      mbar_wait(&bars[Q_FULL], k_item & 1, 200);
      issue_qk(n == 1);
      if (n > 1) issue_qk(n == 2);
      for (int j = 0; j < n; ++j) {
        const int st = gs & 1;
        mbar_wait(&bars[SOFT_DONE + st], (gs >> 1) & 1, 220 + st);
        tc_fence_after();
        if (j <= jd) {
          const int sv = gv & 1;
          mbar_wait(&bars[V_FULL + sv], (gv >> 1) & 1, 230 + sv);
          tc_fence_after();
          if (elect_one()) {
            // A: P_j in TMEM (64 keys = 32 packed columns, 8 per K step); B: V_j [64 keys][d], 16 keys = 2 KB per step
            const uint32_t v_lo = v_lo0 + sv * (KV_TILE >> 4);
            const uint32_t a_p = tmem_base + st * BN;
#pragma unroll
            for (int k = 0; k < BN / 16; ++k)
              mma_ts(tmem_O, a_p + k * 8, desc64(HI, v_lo + k * (2048 >> 4)), IDESC_PV, (k > 0 || j > 0) ? 1u : 0u);
            tc_commit(&bars[V_EMPTY + sv]);
            tc_commit(&bars[PV_DONE]);
          }
          __syncwarp();
          ++gv;
        }
        ++gs;
        if (j + 2 < n) issue_qk(j + 3 == n);
      }
      if (jd >= 0) {
        if (elect_one()) tc_commit(&bars[O_FULL]);
        __syncwarp();
      }
    }
  } else {
    // ===================== softmax / correction / epilogue: one query row per thread =====================
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    const float scale2 = p.scale * LOG2E;
    const int sid = threadIdx.x - 64;
    uint32_t gs = 0, pv_seen = 0, o_seen = 0;            // softmax steps / PV steps / items with an O tile, across items
    int k_item = 0, last_b = -1;
    for (int it = blockIdx.x; it < n_items; it += gridDim.x, ++k_item) {
    const Item m = decode(it);
    const int n = m.n, jd = m.jd, h = m.h, b = m.b, t0 = m.t0;
    const int t = t0 + r;
    const bool row_ok = t < T;

    // per-item tables: w / shift of this head, key bits of this batch row (rebuilt only when the batch row changes)
    const uint32_t* kbits = kbits_all + (kbits_per_row ? b * n_tiles * 2 : 0);
    const bool rebuild = !kbits_per_row && b != last_b;
    if (p.active || rebuild) {
      named_bar_sync(1, 128);                            // every thread is done with the previous item's tables
      if (p.active) {
        w_s[sid] = p.w[h * HD + sid];
        shift_s[sid] = p.shift[h * HD + sid];
      }
      if (rebuild) {
        const uint8_t* km = p.key_mask ? p.key_mask + (int64_t)b * S : nullptr;
        for (int w0 = warp - 2; w0 < n_tiles * 2; w0 += 4 * 16) {      // 16 independent byte loads in flight per lane
          uint8_t ok[16];
#pragma unroll
          for (int u = 0; u < 16; ++u) {
            const int s = (w0 + 4 * u) * 32 + lane;
            ok[u] = (s < S) ? (km == nullptr ? (uint8_t)1 : km[s]) : (uint8_t)0;
          }
#pragma unroll
          for (int u = 0; u < 16; ++u) {
            const int wi = w0 + 4 * u;
            const uint32_t word = __ballot_sync(0xffffffffu, ok[u] != 0);
            if (lane == 0 && wi < n_tiles * 2) kbits_all[wi] = word;
          }
        }
        last_b = b;
      }
      named_bar_sync(1, 128);
    }

    float z1 = 0.f;
    mbar_wait(&bars[Q_FULL], k_item & 1, 300);
    if (p.active) {
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
    mbar_arrive(&bars[Q_EMPTY]);                         // this thread no longer reads sQ

    float m_c = -INFINITY, l_c = 0.f, m_x = -INFINITY, l_x = 0.f;   // log2 domain
    for (int j = 0; j < n; ++j, ++gs) {
      const int st = gs & 1;
      mbar_wait(&bars[S_FULL + st], (gs >> 1) & 1, 310 + st);
      tc_fence_after();
      const uint32_t taddr = tmem_base + lane_off + st * BN;
      const int key0 = j * BN;
      uint32_t vw[2], xw[2];
      bool all_vis_t = row_ok, any_exc_t = false, all_exc_t = true;
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const uint32_t inr = low_bits3p(S - key0 - 32 * i);
        const uint32_t cz = p.causal ? low_bits3p(t + off - key0 - 32 * i + 1) : 0xFFFFFFFFu;
        vw[i] = row_ok ? (kbits[j * 2 + i] & cz) : 0u;
        xw[i] = (p.active && row_ok) ? (inr & ~vw[i]) : 0u;
        all_vis_t = all_vis_t && (vw[i] == 0xFFFFFFFFu);
        any_exc_t = any_exc_t || (xw[i] != 0u);
        all_exc_t = all_exc_t && (xw[i] == 0xFFFFFFFFu);
      }
      const bool all_vis = __all_sync(0xffffffffu, all_vis_t);
      const bool any_exc = __any_sync(0xffffffffu, any_exc_t);
      float s[64];
      tmem_ld32(taddr, *reinterpret_cast<float(*)[32]>(s));
      tmem_ld32(taddr + 32, *reinterpret_cast<float(*)[32]>(s + 32));
      tc_wait_ld();
      if (j <= jd) {
        float mv = -INFINITY;
        if (all_vis) {
          mv = row_max64(s);
        } else {
#pragma unroll
          for (int c = 0; c < 64; ++c)
            if ((vw[c >> 5] >> (c & 31)) & 1u) mv = fmaxf(mv, s[c]);
        }
        mv *= scale2;
        const float m_new = fmaxf(m_c, mv);
        const bool need = (m_c != -INFINITY) && (m_new > m_c + RESCALE_THRESHOLD);
        if (m_c == -INFINITY) m_c = m_new;
        if (__any_sync(0xffffffffu, need)) {            // lazy rescale of O and l_c (warp-collective TMEM ops)
          const float alpha = need ? ex2(m_c - m_new) : 1.f;
          if (j > 0) {
            mbar_wait(&bars[PV_DONE], (pv_seen - 1) & 1, 320);   // PV_{j-1} of this item = PV number pv_seen - 1
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
        const float neg_m = (m_c == -INFINITY) ? 0.f : -m_c;
        if (all_vis) {
          l_c += exp_pack_store64(s, scale2, neg_m, taddr);
        } else {
          float rowsum = 0.f;
#pragma unroll
          for (int i = 0; i < 2; ++i) {
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
          for (int c = 0; c < 64; ++c)
            if ((xw[c >> 5] >> (c & 31)) & 1u) mx = fmaxf(mx, s[c]);
          mx *= scale2;
          const float mx_new = fmaxf(m_x, mx);
          float es = 0.f;
#pragma unroll
          for (int c = 0; c < 64; ++c)
            if ((xw[c >> 5] >> (c & 31)) & 1u) es += ex2(fmaf(s[c], scale2, -mx_new));
          l_x = l_x * ((m_x == -INFINITY) ? 0.f : ex2(m_x - mx_new)) + es;
          m_x = mx_new;
        }
        tc_wait_st();
        ++pv_seen;
      } else if (__all_sync(0xffffffffu, all_exc_t)) {
        // tile above the causal diagonal, every key excluded: row-sum of exp only
        const float mx_new = fmaxf(m_x, row_max64(s) * scale2);
        l_x = l_x * ((m_x == -INFINITY) ? 0.f : ex2(m_x - mx_new)) +
              expsum32(*reinterpret_cast<const float(*)[32]>(s), scale2, -mx_new) +
              expsum32(*reinterpret_cast<const float(*)[32]>(s + 32), scale2, -mx_new);
        m_x = mx_new;
      } else if (any_exc) {
        float mx = -INFINITY;
#pragma unroll
        for (int c = 0; c < 64; ++c)
          if ((xw[c >> 5] >> (c & 31)) & 1u) mx = fmaxf(mx, s[c]);
        mx *= scale2;
        const float mx_new = fmaxf(m_x, mx);
        float es = 0.f;
#pragma unroll
        for (int c = 0; c < 64; ++c) {
          const float e = ex2(fmaf(s[c], scale2, -mx_new));
          es += ((xw[c >> 5] >> (c & 31)) & 1u) ? e : 0.f;
        }
        l_x = l_x * ((m_x == -INFINITY) ? 0.f : ex2(m_x - mx_new)) + es;
        m_x = mx_new;
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
      mbar_wait(&bars[O_FULL], o_seen & 1, 330);
      tc_fence_after();
      ++o_seen;
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
    tc_fence_before();      // O reads above are ordered before the next item's first SOFT_DONE arrive -> PV overwrite
    }                       // item loop
  }
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

int fwd64p(const AttnParams& p, cudaStream_t st) {
  using namespace f64p;
  CUtensorMap tq, tk, tv;
  int rc;
  if ((rc = make_map(&tq, p.q, p.B, p.H, p.T, p.q_sb, p.q_sh, p.q_st, BM))) return rc;
  if ((rc = make_map(&tk, p.k, p.B, p.Hkv, p.S, p.k_sb, p.k_sh, p.k_ss, BN))) return rc;
  if ((rc = make_map(&tv, p.v, p.B, p.Hkv, p.S, p.v_sb, p.v_sh, p.v_ss, BN))) return rc;
  const size_t smem = smem_bytes(p.B, p.S);
  static size_t configured = 0;
  if (smem > configured) {
    MIMIC_CUDA_OK(cudaFuncSetAttribute(attn_fwd_tc64p_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  static int n_sm = 0;
  if (n_sm == 0) {
    int dev = 0;
    MIMIC_CUDA_OK(cudaGetDevice(&dev));
    MIMIC_CUDA_OK(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
  }
  const long long n_items = (long long)((p.T + BM - 1) / BM) * p.H * p.B;
  const int grid = (int)(n_items < 2LL * n_sm ? n_items : 2LL * n_sm);   // two resident CTAs per SM, each loops over its items
  attn_fwd_tc64p_kernel<<<grid, THREADS, smem, st>>>(tq, tk, tv, p);
  MIMIC_LAUNCH_OK("attn_fwd_tc64p_kernel");
  return MIMIC_OK;
}

}  // namespace tc
}  // namespace mimic
