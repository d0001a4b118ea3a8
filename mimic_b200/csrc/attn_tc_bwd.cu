// Shifted self-attention BACKWARD on Blackwell tensor cores (bf16, head_dim 128).
// Math: SURVEY.md Appendix A (backward of SDPA + do_shift, src/shift_encoder.py:352-359, 682-732).
//
// One CTA per (128-key tile, kv head, batch); it keeps K and V resident in shared memory, walks the
// query heads of the GQA group and their 64-row query tiles, and accumulates dK, dV in TMEM:
//
//   S^T  = K Q_i^T      SS mma, M=128 keys, N=64 queries   -> TMEM SdQ[i&1]       (keys on TMEM lanes)
//   dP^T = V dO_i^T     SS mma                              -> TMEM dP[i&1]        (causal tiles only)
//   softmax warps (one key row per thread, 64 query columns):
//        P^T  = exp2(s - lse_c[t])            visible entries
//        dS^T = P^T (dP^T - Delta'[t])        visible;   -dz[t] exp2(s - lse_full[t])   excluded (future/masked)
//        -> bf16 P^T, dS^T back into TMEM (A operands) and dS^T into smem ([keys][queries], B operand)
//   dV  += P^T dO_i     TS mma (A from TMEM, B = dO_i MN-major)           (causal tiles only)
//   dK  += dS^T Q_i     TS mma   (dS carries no softmax scale: dK x scale in the epilogue, dQ x scale at conversion)
//   dQ^T = K^T dS^T     SS mma, M=128 (d), N=64 queries (A = K MN-major, B = dS^T MN-major) -> TMEM SdQ[i&1]
//   drain warps: dQ^T TMEM -> registers -> fp32 smem [64][128] -> ONE bulk reduce-add
//        (cp.reduce.async.bulk .add.f32) into the fp32 dq accumulator — no per-element atomics.
// Delta' folds the visible part of the shift term: Delta' = g.SA + dz exp(lse_c - lse_full) (prep kernel).
// TMEM map (512 cols): dV [0,128) dK [128,256) SdQ0 [256,320) SdQ1 [320,384) dP0 [384,448) dP1 [448,512).
#include <cuda.h>

#include <cstdlib>

#include "attn_common.cuh"
#include "tc_host.cuh"
#include "tc_ptx.cuh"

namespace mimic {
namespace tc {

using namespace ptx;

namespace bw {
constexpr int BN = 128;                    // keys per CTA
constexpr int BQ = 64;                     // queries per iteration
constexpr int HD = 128;
constexpr int KV_HALF = 128 * 128;         // [128 keys][128 B]
constexpr int KV_TILE = 2 * KV_HALF;       // 32 KB
constexpr int Q_HALF = BQ * 128;           // [64 rows][128 B] = 8 KB
constexpr int Q_TILE = 2 * Q_HALF;         // 16 KB
constexpr int DS_TILE = 128 * 128;         // dS^T [128 keys][64 queries] bf16 = 16 KB
constexpr int STAGE_BYTES = 32 * HD * 4;   // fp32 dQ staging for 32 query rows, 16 KB (two bulk reduces per tile)
constexpr int NQS = 3;                     // Q / dO TMA stages
constexpr int THREADS = 448;                // warp 0 TMA, warp 1 MMA, warps 2-9 softmax (2 per TMEM quarter), warps 10-13 drain
constexpr float LOG2E = 1.4426950408889634f;

constexpr int OFF_K = 0, OFF_V = KV_TILE, OFF_Q = 2 * KV_TILE, OFF_DO = OFF_Q + NQS * Q_TILE, OFF_DS = OFF_DO + NQS * Q_TILE,
              OFF_STAGE = OFF_DS + 2 * DS_TILE, OFF_STATS = OFF_STAGE + STAGE_BYTES, OFF_BAR = OFF_STATS + 2 * BQ * 16,
              OFF_TMEM = OFF_BAR + 24 * 8, SMEM_BYTES = OFF_TMEM + 16;

enum Bar { KV_FULL = 0, QDO_FULL = 1, QDO_EMPTY = 4, S_FULL = 7, SOFT_DONE = 9, DQ_FULL = 11, DRAIN_DONE = 13, DKV_FULL = 15, N_BARS = 16 };

constexpr uint32_t TM_DV = 0, TM_DK = 128, TM_SDQ = 256, TM_DP = 384;
constexpr uint32_t IDESC_ST = idesc_bf16(128, BQ, 0, 0);     // S^T, dP^T : A, B K-major
constexpr uint32_t IDESC_DKV = idesc_bf16(128, HD, 0, 1);    // dV, dK    : A in TMEM, B MN-major
constexpr uint32_t IDESC_DQT = idesc_bf16(128, BQ, 1, 1);    // dQ^T      : A = K MN-major, B = dS^T MN-major
}  // namespace bw

__device__ __forceinline__ uint32_t low_bits_b(int n) { return n >= 32 ? 0xFFFFFFFFu : (n <= 0 ? 0u : ((1u << n) - 1u)); }

__global__ void __launch_bounds__(bw::THREADS, 1)
attn_bwd_tc_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                   const __grid_constant__ CUtensorMap tm_v, const __grid_constant__ CUtensorMap tm_do, AttnParams p) {
  using namespace bw;
  // address the dynamic window directly (no pointer rounding through integers): keeps every access LDS/STS
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) {
    if (threadIdx.x == 0) printf("[mimic tcgen05] dynamic shared memory is not 1024-byte aligned\n");
    __trap();
  }
  uint8_t* sK = smem + OFF_K;
  uint8_t* sV = smem + OFF_V;
  uint8_t* sQ = smem + OFF_Q;
  uint8_t* sDO = smem + OFF_DO;
  uint8_t* sDS = smem + OFF_DS;
  float* stage = reinterpret_cast<float*>(smem + OFF_STAGE);
  float4* stats = reinterpret_cast<float4*>(smem + OFF_STATS);      // [2][64] (lse_c*log2e, Delta', dz, lse_full*log2e)
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + OFF_TMEM);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kt = blockIdx.x, kvh = blockIdx.y, b = blockIdx.z;
  // timeline trace (debug): role r writes record k of iteration i at trace[(i * 8 + r) * 8 + k]
  const bool tracing = p.trace != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0;
  auto stamp = [&](int role, int i, int k) {
    if (tracing && i < 24) p.trace[(i * 8 + role) * 8 + k] = clock64();
  };
  const int G = p.H / p.Hkv;
  const int T = p.T, S = p.S, off = p.S - p.T;
  const int s0 = kt * BN;
  const int n_qt = (T + BQ - 1) / BQ;
  int qt_first = 0;
  if (!p.active && p.causal) {
    const int tmin = s0 - off;                         // first query row that can see key s0
    qt_first = tmin <= 0 ? 0 : tmin / BQ;
  }
  const int nq = n_qt - qt_first;                      // query tiles per head (>= 1 for every key tile in range)
  const int n = nq > 0 ? G * nq : 0;                   // iterations
  auto need_c = [&](int qt) { return !p.causal || (min(qt * BQ + BQ - 1, T - 1) + off >= s0); };

  if (warp == 0) {
    if (lane == 0) {
      mbar_init(&bars[KV_FULL], 1);
      for (int i = 0; i < NQS; ++i) {
        mbar_init(&bars[QDO_FULL + i], 1);
        mbar_init(&bars[QDO_EMPTY + i], 1);
      }
      for (int i = 0; i < 2; ++i) {
        mbar_init(&bars[S_FULL + i], 1);
        mbar_init(&bars[SOFT_DONE + i], 256);
        mbar_init(&bars[DQ_FULL + i], 1);
        mbar_init(&bars[DRAIN_DONE + i], 128);
      }
      mbar_init(&bars[DKV_FULL], 1);
      fence_barrier_init();
      prefetch_tmap(&tm_q);
      prefetch_tmap(&tm_k);
      prefetch_tmap(&tm_v);
      prefetch_tmap(&tm_do);
    }
    __syncwarp();
    tmem_alloc(tmem_slot, 512);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      mbar_expect_tx(&bars[KV_FULL], 2 * KV_TILE);
      tma_load_4d(sK, &tm_k, &bars[KV_FULL], 0, s0, kvh, b);
      tma_load_4d(sK + KV_HALF, &tm_k, &bars[KV_FULL], 64, s0, kvh, b);
      tma_load_4d(sV, &tm_v, &bars[KV_FULL], 0, s0, kvh, b);
      tma_load_4d(sV + KV_HALF, &tm_v, &bars[KV_FULL], 64, s0, kvh, b);
      for (int i = 0; i < n; ++i) {
        const int st = i % NQS;
        const int h = kvh * G + i / nq, qt = qt_first + i % nq;
        const bool nc = need_c(qt);
        if (!(p.debug_flags & 8)) mbar_wait(&bars[QDO_EMPTY + st], ((i / NQS) & 1) ^ 1, 100 + st);
        stamp(0, i, 0);
        mbar_expect_tx(&bars[QDO_FULL + st], nc ? 2 * Q_TILE : Q_TILE);
        tma_load_4d(sQ + st * Q_TILE, &tm_q, &bars[QDO_FULL + st], 0, qt * BQ, h, b);
        tma_load_4d(sQ + st * Q_TILE + Q_HALF, &tm_q, &bars[QDO_FULL + st], 64, qt * BQ, h, b);
        if (nc) {
          tma_load_4d(sDO + st * Q_TILE, &tm_do, &bars[QDO_FULL + st], 0, qt * BQ, h, b);
          tma_load_4d(sDO + st * Q_TILE + Q_HALF, &tm_do, &bars[QDO_FULL + st], 64, qt * BQ, h, b);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // All 32 lanes run the control flow (waits, address arithmetic stay warp-uniform); one elected lane issues.
    if (n > 0) {
      constexpr uint32_t HI = desc_hi_sw128(1024);
      const uint32_t k_km = desc_lo(smem_u32(sK), 16), v_km = desc_lo(smem_u32(sV), 16);        // K-major views
      const uint32_t k_mn = desc_lo(smem_u32(sK), KV_HALF);                                      // K^T as MN-major A
      const uint32_t q_km0 = desc_lo(smem_u32(sQ), 16), do_km0 = desc_lo(smem_u32(sDO), 16);
      const uint32_t q_mn0 = desc_lo(smem_u32(sQ), Q_HALF), do_mn0 = desc_lo(smem_u32(sDO), Q_HALF);
      const uint32_t ds_mn0 = desc_lo(smem_u32(sDS), 16);
      mbar_wait(&bars[KV_FULL], 0, 200);
      auto issue_sdp = [&](int i) {
        const int st = i & 1, qs = i % NQS;
        const int qt = qt_first + i % nq;
        const bool nc = need_c(qt);
        if (!(p.debug_flags & 16)) mbar_wait(&bars[QDO_FULL + qs], (i / NQS) & 1, 210 + qs);
        if (lane == 0) stamp(1, i, 0);
        if (!(p.debug_flags & 2)) mbar_wait(&bars[DRAIN_DONE + st], ((i >> 1) & 1) ^ 1, 215 + st);   // dQ^T(i-2) left this TMEM buffer
        if (lane == 0) stamp(1, i, 1);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t q_km = q_km0 + qs * (Q_TILE >> 4), do_km = do_km0 + qs * (Q_TILE >> 4);
          const uint32_t d_s = tmem_base + TM_SDQ + st * BQ, d_p = tmem_base + TM_DP + st * BQ;
          if (!(p.debug_flags & 128))
#pragma unroll
          for (int k = 0; k < HD / 16; ++k) {
            const uint32_t ko_kv = ((k >> 2) * KV_HALF + (k & 3) * 32) >> 4, ko_q = ((k >> 2) * Q_HALF + (k & 3) * 32) >> 4;
            if (k == 0) mma_ss_c<false>(d_s, desc64(HI, k_km + ko_kv), desc64(HI, q_km + ko_q), IDESC_ST);
            else mma_ss_c<true>(d_s, desc64(HI, k_km + ko_kv), desc64(HI, q_km + ko_q), IDESC_ST);
          }
          if (nc && !(p.debug_flags & 128)) {
#pragma unroll
            for (int k = 0; k < HD / 16; ++k) {
              const uint32_t ko_kv = ((k >> 2) * KV_HALF + (k & 3) * 32) >> 4, ko_q = ((k >> 2) * Q_HALF + (k & 3) * 32) >> 4;
              if (k == 0) mma_ss_c<false>(d_p, desc64(HI, v_km + ko_kv), desc64(HI, do_km + ko_q), IDESC_ST);
              else mma_ss_c<true>(d_p, desc64(HI, v_km + ko_kv), desc64(HI, do_km + ko_q), IDESC_ST);
            }
          }
          tc_commit(&bars[S_FULL + st]);
        }
        __syncwarp();
        if (lane == 0) stamp(1, i, 2);
      };
      bool dv_started = false;
      issue_sdp(0);
      for (int i = 0; i < n; ++i) {
        const int st = i & 1, qs = i % NQS;
        const int qt = qt_first + i % nq;
        const bool nc = need_c(qt);
        if (i + 1 < n) issue_sdp(i + 1);
        mbar_wait(&bars[SOFT_DONE + st], (i >> 1) & 1, 220 + st);
        if (lane == 0) stamp(1, i, 3);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t q_mn = q_mn0 + qs * (Q_TILE >> 4), do_mn = do_mn0 + qs * (Q_TILE >> 4);
          const uint32_t ds_mn = ds_mn0 + st * (DS_TILE >> 4);
          const uint32_t a_p = tmem_base + TM_SDQ + st * BQ, a_ds = tmem_base + TM_DP + st * BQ;
          if (nc && !(p.debug_flags & 64)) {
            // dV += P^T dO_i : K = 16 queries per step (8 packed TMEM cols: queries 0-31 at cols 0.., 32-63 at cols 48..)
            if (dv_started) {
#pragma unroll
              for (int k = 0; k < BQ / 16; ++k)
                mma_ts_c<true>(tmem_base + TM_DV, a_p + (k < 2 ? k * 8 : 48 + (k - 2) * 8), desc64(HI, do_mn + k * (2048 >> 4)), IDESC_DKV);
            } else {
#pragma unroll
              for (int k = 0; k < BQ / 16; ++k) {
                if (k == 0) mma_ts_c<false>(tmem_base + TM_DV, a_p, desc64(HI, do_mn), IDESC_DKV);
                else mma_ts_c<true>(tmem_base + TM_DV, a_p + (k < 2 ? k * 8 : 48 + (k - 2) * 8), desc64(HI, do_mn + k * (2048 >> 4)), IDESC_DKV);
              }
            }
          }
          // dK += dS^T Q_i
          if (p.debug_flags & 64) {
          } else if (i > 0) {
#pragma unroll
            for (int k = 0; k < BQ / 16; ++k)
              mma_ts_c<true>(tmem_base + TM_DK, a_ds + (k < 2 ? k * 8 : 48 + (k - 2) * 8), desc64(HI, q_mn + k * (2048 >> 4)), IDESC_DKV);
          } else {
#pragma unroll
            for (int k = 0; k < BQ / 16; ++k) {
              if (k == 0) mma_ts_c<false>(tmem_base + TM_DK, a_ds, desc64(HI, q_mn), IDESC_DKV);
              else mma_ts_c<true>(tmem_base + TM_DK, a_ds + (k < 2 ? k * 8 : 48 + (k - 2) * 8), desc64(HI, q_mn + k * (2048 >> 4)), IDESC_DKV);
            }
          }
          // dQ^T = K^T dS^T : K = 16 keys per step
          if (!(p.debug_flags & 32))
#pragma unroll
          for (int k = 0; k < BN / 16; ++k) {
            if (k == 0) mma_ss_c<false>(a_p, desc64(HI, k_mn), desc64(HI, ds_mn), IDESC_DQT);
            else mma_ss_c<true>(a_p, desc64(HI, k_mn + k * (2048 >> 4)), desc64(HI, ds_mn + k * (2048 >> 4)), IDESC_DQT);
          }
          tc_commit(&bars[QDO_EMPTY + qs]);
          tc_commit(&bars[DQ_FULL + st]);
        }
        __syncwarp();
        if (lane == 0) stamp(1, i, 4);
        if (nc) dv_started = true;
      }
      if (elect_one()) tc_commit(&bars[DKV_FULL]);
      __syncwarp();
    }
  } else if (warp < 10) {
    // ===================== softmax warps: one key row x 32 query columns per thread =====================
    // warps 2-5 take query columns 0-31 of the tile, warps 6-9 columns 32-63 (two warps per TMEM lane quarter)
    const int quarter = warp & 3;
    const int hf = warp >= 6 ? 1 : 0;
    const int r = quarter * 32 + lane;
    const int s = s0 + r;
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    const float scale2 = p.scale * LOG2E;
    const int sid = threadIdx.x - 64;                   // 0..255 inside the softmax group
    const bool key_in = s < S;
    const bool key_ok = key_in && (p.key_mask == nullptr || p.key_mask[(int64_t)b * S + s] != 0);
    const uint32_t pcol = hf ? 48 : 0;                  // where this half's packed P^T / dS^T columns go

    auto load_stats = [&](int h, int qt, float4& v) {  // threads 0..63 of the group: column sid of tile (h, qt)
      const int t = qt * BQ + sid;
      v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (t < T) {
        const int64_t ridx = ((int64_t)b * p.H + h) * T + t;
        v.x = -p.lse_c[ridx] * LOG2E;                   // negated: exp2(fma(s, scale2, -lse))
        v.y = p.delta_ws[ridx];
        v.z = -p.dz_ws[ridx];
        v.w = p.active ? -p.lse_full[ridx] * LOG2E : 0.f;
      }
    };
    float4 nxt;
    if (n > 0 && sid < BQ) {
      load_stats(kvh * G, qt_first, nxt);
      stats[sid] = nxt;
    }
    int cur_h = kvh * G, cur_qt = qt_first;             // (head, query tile) of iteration i, advanced without div/mod
    for (int i = 0; i < n; ++i) {
      const int st = i & 1;
      const int qt = cur_qt;
      const int t0 = qt * BQ;
      const bool nc = need_c(qt);
      if (++cur_qt == n_qt) {
        cur_qt = qt_first;
        ++cur_h;
      }
      if (i + 1 < n && sid < BQ) load_stats(cur_h, cur_qt, nxt);   // prefetch next iteration's row statistics
      const int srole = (lane == 0 && (warp == 2 || warp == 6)) ? (warp == 2 ? 2 : 3) : -1;
      if (srole >= 0) stamp(srole, i, 0);
      if (!(p.debug_flags & 2048)) named_bar_sync(1, 256);      // stats[st] complete; stats[st^1] no longer read
      if (srole >= 0) stamp(srole, i, 1);
      mbar_wait(&bars[S_FULL + st], (i >> 1) & 1, 310 + st);
      if (srole >= 0) stamp(srole, i, 2);
      tc_fence_after();
      const uint32_t t_s = tmem_base + lane_off + TM_SDQ + st * BQ;
      const uint32_t t_dp = tmem_base + lane_off + TM_DP + st * BQ;
      uint8_t* ds_row = sDS + st * DS_TILE + r * 128;
      const float4* stt = stats + st * BQ + 32 * hf;
      // visibility of queries t0+32*hf+c (c = 0..31) for key s: visible iff s <= t + off
      const uint32_t trow = low_bits_b(T - t0 - 32 * hf);
      const uint32_t cz = p.causal ? ~low_bits_b(s - off - t0 - 32 * hf) : 0xFFFFFFFFu;
      const uint32_t vw = key_ok ? (trow & cz) : 0u;
      const uint32_t xw = (p.active && key_in) ? (trow & ~vw) : 0u;
      float sv[32], dpv[32];
      tmem_ld32(t_s + 32 * hf, sv);
      if (nc) tmem_ld32(t_dp + 32 * hf, dpv);
      tc_wait_ld();
      if (srole >= 0) stamp(srole, i, 3);
      uint32_t pk[16], dk_[16];
      const bool all_v = __all_sync(0xffffffffu, vw == 0xFFFFFFFFu);
      const bool all_x = __all_sync(0xffffffffu, vw == 0u && xw == 0xFFFFFFFFu);
      if (p.debug_flags & 4) {
#pragma unroll
        for (int c2 = 0; c2 < 16; ++c2) { pk[c2] = __float_as_uint(sv[c2]); dk_[c2] = __float_as_uint(sv[c2 + 16]); }
      } else if (all_v && nc) {
        // interior causal tile: every entry visible.  dS (unscaled) = P (dP - Delta'); scale is applied to dK / dQ later
#pragma unroll
        for (int c2 = 0; c2 < 16; ++c2) {
          const float4 a = stt[2 * c2], bq = stt[2 * c2 + 1];
          const float p0 = ex2(fmaf(sv[2 * c2], scale2, a.x)), p1 = ex2(fmaf(sv[2 * c2 + 1], scale2, bq.x));
          pk[c2] = pack_bf16(p0, p1);
          dk_[c2] = pack_bf16(p0 * (dpv[2 * c2] - a.y), p1 * (dpv[2 * c2 + 1] - bq.y));
        }
      } else if (all_x) {
        // tile above the diagonal (shift term only): dS = -dz exp2(s - lse_full)
#pragma unroll
        for (int c2 = 0; c2 < 16; ++c2) {
          const float4 a = stt[2 * c2], bq = stt[2 * c2 + 1];
          pk[c2] = 0u;
          dk_[c2] = pack_bf16(a.z * ex2(fmaf(sv[2 * c2], scale2, a.w)), bq.z * ex2(fmaf(sv[2 * c2 + 1], scale2, bq.w)));
        }
      } else {
#pragma unroll
        for (int c2 = 0; c2 < 16; ++c2) {
          float pv[2], dsv[2];
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int c = 2 * c2 + e;
            const float4 stc = stt[c];
            float pe = 0.f, de = 0.f;
            if ((vw >> c) & 1u) {
              pe = ex2(fmaf(sv[c], scale2, stc.x));
              de = pe * ((nc ? dpv[c] : 0.f) - stc.y);
            } else if ((xw >> c) & 1u) {
              de = stc.z * ex2(fmaf(sv[c], scale2, stc.w));
            }
            pv[e] = pe;
            dsv[e] = de;
          }
          pk[c2] = pack_bf16(pv[0], pv[1]);
          dk_[c2] = pack_bf16(dsv[0], dsv[1]);
        }
      }
      if (srole >= 0) stamp(srole, i, 4);
      if (nc) tmem_st16(t_s + pcol, pk);                        // P^T  (A operand of dV)
      tmem_st16(t_dp + pcol, dk_);                              // dS^T (A operand of dK)
      // dS^T row to smem: 32 queries = 64 B = four 16-byte chunks, 128B-swizzled by key row
      if (!(p.debug_flags & 512))
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) {
        const int chunk = hf * 4 + ch;
        *reinterpret_cast<uint4*>(ds_row + ((chunk ^ (r & 7)) << 4)) =
            make_uint4(dk_[4 * ch], dk_[4 * ch + 1], dk_[4 * ch + 2], dk_[4 * ch + 3]);
      }
      if (i + 1 < n && sid < BQ) stats[(st ^ 1) * BQ + sid] = nxt;
      tc_wait_st();
      if (!(p.debug_flags & 1024)) fence_proxy_async();         // smem dS^T visible to the tensor-core (async) proxy
      tc_fence_before();
      mbar_arrive(&bars[SOFT_DONE + st]);
      if (srole >= 0) stamp(srole, i, 5);
    }
    // ---- epilogue: warps 2-5 write this key's dV row, warps 6-9 its dK row (x softmax scale) ----
    if (n > 0) {
      mbar_wait(&bars[DKV_FULL], 0, 330);
      tc_fence_after();
    }
    {
      const int which = hf;                              // 0: dV, 1: dK
      const float mul = which ? p.scale : 1.f;
      __nv_bfloat16* dst0 = reinterpret_cast<__nv_bfloat16*>(which ? p.dk : p.dv) + (((int64_t)b * p.Hkv + kvh) * S + s) * HD;
#pragma unroll
      for (int i4 = 0; i4 < 4; ++i4) {
        float o[32];
        if (n > 0) {
          tmem_ld32(tmem_base + lane_off + (which ? TM_DK : TM_DV) + 32 * i4, o);
          tc_wait_ld();
        } else {
#pragma unroll
          for (int c = 0; c < 32; ++c) o[c] = 0.f;
        }
        if (key_in) {
#pragma unroll
          for (int c8 = 0; c8 < 4; ++c8) {
            float f[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) f[e] = o[8 * c8 + e] * mul;
            store8(dst0 + 32 * i4 + 8 * c8, f);
          }
        }
      }
    }
    tc_fence_before();
  } else {
    // ===================== drain warps: dQ^T (TMEM, d on lanes) -> smem fp32 [64][128] -> bulk reduce-add =====================
    const int quarter = warp & 3;
    const int c = quarter * 32 + lane;                  // head-dim index owned by this thread
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    const int did = threadIdx.x - 320;                  // 0..127
    for (int i = 0; i < n; ++i) {
      const int st = i & 1;
      const int h = kvh * G + i / nq, qt = qt_first + i % nq;
      const bool dtr = (warp == 10 && lane == 0);
      if (dtr) stamp(4, i, 0);
      mbar_wait(&bars[DQ_FULL + st], (i >> 1) & 1, 410 + st);
      if (dtr) stamp(4, i, 1);
      tc_fence_after();
      float v0[32], v1[32];
      tmem_ld32(tmem_base + lane_off + TM_SDQ + st * BQ, v0);
      tmem_ld32(tmem_base + lane_off + TM_SDQ + st * BQ + 32, v1);
      tc_wait_ld();
      tc_fence_before();
      mbar_arrive(&bars[DRAIN_DONE + st]);              // TMEM buffer may be overwritten by S^T(i+2)
      if (dtr) stamp(4, i, 2);
      if (p.debug_flags & 256) continue;
      const int rows = min(BQ, T - qt * BQ);
      float* gdst = p.dq_acc + (((int64_t)b * p.H + h) * T + (int64_t)qt * BQ) * HD;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        if (did == 0) bulk_wait_read0();                // previous bulk reduce finished reading the staging tile
        named_bar_sync(2, 128);
#pragma unroll
        for (int t = 0; t < 32; ++t) stage[t * HD + c] = half ? v1[t] : v0[t];
        fence_proxy_async();
        named_bar_sync(3, 128);
        if (did == 0 && !(p.debug_flags & 1)) {
          const int r0 = half * 32, nr = min(32, rows - r0);
          if (nr > 0) {
            bulk_reduce_add_f32(gdst + (int64_t)r0 * HD, stage, (uint32_t)nr * HD * 4);
            bulk_commit();
          }
        }
      }
      if (dtr) stamp(4, i, 3);
    }
    if (did == 0) bulk_wait0();
  }
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

int bwd(const AttnParams& p, cudaStream_t st) {
  using namespace bw;
  int rc;
  // the softmax scale is NOT applied to dS in the main kernel: dK is scaled in its epilogue, dQ in the
  // fp32 -> bf16 conversion; the dq seed (dz * w) is pre-divided so that it comes out unscaled
  AttnParams pp = p;
  pp.dq_seed_mul = 1.f / p.scale;
  pp.dq_out_mul = p.scale;
  pp.trace = debug_trace_capacity() >= 24 * 64 ? debug_trace_buffer() : nullptr;
  if (const char* e = getenv("MIMIC_DEBUG_FLAGS")) pp.debug_flags = atoi(e);   // bit 0: skip the dQ reduce (timing experiments only)
  if ((rc = simt::bwd_prep_bf16_d128(pp, st))) return rc;
  CUtensorMap tq, tk, tv, tdo;
  if ((rc = make_map(&tq, p.q, p.B, p.H, p.T, p.q_sb, p.q_sh, p.q_st, BQ))) return rc;
  if ((rc = make_map(&tk, p.k, p.B, p.Hkv, p.S, p.k_sb, p.k_sh, p.k_ss, BN))) return rc;
  if ((rc = make_map(&tv, p.v, p.B, p.Hkv, p.S, p.v_sb, p.v_sh, p.v_ss, BN))) return rc;
  if ((rc = make_map(&tdo, p.d_out, p.B, p.H, p.T, p.go_sb, p.go_sh, p.go_st, BQ))) return rc;
  static bool configured = false;
  if (!configured) {
    MIMIC_CUDA_OK(cudaFuncSetAttribute(attn_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    configured = true;
  }
  dim3 grid((p.S + BN - 1) / BN, p.Hkv, p.B);
  attn_bwd_tc_kernel<<<grid, THREADS, SMEM_BYTES, st>>>(tq, tk, tv, tdo, pp);
  MIMIC_LAUNCH_OK("attn_bwd_tc_kernel");
  return simt::convert_dq_bf16(pp, st);
}

}  // namespace tc
}  // namespace mimic
