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
// TMEM map (512 cols): dV [0,128) dK [128,256) S0 [256,320) S1 [320,384) dP0 [384,448) dP1 [448,512).
// S[st] holds S^T then P^T; dP[st] holds dP^T, then dS^T, then dQ^T (so S^T(i+2) only waits for dV(i), see the issuers).
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
constexpr int STAGE_BYTES = 32 * HD * 4;   // fp32 dQ staging: per drain warp two 2 KB buffers of [16 rows][32 cols] (16 KB in all)
constexpr int NQS = 3;                     // Q / dO TMA stages
constexpr int THREADS = 736;                // warp 0 TMA, warps 1 / 22 MMA issuers A / B, warps 2-9 / 10-17 softmax groups 0 / 1, warps 18-21 drain
constexpr float LOG2E = 1.4426950408889634f;

constexpr int OFF_K = 0, OFF_V = KV_TILE, OFF_Q = 2 * KV_TILE, OFF_DO = OFF_Q + NQS * Q_TILE, OFF_DS = OFF_DO + NQS * Q_TILE,
              OFF_STAGE = OFF_DS + 2 * DS_TILE, OFF_STATS = OFF_STAGE + STAGE_BYTES, OFF_BAR = OFF_STATS + 4 * BQ * 16,
              OFF_TMEM = OFF_BAR + 24 * 8, SMEM_BYTES = OFF_TMEM + 16;

enum Bar { KV_FULL = 0, QDO_FULL = 1, QDO_EMPTY = 4, S_FULL = 7, SOFT_DONE = 9, DQ_FULL = 11, DRAIN_DONE = 13, DKV_FULL = 15, DV_DONE = 16, DP_FULL = 18, N_BARS = 20 };

constexpr uint32_t TM_DV = 0, TM_DK = 128, TM_SDQ = 256, TM_DP = 384;
constexpr uint32_t IDESC_ST = idesc_bf16(128, BQ, 0, 0);     // S^T, dP^T : A, B K-major
constexpr uint32_t IDESC_DKV = idesc_bf16(128, HD, 0, 1);    // dV, dK    : A in TMEM, B MN-major
constexpr uint32_t IDESC_DQT = idesc_bf16(128, BQ, 1, 1);    // dQ^T      : A = K MN-major, B = dS^T MN-major
}  // namespace bw

__device__ __forceinline__ uint32_t low_bits_b(int n) { return n >= 32 ? 0xFFFFFFFFu : (n <= 0 ? 0u : ((1u << n) - 1u)); }

// DBG = true builds the instrumented twin (timeline stamps + MIMIC_DEBUG_FLAGS experiments); the production
// instantiation carries none of it: the single MMA-issuing warp is instruction-fetch bound, every instruction counts.
template <bool DBG>
__global__ void __launch_bounds__(bw::THREADS, 1)
attn_bwd_tc_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                   const __grid_constant__ CUtensorMap tm_v, const __grid_constant__ CUtensorMap tm_do,
                   const __grid_constant__ CUtensorMap tm_dk, const __grid_constant__ CUtensorMap tm_dv,
                   const __grid_constant__ CUtensorMap tm_dq, AttnParams p, int debug_life) {
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
  float4* stats = reinterpret_cast<float4*>(smem + OFF_STATS);      // [group][2][64] (-lse_c*log2e, Delta', -dz, -lse_full*log2e)
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + OFF_TMEM);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kt = blockIdx.x, kvh = blockIdx.y, b = blockIdx.z;
  // timeline trace (debug): role r writes record k of iteration i at trace[(i * 8 + r) * 8 + k]
  const bool tracing = DBG && p.trace != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0;
  auto stamp = [&](int role, int i, int k) {
    if constexpr (DBG) {
      if (tracing && i < 24) p.trace[(i * 8 + role) * 8 + k] = clock64();
    }
  };
  auto dbg = [&](int bit) { return DBG && (p.debug_flags & bit) != 0; };
  // per-CTA life cycle (debug): trace[2048 + 8 cta + k] = SM id / entry / prologue done / K,V landed / last MMA done (DKV_FULL
  // seen) / dK,dV stores issued / exit, in nanoseconds of %globaltimer — where do the ~17 us a CTA costs beyond its iterations go?
  const int cta_lin = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
  auto life = [&](int k) {
    if constexpr (DBG) {
      if (p.trace != nullptr && debug_life) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        p.trace[2048 + 8 * cta_lin + k] = t;
      }
    }
  };
  if (threadIdx.x == 0) {
    if constexpr (DBG) {
      if (p.trace != nullptr && debug_life) {
        unsigned smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        p.trace[2048 + 8 * cta_lin] = smid;
      }
    }
    life(1);
  }
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
        mbar_init(&bars[DV_DONE + i], 1);
        mbar_init(&bars[DP_FULL + i], 1);
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
  if (threadIdx.x == 0) life(2);

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      mbar_expect_tx(&bars[KV_FULL], 2 * KV_TILE);
      tma_load_4d(sK, &tm_k, &bars[KV_FULL], 0, s0, kvh, b);
      tma_load_4d(sK + KV_HALF, &tm_k, &bars[KV_FULL], 64, s0, kvh, b);
      tma_load_4d(sV, &tm_v, &bars[KV_FULL], 0, s0, kvh, b);
      tma_load_4d(sV + KV_HALF, &tm_v, &bars[KV_FULL], 64, s0, kvh, b);
      int st = 0, ph = 1, h = kvh * G, qt = qt_first;   // ring stage, EMPTY-wait parity, (head, query tile) of iteration i
      for (int i = 0; i < n; ++i) {
        const bool nc = need_c(qt);
        if (!dbg(8)) mbar_wait(&bars[QDO_EMPTY + st], ph, 100 + st);
        stamp(0, i, 0);
        mbar_expect_tx(&bars[QDO_FULL + st], nc ? 2 * Q_TILE : Q_TILE);
        tma_load_4d(sQ + st * Q_TILE, &tm_q, &bars[QDO_FULL + st], 0, qt * BQ, h, b);
        tma_load_4d(sQ + st * Q_TILE + Q_HALF, &tm_q, &bars[QDO_FULL + st], 64, qt * BQ, h, b);
        if (nc) {
          tma_load_4d(sDO + st * Q_TILE, &tm_do, &bars[QDO_FULL + st], 0, qt * BQ, h, b);
          tma_load_4d(sDO + st * Q_TILE + Q_HALF, &tm_do, &bars[QDO_FULL + st], 64, qt * BQ, h, b);
        }
        if (DBG && tracing) {                            // debug: when did this stage's bytes land?
          stamp(0, i, 1);
          mbar_wait(&bars[QDO_FULL + st], ph ^ 1, 110 + st);
          stamp(0, i, 2);
        }
        if (++qt == n_qt) {
          qt = qt_first;
          ++h;
        }
        if (++st == NQS) {
          st = 0;
          ph ^= 1;
        }
      }
    }
  } else if (warp == 1 || warp == 22) {
    // ===================== two MMA issuers =====================
    // One issuing warp runs at ~8.7 cycles per instruction (~350 instructions per iteration): it was the bottleneck.
    // The stream is split by operand so each accumulator keeps ONE issuer, and ordered so that the next S^T only
    // waits for dV:
    //   warp 1  (A), after SOFT_DONE(i):  dV(i) | commit DV_DONE[st]   dK(i)   dQ^T(i) -> dP[st] | commit DQ_FULL, QDO_EMPTY
    //   warp 22 (B):  S^T(i+2) -> S[st]   after DV_DONE(i)   (dV(i) has read P^T(i))            | commit S_FULL[st]
    //                 dP^T(i+2) -> dP[st] after DRAIN_DONE(i) (dQ^T(i) drained; dK(i) read dS^T) | commit DP_FULL[st]
    // The softmax group starts its exponentials on S^T(i+2) while dK(i), dQ^T(i), the drain and dP^T(i+2) are still
    // running.  Every cross-issuer hazard is an explicit mbarrier; inside one issuer the tensor pipe runs in issue order.
    // All 32 lanes run the control flow (waits, address arithmetic stay warp-uniform); one elected lane issues.
    if (n > 0) {
      constexpr uint32_t HI = desc_hi_sw128(1024);
      const uint32_t k_km = desc_lo(smem_u32(sK), 16), v_km = desc_lo(smem_u32(sV), 16);        // K-major views
      const uint32_t k_mn = desc_lo(smem_u32(sK), KV_HALF);                                      // K^T as MN-major A
      const uint32_t q_km0 = desc_lo(smem_u32(sQ), 16), do_km0 = desc_lo(smem_u32(sDO), 16);
      const uint32_t q_mn0 = desc_lo(smem_u32(sQ), Q_HALF), do_mn0 = desc_lo(smem_u32(sDO), Q_HALF);
      const uint32_t ds_mn0 = desc_lo(smem_u32(sDS), 16);
      auto next_qt = [&](int qt) { return qt + 1 == n_qt ? qt_first : qt + 1; };
      mbar_wait(&bars[KV_FULL], 0, 200);
      if (warp == 1 && lane == 0) life(3);
      if (warp == 22) {
        // ---------------- issuer B: S^T / dP^T of iteration i2 = 0 .. n-1 ----------------
        int qt2 = qt_first, qs2 = 0, qph2 = 0;
        for (int i2 = 0; i2 < n; ++i2) {
          const int st2 = i2 & 1;
          const bool nc2 = need_c(qt2);
          const uint32_t prev_par = ((i2 >> 1) & 1) ^ 1;       // parity of the phase completed by iteration i2 - 2 (passes for i2 < 2)
          mbar_wait(&bars[QDO_FULL + qs2], qph2, 210 + qs2);
          if (lane == 0) stamp(5, i2, 0);
          mbar_wait(&bars[DV_DONE + st2], prev_par, 212 + st2);
          if (lane == 0) stamp(5, i2, 1);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t q_km = q_km0 + qs2 * (Q_TILE >> 4), d_s = tmem_base + TM_SDQ + st2 * BQ;
#pragma unroll
            for (int k = 0; k < HD / 16; ++k) {
              const uint32_t ko_kv = ((k >> 2) * KV_HALF + (k & 3) * 32) >> 4, ko_q = ((k >> 2) * Q_HALF + (k & 3) * 32) >> 4;
              if (k == 0) mma_ss_c<false>(d_s, desc64(HI, k_km + ko_kv), desc64(HI, q_km + ko_q), IDESC_ST);
              else mma_ss_c<true>(d_s, desc64(HI, k_km + ko_kv), desc64(HI, q_km + ko_q), IDESC_ST);
            }
            tc_commit(&bars[S_FULL + st2]);
          }
          __syncwarp();
          if (!dbg(2)) mbar_wait(&bars[DRAIN_DONE + st2], prev_par, 215 + st2);
          if (lane == 0) stamp(5, i2, 2);
          tc_fence_after();
          if (elect_one()) {
            if (nc2) {
              const uint32_t do_km = do_km0 + qs2 * (Q_TILE >> 4), d_p = tmem_base + TM_DP + st2 * BQ;
#pragma unroll
              for (int k = 0; k < HD / 16; ++k) {
                const uint32_t ko_kv = ((k >> 2) * KV_HALF + (k & 3) * 32) >> 4, ko_q = ((k >> 2) * Q_HALF + (k & 3) * 32) >> 4;
                if (k == 0) mma_ss_c<false>(d_p, desc64(HI, v_km + ko_kv), desc64(HI, do_km + ko_q), IDESC_ST);
                else mma_ss_c<true>(d_p, desc64(HI, v_km + ko_kv), desc64(HI, do_km + ko_q), IDESC_ST);
              }
            }
            tc_commit(&bars[DP_FULL + st2]);                   // without MMAs in between: dP[st2] is simply free for dS^T
          }
          __syncwarp();
          if (lane == 0) stamp(5, i2, 3);
          qt2 = next_qt(qt2);
          if (++qs2 == NQS) {
            qs2 = 0;
            qph2 ^= 1;
          }
        }
      } else {
        // ---------------- issuer A: dV / dK / dQ^T of iteration i ----------------
        int qt_i = qt_first, qs = 0;
        bool dv_started = false;
        for (int i = 0; i < n; ++i) {
          const int st = i & 1;
          const bool nc = need_c(qt_i);
          mbar_wait(&bars[SOFT_DONE + st], (i >> 1) & 1, 220 + st);
          if (lane == 0) stamp(1, i, 3);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t q_mn = q_mn0 + qs * (Q_TILE >> 4), do_mn = do_mn0 + qs * (Q_TILE >> 4);
            const uint32_t ds_mn = ds_mn0 + st * (DS_TILE >> 4);
            const uint32_t a_p = tmem_base + TM_SDQ + st * BQ, a_ds = tmem_base + TM_DP + st * BQ;
            if (nc) {
              // dV += P^T dO_i : K = 16 queries per step; step k's 8 packed TMEM cols sit at [16k, 16k+8) of the buffer
#pragma unroll
              for (int k = 0; k < BQ / 16; ++k)
                mma_ts(tmem_base + TM_DV, a_p + 16 * k, desc64(HI, do_mn + k * (2048 >> 4)), IDESC_DKV, (k > 0 || dv_started) ? 1u : 0u);
            }
            tc_commit(&bars[DV_DONE + st]);                    // S[st] may take S^T(i+2)
            // dK += dS^T Q_i   (A operand: dS^T in dP[st])
#pragma unroll
            for (int k = 0; k < BQ / 16; ++k)
              mma_ts(tmem_base + TM_DK, a_ds + 16 * k, desc64(HI, q_mn + k * (2048 >> 4)), IDESC_DKV, (k > 0 || i > 0) ? 1u : 0u);
            // dQ^T = K^T dS^T : K = 16 keys per step, B from shared memory; lands in dP[st] over dS^T(i) (dK(i) has read it)
#pragma unroll
            for (int k = 0; k < BN / 16; ++k) {
              if (k == 0) mma_ss_c<false>(a_ds, desc64(HI, k_mn), desc64(HI, ds_mn), IDESC_DQT);
              else mma_ss_c<true>(a_ds, desc64(HI, k_mn + k * (2048 >> 4)), desc64(HI, ds_mn + k * (2048 >> 4)), IDESC_DQT);
            }
            tc_commit(&bars[DQ_FULL + st]);
            tc_commit(&bars[QDO_EMPTY + qs]);
          }
          __syncwarp();
          if (lane == 0) stamp(1, i, 4);
          if (nc) dv_started = true;
          qt_i = next_qt(qt_i);
          if (++qs == NQS) qs = 0;
        }
        if (elect_one()) tc_commit(&bars[DKV_FULL]);
        __syncwarp();
      }
    }
  } else if (warp < 18) {
    // ===================== softmax warps: two ping-pong groups of 8 warps =====================
    // Group g (warps 2+8g .. 9+8g) owns the iterations with i & 1 == g, i.e. TMEM / dS^T buffer st = g, so one group's
    // TMEM loads overlap the other's exponentials and stores.  Inside a group, warp (quarter, hf) owns key rows of its
    // TMEM lane quarter and query columns [32 hf, 32 hf + 32), processed as two 16-column passes (register budget).
    const int g = (warp - 2) >> 3;
    const int hf = ((warp - 2) >> 2) & 1;
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;
    const int s = s0 + r;
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    const float scale2 = p.scale * LOG2E;
    const int sid = threadIdx.x - 64 - 256 * g;         // 0..255 inside the group
    const bool key_in = s < S;
    const bool key_ok = key_in && (p.key_mask == nullptr || p.key_mask[(int64_t)b * S + s] != 0);
    const int st = g;

    // Row statistics of the group's NEXT iteration are fetched as raw values at the top of an iteration and only
    // touched (negate / scale) when they are stored to shared memory at its end, so the global-load latency never sits
    // on the softmax critical path.  Threads 0..63 of the group own column sid of the tile.
    auto load_stats = [&](int h, int qt, float4& v) {
      const int t = min(qt * BQ + sid, T - 1);          // clamped: no branch around the loads
      const int64_t ridx = ((int64_t)b * p.H + h) * T + t;
      v.x = p.lse_c[ridx];
      v.y = p.delta_ws[ridx];
      v.z = p.dz_ws[ridx];
      v.w = p.active ? p.lse_full[ridx] : 0.f;
    };
    auto finish_stats = [&](int qt, float4 v) {         // (-lse_c*log2e, Delta', -dz, -lse_full*log2e); rows >= T are zero
      const bool in = qt * BQ + sid < T;
      return in ? make_float4(-v.x * LOG2E, v.y, -v.z, -v.w * LOG2E) : make_float4(0.f, 0.f, 0.f, 0.f);
    };
    int cur_h = kvh * G, cur_qt = qt_first;             // (head, query tile) of the group's current iteration
    auto advance = [&](int& h, int& qt) {
      if (++qt == n_qt) {
        qt = qt_first;
        ++h;
      }
    };
    if (g == 1) advance(cur_h, cur_qt);
    float4* gstats = stats + g * 2 * BQ;                // [2][64] per group
    float4 nxt = make_float4(0.f, 0.f, 0.f, 0.f);
    if (g < n && sid < BQ) {
      load_stats(cur_h, cur_qt, nxt);
      gstats[sid] = finish_stats(cur_qt, nxt);
    }
    int sb = 0;                                          // statistics buffer of the current iteration
    for (int i = g; i < n; i += 2) {
      const int qt = cur_qt;
      const int t0 = qt * BQ;
      const bool nc = need_c(qt);
      advance(cur_h, cur_qt);
      advance(cur_h, cur_qt);                            // -> iteration i + 2
      if (i + 2 < n && sid < BQ) load_stats(cur_h, cur_qt, nxt);
      const int srole = (lane == 0 && (warp == 2 || warp == 10)) ? 2 + g : -1;
      if (srole >= 0) stamp(srole, i, 0);
      named_bar_sync(1 + g, 256);                        // gstats[sb] complete; gstats[sb^1] no longer read
      if (srole >= 0) stamp(srole, i, 1);
      mbar_wait(&bars[S_FULL + st], (i >> 1) & 1, 310 + st);
      if (srole >= 0) stamp(srole, i, 2);
      tc_fence_after();
      uint8_t* ds_row = sDS + st * DS_TILE + r * 128;
      // ---- phase 1 (needs S^T only): val = P for visible entries, -dz exp2(s - lse_full) for excluded ones, packed bf16.
      // P^T goes to TMEM (A operand of dV); val stays in registers for phase 2.
      uint32_t val[2][8];
      uint32_t vws[2], xws[2];
#pragma unroll
      for (int pass = 0; pass < 2; ++pass) {
        const int cq = 2 * hf + pass;                    // 16-column group of the 64-wide tile
        const uint32_t t_s = tmem_base + lane_off + TM_SDQ + st * BQ + 16 * cq;
        const float4* stt = gstats + sb * BQ + 16 * cq;
        // visibility of queries t0+16*cq+c (c = 0..15) for key s: visible iff s <= t + off
        const uint32_t trow = low_bits_b(T - t0 - 16 * cq) & 0xFFFFu;
        const uint32_t cz = (p.causal ? ~low_bits_b(s - off - t0 - 16 * cq) : 0xFFFFFFFFu) & 0xFFFFu;
        const uint32_t vw = key_ok ? (trow & cz) : 0u;
        const uint32_t xw = (p.active && key_in) ? (trow & ~vw) : 0u;
        vws[pass] = vw;
        xws[pass] = xw;
        float sv[16];
        tmem_ld16(t_s, sv);
        tc_wait_ld();
        if (srole >= 0 && pass == 0) stamp(srole, i, 3);
        const bool all_v = __all_sync(0xffffffffu, vw == 0xFFFFu);
        const bool all_x = __all_sync(0xffffffffu, vw == 0u && xw == 0xFFFFu);
        if (all_v) {
#pragma unroll
          for (int c2 = 0; c2 < 8; ++c2) {
            const float4 a = stt[2 * c2], bq = stt[2 * c2 + 1];
            val[pass][c2] = pack_bf16(ex2(fmaf(sv[2 * c2], scale2, a.x)), ex2(fmaf(sv[2 * c2 + 1], scale2, bq.x)));
          }
          if (nc) tmem_st8(t_s, val[pass]);
        } else if (all_x) {
#pragma unroll
          for (int c2 = 0; c2 < 8; ++c2) {
            const float4 a = stt[2 * c2], bq = stt[2 * c2 + 1];
            val[pass][c2] = pack_bf16(a.z * ex2(fmaf(sv[2 * c2], scale2, a.w)), bq.z * ex2(fmaf(sv[2 * c2 + 1], scale2, bq.w)));
          }
          if (nc) {
            const uint32_t zero[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
            tmem_st8(t_s, zero);
          }
        } else {
          uint32_t pk[8];
#pragma unroll
          for (int c2 = 0; c2 < 8; ++c2) {
            float ve[2], pe[2];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int c = 2 * c2 + e;
              const float4 stc = stt[c];
              ve[e] = 0.f;
              pe[e] = 0.f;
              if ((vw >> c) & 1u) {
                pe[e] = ex2(fmaf(sv[c], scale2, stc.x));
                ve[e] = pe[e];
              } else if ((xw >> c) & 1u) {
                ve[e] = stc.z * ex2(fmaf(sv[c], scale2, stc.w));
              }
            }
            val[pass][c2] = pack_bf16(ve[0], ve[1]);
            pk[c2] = pack_bf16(pe[0], pe[1]);
          }
          if (nc) tmem_st8(t_s, pk);
        }
      }
      if (srole >= 0) stamp(srole, i, 4);
      // ---- phase 2 (needs dP^T; dP[st] is free for dS^T once DP_FULL fired): dS = P (dP - Delta') on visible entries
      mbar_wait(&bars[DP_FULL + st], (i >> 1) & 1, 312 + st);
      tc_fence_after();
#pragma unroll
      for (int pass = 0; pass < 2; ++pass) {
        const int cq = 2 * hf + pass;
        const uint32_t t_dp = tmem_base + lane_off + TM_DP + st * BQ + 16 * cq;
        const float4* stt = gstats + sb * BQ + 16 * cq;
        const uint32_t vw = vws[pass], xw = xws[pass];
        uint32_t dk_[8];
        const bool any_v = __any_sync(0xffffffffu, vw != 0u);
        if (any_v && nc) {
          float dpv[16];
          tmem_ld16(t_dp, dpv);
          tc_wait_ld();
#pragma unroll
          for (int c2 = 0; c2 < 8; ++c2) {
            // unpack bf16 pair: low half = element 2 c2, high half = element 2 c2 + 1
            const float v0 = __uint_as_float(val[pass][c2] << 16), v1 = __uint_as_float(val[pass][c2] & 0xFFFF0000u);
            const float d0 = ((vw >> (2 * c2)) & 1u) ? v0 * (dpv[2 * c2] - stt[2 * c2].y) : v0;
            const float d1 = ((vw >> (2 * c2 + 1)) & 1u) ? v1 * (dpv[2 * c2 + 1] - stt[2 * c2 + 1].y) : v1;
            dk_[c2] = pack_bf16(d0, d1);
          }
        } else if (any_v) {
          // visible entries of a tile without dP^T (cannot happen for causal tiles: nc covers every tile with a visible
          // entry; kept for non-causal edge shapes): dS = -P Delta'
#pragma unroll
          for (int c2 = 0; c2 < 8; ++c2) {
            const float v0 = __uint_as_float(val[pass][c2] << 16), v1 = __uint_as_float(val[pass][c2] & 0xFFFF0000u);
            const float d0 = ((vw >> (2 * c2)) & 1u) ? -v0 * stt[2 * c2].y : v0;
            const float d1 = ((vw >> (2 * c2 + 1)) & 1u) ? -v1 * stt[2 * c2 + 1].y : v1;
            dk_[c2] = pack_bf16(d0, d1);
          }
        } else {
#pragma unroll
          for (int c2 = 0; c2 < 8; ++c2) dk_[c2] = val[pass][c2];
        }
        (void)xw;
        tmem_st8(t_dp, dk_);                                    // dS^T (A operand of dK)
        // dS^T row to smem: 16 queries = 32 B = two 16-byte chunks, 128B-swizzled by key row
#pragma unroll
        for (int ch = 0; ch < 2; ++ch) {
          const int chunk = cq * 2 + ch;
          *reinterpret_cast<uint4*>(ds_row + ((chunk ^ (r & 7)) << 4)) =
              make_uint4(dk_[4 * ch], dk_[4 * ch + 1], dk_[4 * ch + 2], dk_[4 * ch + 3]);
        }
      }
      if (i + 2 < n && sid < BQ) gstats[(sb ^ 1) * BQ + sid] = finish_stats(cur_qt, nxt);
      sb ^= 1;
      tc_wait_st();
      fence_proxy_async();                                      // smem dS^T visible to the tensor-core (async) proxy
      tc_fence_before();
      mbar_arrive(&bars[SOFT_DONE + st]);
      if (srole >= 0) stamp(srole, i, 5);
    }
    // ---- epilogue: group 0 writes this key's dV row, group 1 its dK row (x scale); hf picks the 64-column half ----
    if (n > 0) {
      mbar_wait(&bars[DKV_FULL], 0, 330);
      tc_fence_after();
    }
    if (threadIdx.x == 64) life(4);
    {
      // dV / dK leave through shared memory (the K and V tiles are dead once DKV_FULL fired) and four TMA tile stores:
      // thread-per-row global stores touch 32 different lines per warp instruction (~4000 sector writes per CTA).
      // (Measured again in round 2, A/B in one process: writing the half rows straight from registers so that the CTA
      // need not wait ~1.3 us for the TMA unit to drain the four boxes is SLOWER — 0.859 vs 0.803 ms at T = 512,
      // 1.129 vs 1.069 at T = 1024.)
      const int which = g;                               // 0: dV -> the V tile's smem, 1: dK -> the K tile's
      const int half = hf;                               // 64-column half of the row = one [128 rows][128 B] TMA box
      const float mul = which ? p.scale : 1.f;
      uint8_t* srow = (which ? sK : sV) + half * KV_HALF + r * 128;
      mbar_wait(&bars[KV_FULL], 0, 331);                 // n == 0: the K/V loads issued in the prologue must have landed
#pragma unroll
      for (int i4 = 0; i4 < 2; ++i4) {
        float o[32];
        if (n > 0) {
          tmem_ld32(tmem_base + lane_off + (which ? TM_DK : TM_DV) + 64 * half + 32 * i4, o);
          tc_wait_ld();
        } else {
#pragma unroll
          for (int c = 0; c < 32; ++c) o[c] = 0.f;
        }
#pragma unroll
        for (int c8 = 0; c8 < 4; ++c8) {
          uint32_t pk[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) pk[e] = pack_bf16(o[8 * c8 + 2 * e] * mul, o[8 * c8 + 2 * e + 1] * mul);
          *reinterpret_cast<uint4*>(srow + (((4 * i4 + c8) ^ (r & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        }
      }
      fence_proxy_async();                               // generic-proxy smem writes -> visible to the TMA (async proxy)
      named_bar_sync(5, 512);                            // both softmax groups
      if (threadIdx.x == 64) {
        tma_store_4d(&tm_dv, sV, 0, s0, kvh, b);         // rows beyond S are clipped by the TMA unit
        tma_store_4d(&tm_dv, sV + KV_HALF, 64, s0, kvh, b);
        tma_store_4d(&tm_dk, sK, 0, s0, kvh, b);
        tma_store_4d(&tm_dk, sK + KV_HALF, 64, s0, kvh, b);
        bulk_commit();
        life(5);
        bulk_wait_read0();
        life(6);
      }
    }
    tc_fence_before();
  } else {
    // ===================== drain warps: dQ^T (TMEM, d on lanes) -> smem fp32 [64][128] -> bulk reduce-add =====================
    const int quarter = warp & 3;
    const int c = quarter * 32 + lane;                  // head-dim index owned by this thread
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    const int did = threadIdx.x - 576;                  // 0..127
    int h = kvh * G, qt = qt_first - 1;
    for (int i = 0; i < n; ++i) {
      const int st = i & 1;
      if (++qt == n_qt) {
        qt = qt_first;
        ++h;
      }
      const bool dtr = (warp == 18 && lane == 0);
      if (dtr) stamp(4, i, 0);
      // one key tile (S <= 128, the query pass): no other CTA adds to these dq rows, so they leave as bf16 from here —
      // dq = scale * dQ + dz * w[h] — and neither the fp32 accumulator, its seed, nor the conversion launch exist.
      // dz of the tile's 64 rows and this thread's w column are fetched before the wait.
      float dz_lo = 0.f, dz_hi = 0.f, w_c = 0.f;
      if (p.dq_direct && p.active) {
        const int64_t r0 = ((int64_t)b * p.H + h) * T + qt * BQ;
        if (qt * BQ + lane < T) dz_lo = p.dz_ws[r0 + lane];
        if (qt * BQ + 32 + lane < T) dz_hi = p.dz_ws[r0 + 32 + lane];
        w_c = p.w[h * HD + c];
      }
      mbar_wait(&bars[DQ_FULL + st], (i >> 1) & 1, 410 + st);
      if (dtr) stamp(4, i, 1);
      tc_fence_after();
      float v0[32], v1[32];
      tmem_ld32(tmem_base + lane_off + TM_DP + st * BQ, v0);
      tmem_ld32(tmem_base + lane_off + TM_DP + st * BQ + 32, v1);
      tc_wait_ld();
      tc_fence_before();
      mbar_arrive(&bars[DRAIN_DONE + st]);              // dP[st] may take dP^T(i+2) / dS^T(i+2)
      if (dtr) stamp(4, i, 2);
      if (dbg(256)) continue;
      if (p.dq_direct) {
        __nv_bfloat16* dqp = reinterpret_cast<__nv_bfloat16*>(p.dq) + (((int64_t)b * p.H + h) * T + qt * BQ) * HD + c;
        const int rows_left = T - qt * BQ;               // (warp-uniform)
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float s0 = __shfl_sync(0xffffffffu, dz_lo, j), s1 = __shfl_sync(0xffffffffu, dz_hi, j);
          if (j < rows_left) dqp[(int64_t)j * HD] = __float2bfloat16_rn(fmaf(p.dq_out_mul, v0[j], s0 * w_c));
          if (32 + j < rows_left) dqp[(int64_t)(32 + j) * HD] = __float2bfloat16_rn(fmaf(p.dq_out_mul, v1[j], s1 * w_c));
        }
        if (dtr) stamp(4, i, 3);
        continue;
      }
      // Each drain warp reduces its own 32-column slice of dQ on its own: four 16-row chunks through two 2 KB staging
      // buffers of the warp, one TMA tile reduce-add (fp32 add at L2) per chunk.  No barrier couples the four warps, and
      // while the TMA unit reads one buffer the lanes fill the other.  (Measured: skipping the waits or the fences changes nothing; the
      // drain as a whole costs ~10 % — its 64 KB of shared-memory traffic and the L2 adds, not its synchronisation.)  Rows past T are clipped by the TMA unit.
      float* wstage = stage + quarter * (2 * 16 * 32);
#pragma unroll
      for (int q4 = 0; q4 < 4; ++q4) {
        float* buf = wstage + (q4 & 1) * (16 * 32);
        if (lane == 0) bulk_wait_read1();                // the chunk that used this buffer two chunks ago has been read
        __syncwarp();
#pragma unroll
        for (int t = 0; t < 16; ++t) buf[t * 32 + lane] = (q4 < 2) ? v0[16 * q4 + t] : v1[16 * (q4 - 2) + t];
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          if (!dbg(1)) tma_reduce_add_3d(&tm_dq, buf, 32 * quarter, qt * BQ + 16 * q4, b * p.H + h);
          bulk_commit();                                 // one group per chunk keeps the group <-> buffer pairing regular
        }
      }
      if (dtr) stamp(4, i, 3);
    }
    if (lane == 0) bulk_wait0();
  }
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
    if (lane == 0) life(7);
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
  static const int dbg_env = [] { const char* e = getenv("MIMIC_DEBUG_FLAGS"); return e ? atoi(e) : 0; }();   // bit 0: skip the dQ reduce (timing experiments only)
  pp.debug_flags = dbg_env;
  static const bool no_direct = getenv("MIMIC_BWD_NO_DIRECT_DQ") != nullptr;   // A/B only
  // one key tile whose CTA visits EVERY query tile (with T > S the first query tiles of an inactive causal handle see no key
  // and are skipped: their dq rows only exist through the accumulator path), no outside gradient on lse_full (the seed
  // uses the op's own dz, dz_ws holds the combined one)
  pp.dq_direct = (p.S <= BN && (p.active || !p.causal || p.S >= p.T) && p.d_lse_full == nullptr && !no_direct) ? 1 : 0;
  if ((rc = simt::bwd_prep_bf16_d128(pp, st))) return rc;
  CUtensorMap tq, tk, tv, tdo, tdk, tdv, tdq;
  if ((rc = make_map_f32_rows(&tdq, p.dq_acc, (int64_t)p.B * p.H, p.T, 16))) return rc;
  if ((rc = make_map(&tq, p.q, p.B, p.H, p.T, p.q_sb, p.q_sh, p.q_st, BQ))) return rc;
  // dk, dv: contiguous [B, Hkv, S, d] (include/mimic_b200.h)
  if ((rc = make_map(&tdk, p.dk, p.B, p.Hkv, p.S, (int64_t)p.Hkv * p.S * HD, (int64_t)p.S * HD, HD, BN))) return rc;
  if ((rc = make_map(&tdv, p.dv, p.B, p.Hkv, p.S, (int64_t)p.Hkv * p.S * HD, (int64_t)p.S * HD, HD, BN))) return rc;
  if ((rc = make_map(&tk, p.k, p.B, p.Hkv, p.S, p.k_sb, p.k_sh, p.k_ss, BN))) return rc;
  if ((rc = make_map(&tv, p.v, p.B, p.Hkv, p.S, p.v_sb, p.v_sh, p.v_ss, BN))) return rc;
  if ((rc = make_map(&tdo, p.d_out, p.B, p.H, p.T, p.go_sb, p.go_sh, p.go_st, BQ))) return rc;
  static bool configured = false;
  if (!configured) {
    MIMIC_CUDA_OK(cudaFuncSetAttribute(attn_bwd_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    MIMIC_CUDA_OK(cudaFuncSetAttribute(attn_bwd_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    configured = true;
  }
  dim3 grid((p.S + BN - 1) / BN, p.Hkv, p.B);
  const int life = debug_trace_capacity() >= 2048 + 8 * (size_t)grid.x * grid.y * grid.z ? 1 : 0;
  if (pp.trace != nullptr || pp.debug_flags != 0) attn_bwd_tc_kernel<true><<<grid, THREADS, SMEM_BYTES, st>>>(tq, tk, tv, tdo, tdk, tdv, tdq, pp, life);
  else attn_bwd_tc_kernel<false><<<grid, THREADS, SMEM_BYTES, st>>>(tq, tk, tv, tdo, tdk, tdv, tdq, pp, 0);
  MIMIC_LAUNCH_OK("attn_bwd_tc_kernel");
  return pp.dq_direct ? MIMIC_OK : simt::convert_dq_bf16(pp, st);
}

}  // namespace tc
}  // namespace mimic
