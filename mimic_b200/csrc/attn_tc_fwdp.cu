// Shifted self-attention forward — ONE PERSISTENT CTA PER SM, two 128-row query tiles ping-ponged over 128-key tiles.
// Same math / masks / statistics / outputs as the other tcgen05 forwards; used for T > 128.
//
// Why this shape: with 128-key tiles every MMA is M=128, N=128 — QK^T (SS) and PV (TS) both sit on the
// 64-cycle floor of the tensor pipe (the 64-key kernels pay 48.6 instead of 32 cycles per SS MMA: operand fetch).  Two
// query tiles per CTA give the pipe the other tile's MMAs to run while one tile is in its softmax, every K/V tile
// fetched by TMA feeds two QK^T and two PV, and a persistent CTA pays barrier init / TMEM alloc / key-bit tables once.
// The item boundary is seamless: Q_x of the next item is requested as soon as the last QK^T_x of this item has
// completed (one step before the item ends), QK^T_x'(0) is issued right behind PV_x(last), and the normalise / store
// epilogue runs on its own warpgroup while the softmax warpgroups are already in the next item.
//
//   warps 0-3   softmax of query tile 0      (192 registers)       TMEM  S0 [0,128)    O0 [256,384)
//   warps 4-7   softmax of query tile 1      (192 registers)             S1 [128,256)  O1 [384,512)
//   warps 8-11  epilogue: O_x / l, + mu*shift, bf16, TMA store (72)      P_x = first 64 columns of S_x (bf16 pairs)
//   warp 12     MMA issuer (one elected lane issues)
//   warp 13 K producer | warp 14 V producer (one thread each, sleeping on the barrier it needs)
//   warp 15     item scheduler (global ticket counter -> shared-memory ring) + Q producer, one polling thread
//   smem: Q0 Q1 | K ring x2 | V ring x2 (32 KB tiles, 128B swizzle) | 16 KB output staging | barriers, stats, key bits
// MMA issue order (tcgen05 ops of one thread execute in order; P_x aliases S_x, so QK_x(j+1) follows PV_x(j)):
//   QK0(0) QK1(0) | per key tile j:  [P0(j)] PV0(j) QK0(j+1)   [P1(j)] PV1(j) QK1(j+1)
// P_x(j) is handed over in FOUR 32-key quarters (one barrier per quarter — a waiter must never fall two phases behind): the PV MMAs of the first quarters
// run while the softmax warps still exponentiate the later ones, which takes ~3/4 of PV_x(j) off the chain
// softmax_x(j) -> PV_x(j) -> QK_x(j+1) -> softmax_x(j+1) that bounds a tile's step.
#include <cuda.h>

#include <atomic>
#include <cstdlib>

#include "attn_common.cuh"
#include "tc_host.cuh"
#include "tc_ptx.cuh"
#include "tc_softmax.cuh"

namespace mimic {
namespace tc {

using namespace ptx;

namespace fp {
constexpr int BM = 128, BN = 128, HD = 128;
constexpr int HALF = 128 * 64 * 2;               // [128 rows][128 B]
constexpr int TILE = 2 * HALF;                   // 32 KB
constexpr int THREADS = 512;          // warps 14, 15 only take part in the register hand-over
constexpr float LOG2E = 1.4426950408889634f, LN2 = 0.6931471805599453f;
constexpr float RESCALE_THRESHOLD = 8.f;
constexpr int OFF_Q = 0, OFF_K = 2 * TILE, OFF_V = 4 * TILE, OFF_OST = 6 * TILE, OFF_BAR = OFF_OST + HALF;
constexpr int NSCHED = 3;                        // depth of the item ring between the scheduler thread and the consumer roles
constexpr int OFF_TMEM = OFF_BAR + 48 * 8, OFF_SCHED = OFF_TMEM + 16, OFF_STATS = OFF_SCHED + 16, OFF_SHIFT = OFF_STATS + 2 * 128 * 8,
              OFF_KBITS = OFF_SHIFT + HD * 4;
constexpr int KBITS_MAX = 16384;
constexpr int SMEM_MAX = OFF_KBITS + KBITS_MAX;
static_assert(SMEM_MAX <= 227 * 1024, "shared memory budget");
enum { Q_FULL = 0, Q_EMPTY = 2, K_FULL = 4, K_EMPTY = 6, V_FULL = 8, V_EMPTY = 10, S_FULL = 12, SOFT_DONE = 14,
       PV_DONE = 16, O_FULL = 18, O_EMPTY = 20, ST_FULL = 22, ST_EMPTY = 24, SCHED_FULL = 26, SCHED_EMPTY = 30, P_PART = 34, Z_FULL = 42 };
constexpr int SCHED_CONSUMERS = 8 + 4 + 1 + 3;   // softmax warps, epilogue warps, MMA warp, the K / V / Q producer threads
constexpr uint32_t IDESC_QK = idesc_bf16(BM, BN, 0, 0);
constexpr uint32_t IDESC_PV = idesc_bf16(BM, HD, 0, 1);
// 512 threads launch with 128 registers each (ptxas sizes a kernel that uses setmaxnreg by whole warpgroups); the epilogue
// warpgroup and the warpgroup of the two single-thread roles give registers back, the two softmax warpgroups take them
#ifndef FWDP_EXCL_POLY_MASK
#define FWDP_EXCL_POLY_MASK 0xA4A4u      // pairs 2, 5, 7 of every 8: three eighths of the exponentials of an exponent-sum-only step
#endif
#ifndef FWDP_REG_S
#define FWDP_REG_S 192
#define FWDP_REG_E 64
#define FWDP_REG_M 64
#endif
constexpr int REG_LAUNCH = 128;
constexpr int REG_SOFTMAX = FWDP_REG_S, REG_EPI = FWDP_REG_E, REG_MISC = FWDP_REG_M;
static_assert(2 * REG_SOFTMAX + REG_EPI + REG_MISC <= 4 * REG_LAUNCH, "register pool: 4 warpgroups x 128 threads x 128 registers");
inline int kbit_words(int S) { return ((S + BN - 1) / BN) * 4; }
}  // namespace fp

__device__ __forceinline__ void mbar_wait_q(uint64_t* bar, uint32_t parity) { mbar_wait(bar, parity, 0); }

template <int N>
__device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N>
__device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }

__device__ __forceinline__ uint32_t low_bitsp(int n) { return n >= 32 ? 0xFFFFFFFFu : (n <= 0 ? 0u : ((1u << n) - 1u)); }

constexpr int N_TICKETS = 256;
__device__ unsigned int g_fwdp_tickets[2 * N_TICKETS];   // zero-initialised; every launch leaves its pair at zero again

struct FwdItem {
  int h, b, kvh;
  int t0[2], jd[2], n[2];
  int nmax, nv;          // key tiles visited by the item (K loads), key tiles with a PV (V loads)
};
struct FwdCounts {       // what the MMA issuer needs of an item
  int jd[2], n[2], nmax, nv;
};

// DBG = true builds the instrumented twin (per-item %globaltimer stamps of CTA 0, scripts/trace_fwdp_items.py); the
// production instantiation carries none of it.
template <bool DBG>
__global__ void __launch_bounds__(fp::THREADS, 1)
attn_fwd_tcp_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                    const __grid_constant__ CUtensorMap tm_v, const __grid_constant__ CUtensorMap tm_o, AttnParams p,
                    int sec_hb, unsigned int* ticket) {
  using namespace fp;
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint8_t* sQ = smem + OFF_Q;
  uint8_t* sK = smem + OFF_K;
  uint8_t* sV = smem + OFF_V;
  uint8_t* sO = smem + OFF_OST;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + OFF_TMEM);
  volatile int* sched = reinterpret_cast<volatile int*>(smem + OFF_SCHED);   // item ring written by the scheduler thread
  float2* stats = reinterpret_cast<float2*>(smem + OFF_STATS);          // [2][128] (1 / l_c, mu)
  float* s_shift = reinterpret_cast<float*>(smem + OFF_SHIFT);          // shift[h] of the epilogue's current item
  uint32_t* kbits_all = reinterpret_cast<uint32_t*>(smem + OFF_KBITS);  // [B][words] when a key mask is given

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int T = p.T, S = p.S, off = p.S - p.T;
  const int n_tiles = (S + BN - 1) / BN;
  const int n_qp = (T + 2 * BM - 1) / (2 * BM);
  const int HB = p.H * p.B;
  const int n_items = n_qp * HB;
  const int G = p.H / p.Hkv;
  const int words = n_tiles * 4;
  const bool have_kbits = p.key_mask != nullptr;

  // item -> (query-tile pair, head, batch).  Order: sections of `sec_hb` (head, batch) streams whose K/V fit L2
  // together; inside a section the pair rank is slowest (descending work under the causal mask) and the head fastest
  // (a GQA group shares its K/V through L2).  Items are handed out DYNAMICALLY: the producer thread draws the CTA's next
  // item from a global ticket counter and publishes it in a small shared-memory ring that every role reads in order
  // (a static round-robin left SMs idle for 14 % of the kernel at T = 4096: the lists do not sum to equal work).
  auto decode = [&](int it) {
    FwdItem m;
    const int per_sec = n_qp * sec_hb;
    int sec = it / per_sec;
    const int nsec = (HB + sec_hb - 1) / sec_hb;
    if (sec >= nsec) sec = nsec - 1;
    const int hs = min(sec_hb, HB - sec * sec_hb);
    const int rem = it - sec * per_sec;
    const int qrank = rem / hs;
    const int hb = sec * sec_hb + (rem - qrank * hs);
    const int qp = n_qp - 1 - qrank;
    m.b = hb / p.H;
    m.h = hb - m.b * p.H;
    m.kvh = m.h / G;
#pragma unroll
    for (int x = 0; x < 2; ++x) {
      m.t0[x] = (qp * 2 + x) * BM;
      if (m.t0[x] >= T) {
        m.jd[x] = -1;
        m.n[x] = 0;
      } else {
        m.jd[x] = n_tiles - 1;
        if (p.causal) {
          const int smax = min(m.t0[x] + BM - 1, T - 1) + off;
          m.jd[x] = smax < 0 ? -1 : min(n_tiles - 1, smax / BN);
        }
        m.n[x] = p.active ? n_tiles : m.jd[x] + 1;
      }
    }
    m.nmax = max(m.n[0], m.n[1]);
    m.nv = max(m.jd[0], m.jd[1]) + 1;
    return m;
  };

  // k-th item of this CTA (every lane of the calling warp waits and reads), >= n_items when the work has run out
  auto sched_get = [&](uint32_t k) -> int {
    mbar_wait_q(&bars[SCHED_FULL + (k % NSCHED)], (k / NSCHED) & 1);
    return sched[k % NSCHED];
  };
  // the calling warp is done with its k-th item (one arrival per consumer warp)
  auto sched_release = [&](uint32_t k) {
    __syncwarp();
    if (lane == 0) mbar_arrive(&bars[SCHED_EMPTY + (k % NSCHED)]);
  };

  // ------------------------------------------------ prologue (once per CTA) ------------------------------------------------
  if (warp == 13) {
    if (lane == 0) {
      for (int i = 0; i < 2; ++i) {
        mbar_init(&bars[Q_FULL + i], 1);
        mbar_init(&bars[Q_EMPTY + i], 1 + 128);      // last QK^T of the item (commit) + the epilogue threads (z1 read)
        mbar_init(&bars[Z_FULL + i], 128);           // z1 of the tile's rows is in stats[].y (epilogue threads)
        mbar_init(&bars[K_FULL + i], 1);
        mbar_init(&bars[K_EMPTY + i], 1);
        mbar_init(&bars[V_FULL + i], 1);
        mbar_init(&bars[V_EMPTY + i], 1);
        mbar_init(&bars[S_FULL + i], 1);
        mbar_init(&bars[SOFT_DONE + i], 128);
        mbar_init(&bars[PV_DONE + i], 1);
        mbar_init(&bars[O_FULL + i], 1);
        mbar_init(&bars[O_EMPTY + i], 128);
        mbar_init(&bars[ST_FULL + i], 128);
        mbar_init(&bars[ST_EMPTY + i], 128);
        for (int q = 0; q < 4; ++q) mbar_init(&bars[P_PART + 4 * i + q], 128);
      }
      for (int i = 0; i < NSCHED; ++i) {
        mbar_init(&bars[SCHED_FULL + i], 1);
        mbar_init(&bars[SCHED_EMPTY + i], SCHED_CONSUMERS);
      }
      fence_barrier_init();
      prefetch_tmap(&tm_q);
      prefetch_tmap(&tm_k);
      prefetch_tmap(&tm_v);
      prefetch_tmap(&tm_o);
    }
    __syncwarp();
    tmem_alloc(tmem_slot, 512);
  } else if (warp < 12 && have_kbits) {
    // key-validity bits of every batch row, once per CTA: 8 independent byte loads in flight per lane
    for (int bb = 0; bb < p.B; ++bb) {
      const uint8_t* km = p.key_mask + (int64_t)bb * S;
      for (int w0 = warp; w0 < words; w0 += 12 * 8) {
        uint8_t ok[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int s = (w0 + 12 * u) * 32 + lane;
          ok[u] = (s < S) ? km[s] : (uint8_t)0;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int wi = w0 + 12 * u;
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
  // debug timeline: trace[(role * 16 + item ordinal) * 16 + event] in ns; roles 0 MMA issuer, 1 epilogue, 2 / 3 softmax tile 0 / 1
  auto stamp = [&](int role, uint32_t k, int ev) {
    if constexpr (DBG) {
      if (blockIdx.x == 0 && p.trace != nullptr && k < 16) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        p.trace[(role * 16 + k) * 16 + ev] = t;
      }
    }
  };

  if (warp == 13 || warp == 14) {
    // ===================== K producer (warp 13) / V producer (warp 14): one thread each, blocking waits =====================
    // (a 2-stage ring makes the issue time of the next load part of the critical path: each stream sleeps on exactly the
    // barrier it needs — a thread multiplexing the streams by polling added ~15 % to the kernel)
    reg_dec<REG_MISC>();
    if (lane == 0) {
      const bool is_k = warp == 13;
      uint8_t* ring = is_k ? sK : sV;
      const CUtensorMap* tm = is_k ? &tm_k : &tm_v;
      const int FULL = is_k ? K_FULL : V_FULL, EMPTY = is_k ? K_EMPTY : V_EMPTY;
      uint32_t g = 0;
      for (uint32_t k = 0;; ++k) {
        const int it = sched_get(k);
        if (it >= n_items) break;
        const FwdItem m = decode(it);
        const int n_loads = is_k ? m.nmax : m.nv;
        for (int j = 0; j < n_loads; ++j, ++g) {
          const int st = g & 1;
          mbar_wait_q(&bars[EMPTY + st], ((g >> 1) & 1) ^ 1);
          mbar_expect_tx(&bars[FULL + st], TILE);
          tma_load_4d(ring + st * TILE, tm, &bars[FULL + st], 0, j * BN, m.kvh, m.b);
          tma_load_4d(ring + st * TILE + HALF, tm, &bars[FULL + st], 64, j * BN, m.kvh, m.b);
        }
        mbar_arrive(&bars[SCHED_EMPTY + (k % NSCHED)]);
      }
    }
  } else if (warp == 15) {
    // ===================== scheduler + Q producer: one thread, polling (neither is latency-critical) =====================
    reg_dec<REG_MISC>();
    if (lane == 0) {
      uint32_t pub = 0, kQ = 0, qc0 = 0, qc1 = 0;          // items published (the last one is the end sentinel); Q-stream ordinal
      bool pub_done = false, hQ = false, eQ = false;
      int xQ = 0;
      FwdItem mQ = decode(0);
      uint32_t idle = 0;
      while (!eQ) {
        bool progressed = false;
        if (!pub_done) {
          const int slot = pub % NSCHED;
          if (pub < (uint32_t)NSCHED || mbar_test_wait(&bars[SCHED_EMPTY + slot], ((pub / NSCHED) & 1) ^ 1)) {
            const int it = (int)atomicAdd(ticket, 1u);
            sched[slot] = it;
            mbar_arrive(&bars[SCHED_FULL + slot]);           // release: the store above is visible to the waiters
            ++pub;
            if (it >= n_items) pub_done = true;
            progressed = true;
          }
        }
        if (!hQ) {
          if (kQ < pub) {
            const int it = sched[kQ % NSCHED];
            if (it >= n_items) eQ = true; else { mQ = decode(it); hQ = true; xQ = 0; }
            progressed = true;
          }
        } else if (xQ >= 2) {
          mbar_arrive(&bars[SCHED_EMPTY + (kQ % NSCHED)]);
          hQ = false;
          ++kQ;
          progressed = true;
        } else if ((xQ ? mQ.n[1] : mQ.n[0]) == 0) {
          ++xQ;                                              // tile without any key tile to visit: no Q needed
          progressed = true;
        } else {
          const uint32_t qc = xQ ? qc1 : qc0;
          if (mbar_test_wait(&bars[Q_EMPTY + xQ], (qc & 1) ^ 1)) {
            mbar_expect_tx(&bars[Q_FULL + xQ], TILE);
            const int tq0 = xQ ? mQ.t0[1] : mQ.t0[0];
            tma_load_4d(sQ + xQ * TILE, &tm_q, &bars[Q_FULL + xQ], 0, tq0, mQ.h, mQ.b);        // rows >= T arrive as zeros
            tma_load_4d(sQ + xQ * TILE + HALF, &tm_q, &bars[Q_FULL + xQ], 64, tq0, mQ.h, mQ.b);
            if (xQ) ++qc1; else ++qc0;
            ++xQ;
            progressed = true;
          }
        }
        if (progressed) {
          idle = 0;
        } else {
          __nanosleep(32);
          if (++idle > (1u << 24)) __trap();                   // a protocol bug must not hang the GPU box
        }
      }
      // the last CTA to run out of work re-arms the ticket counter for the next launch that uses it
      __threadfence();
      if (atomicAdd(ticket + 1, 1u) == gridDim.x - 1) {
        ticket[0] = 0u;
        ticket[1] = 0u;
        __threadfence();
      }
    }
  } else if (warp == 12) {
    // ===================== MMA issuer (warp-uniform control flow, one elected lane issues) =====================
    reg_dec<REG_MISC>();
    constexpr uint32_t HI = desc_hi_sw128(1024);
    const uint32_t q_lo0 = desc_lo(smem_u32(sQ), 16), k_lo0 = desc_lo(smem_u32(sK), 16);
    const uint32_t v_lo0 = desc_lo(smem_u32(sV), HALF);                 // V as MN-major B: d halves 16 KB apart
    uint32_t sc0 = 0, sc1 = 0, qc0 = 0, qc1 = 0, oc0 = 0, oc1 = 0, pp0 = 0, pp1 = 0;
    uint32_t kb = 0, vb = 0;                                            // ring index of the item's first K / V tile
    bool pre0 = false, pre1 = false;
    auto counts = [&](int it) {
      const FwdItem f = decode(it);
      FwdCounts c;
      c.jd[0] = f.jd[0]; c.jd[1] = f.jd[1]; c.n[0] = f.n[0]; c.n[1] = f.n[1]; c.nmax = f.nmax; c.nv = f.nv;
      return c;
    };
    // S_x = Q_x K^T for key tile j of an item with counts (n_x, n_1); ring index kbase + j
    auto issue_qk = [&](int x, int n_x, int n_1, int j, uint32_t kbase) {
      if (j == 0) {
        mbar_wait_q(&bars[Q_FULL + x], (x ? qc1 : qc0) & 1);
        if (x) ++qc1; else ++qc0;
      }
      const uint32_t g = kbase + j;
      const int ks = g & 1;
      mbar_wait_q(&bars[K_FULL + ks], (g >> 1) & 1);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t q_lo = q_lo0 + x * (TILE >> 4), k_lo = k_lo0 + ks * (TILE >> 4);
        const uint32_t d_s = tmem_base + x * BN;
#pragma unroll
        for (int k = 0; k < HD / 16; ++k) {
          const uint32_t koff = ((k >> 2) * HALF + (k & 3) * 32) >> 4;
          if (k == 0) mma_ss_c<false>(d_s, desc64(HI, q_lo + koff), desc64(HI, k_lo + koff), IDESC_QK);
          else mma_ss_c<true>(d_s, desc64(HI, q_lo + koff), desc64(HI, k_lo + koff), IDESC_QK);
        }
        if (x == 1 || j >= n_1) tc_commit(&bars[K_EMPTY + ks]);           // tile 1 issues after tile 0: last reader of K(j)
        tc_commit(&bars[S_FULL + x]);
        if (j == n_x - 1) tc_commit(&bars[Q_EMPTY + x]);                  // no later MMA reads Q_x of this item
      }
      __syncwarp();
    };
    int it = sched_get(0);
    FwdCounts m = counts(it < n_items ? it : 0);
    for (uint32_t k = 0; it < n_items; ++k) {
      // the next item is only needed from the round in which the first tile finishes (its QK^T'(0) goes out behind the
      // tile's last PV): fetch it there, so the scheduler has the whole item to publish it
      int it_next = n_items;
      bool has_next = false, have_nx = false;
      FwdCounts nx = m;
      if (lane == 0) stamp(0, k, 0);
      if (m.n[0] > 0 && !pre0) issue_qk(0, m.n[0], m.n[1], 0, kb);
      if (m.n[1] > 0 && !pre1) issue_qk(1, m.n[1], m.n[1], 0, kb);
      pre0 = pre1 = false;
      if (lane == 0) stamp(0, k, 1);
      for (int j = 0; j < m.nmax; ++j) {
        if (!have_nx && j >= m.nmax - 2) {
          it_next = sched_get(k + 1);
          has_next = it_next < n_items;
          if (has_next) nx = counts(it_next);
          have_nx = true;
        }
#pragma unroll
        for (int x = 0; x < 2; ++x) {
          if (j >= m.n[x]) continue;
          const uint32_t g = vb + j;
          const int vs = g & 1;
          if (j <= m.jd[x]) {
            // what PV_x(j) needs besides P — the V tile and, for the first PV of an item, a drained O_x — is confirmed BEFORE
            // the softmax hand-over, so that nothing but the issue itself follows it (neither depends on this softmax step)
            if (j == 0) {
              if (lane == 0) stamp(0, k, 2 + 2 * x);
              mbar_wait_q(&bars[O_EMPTY + x], ((x ? oc1 : oc0) & 1) ^ 1);
              if (lane == 0) stamp(0, k, 3 + 2 * x);
            }
            mbar_wait_q(&bars[V_FULL + vs], (g >> 1) & 1);
          }
          if (j <= m.jd[x]) {
            // A: P_x in TMEM (128 keys = 64 packed columns, 8 per K step); B: V_j [128 keys][d], 16 keys = 2 KB per step.
            // Quarter q of P_x (keys 32 q .. 32 q + 31) arrives with phase q of the step on P_PART; the last one also means
            // "S_x is consumed" (there is no separate SOFT_DONE arrival on a step that has a PV).
            const uint32_t v_lo = v_lo0 + vs * (TILE >> 4);
            const uint32_t a_p = tmem_base + x * BN, d_o = tmem_base + 256 + x * BN;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              mbar_wait_q(&bars[P_PART + 4 * x + q], (x ? pp1 : pp0) & 1);
              if (j == 0 && q == 0 && lane == 0) stamp(0, k, 6 + x);
              if (j == m.jd[x] && q == 3 && lane == 0) stamp(0, k, 13 + x);
              tc_fence_after();
              if (elect_one()) {
#pragma unroll
                for (int k = 2 * q; k < 2 * q + 2; ++k)
                  mma_ts(d_o, a_p + k * 8, desc64(HI, v_lo + k * (2048 >> 4)), IDESC_PV, (j > 0 || k > 0) ? 1u : 0u);
                if (q == 3) {
                  if (x == 1 || j > m.jd[1]) tc_commit(&bars[V_EMPTY + vs]);
                  tc_commit(&bars[PV_DONE + x]);
                  if (j == m.jd[x]) tc_commit(&bars[O_FULL + x]);
                }
              }
              __syncwarp();
            }
            if (x) ++pp1; else ++pp0;
            if (j == m.jd[x]) {
              if (x) ++oc1; else ++oc0;
            }
          } else {
            mbar_wait_q(&bars[SOFT_DONE + x], (x ? sc1 : sc0) & 1);          // a step without PV: S_x has been consumed
            if (x) ++sc1; else ++sc0;
            tc_fence_after();
          }
          if (j + 1 < m.n[x]) {
            issue_qk(x, m.n[x], m.n[1], j + 1, kb);
          } else if (has_next && nx.n[x] > 0) {
            if (lane == 0) stamp(0, k, 9 + 2 * x);
            issue_qk(x, nx.n[x], nx.n[1], 0, kb + m.nmax);               // next item's first scores, behind PV_x(last)
            if (lane == 0) stamp(0, k, 10 + 2 * x);
            if (x) pre1 = true; else pre0 = true;
          }
        }
      }
      if (!have_nx) {                                        // item without any key tile
        it_next = sched_get(k + 1);
        if (it_next < n_items) nx = counts(it_next);
      }
      kb += m.nmax;
      vb += m.nv;
      m = nx;
      it = it_next;
      if (lane == 0) stamp(0, k, 8);
      sched_release(k);
    }
  } else if (warp >= 8) {
    // ===================== epilogue warpgroup: O_x / l_c + mu * shift -> bf16 -> staging -> TMA store =====================
    reg_dec<REG_EPI>();
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    const bool issuer = threadIdx.x == 256;
    uint32_t oc[2] = {0, 0}, stc[2] = {0, 0}, qce[2] = {0, 0};
    for (uint32_t k = 0;; ++k) {
      const int it = sched_get(k);
      if (it >= n_items) break;
      const FwdItem m = decode(it);
      if (issuer) stamp(1, k, 0);
      // shift[h] into shared memory while the item is still being computed (the epilogue proper only starts when the
      // softmax warps publish the row statistics at the item's end; eight dependent global loads per 32-column chunk
      // inside it cost ~2.4 us per tile — the MMA issuer sat that long in front of the next item's first PV).  The last
      // named barrier of the previous item lies behind every read of the row.
      s_shift[r] = p.active ? __ldg(p.shift + m.h * HD + r) : 0.f;
      // z1 = q_t . w[h] + bias[h] of both query tiles — only needed for mu at the item's END — is formed HERE, by the
      // warpgroup that idles while the item is computed: on the softmax warps it cost ~2 us of every item (they are the
      // busiest resource of the kernel: anything on them is on the critical path).  It travels in the .y slot of the row's
      // statistics entry (free until the softmax thread writes (1 / l, mu) there at the item's end).
#pragma unroll
      for (int x = 0; x < 2; ++x) {
        if ((x ? m.n[1] : m.n[0]) == 0) continue;          // (no Q load for such a tile: same rule as the Q producer)
        mbar_wait_q(&bars[Q_FULL + x], qce[x] & 1);
        ++qce[x];
        if (p.active) {
          const float4* w4 = reinterpret_cast<const float4*>(p.w + m.h * HD);
          const uint8_t* sQx = sQ + x * TILE;
          float za = 0.f, zb = 0.f, zc = 0.f, zd = 0.f;      // four independent chains
#pragma unroll 1
          for (int hf = 0; hf < 2; ++hf)
#pragma unroll 2
            for (int c = 0; c < 8; ++c) {                    // (64 registers here: two chunks in flight, not sixteen)
              const uint4 u = *reinterpret_cast<const uint4*>(sQx + hf * HALF + r * 128 + ((c ^ (r & 7)) << 4));
              const __nv_bfloat162* hh = reinterpret_cast<const __nv_bfloat162*>(&u);
              const float4 wa = __ldg(w4 + (hf * 64 + c * 8) / 4), wb = __ldg(w4 + (hf * 64 + c * 8) / 4 + 1);
              const float2 f0 = __bfloat1622float2(hh[0]), f1 = __bfloat1622float2(hh[1]);
              const float2 f2 = __bfloat1622float2(hh[2]), f3 = __bfloat1622float2(hh[3]);
              za = fmaf(f0.x, wa.x, za); zb = fmaf(f0.y, wa.y, zb); zc = fmaf(f1.x, wa.z, zc); zd = fmaf(f1.y, wa.w, zd);
              za = fmaf(f2.x, wb.x, za); zb = fmaf(f2.y, wb.y, zb); zc = fmaf(f3.x, wb.z, zc); zd = fmaf(f3.y, wb.w, zd);
            }
          stats[x * 128 + r].y = (za + zb) + (zc + zd) + __ldg(p.bias + m.h);
          mbar_arrive(&bars[Z_FULL + x]);
        }
        mbar_arrive(&bars[Q_EMPTY + x]);                   // this thread no longer reads sQ_x
      }
#pragma unroll
      for (int x = 0; x < 2; ++x) {
        if (m.t0[x] >= T) continue;
        const bool has_pv = m.jd[x] >= 0;
        mbar_wait_q(&bars[ST_FULL + x], stc[x] & 1);
        if (issuer) stamp(1, k, 1 + 4 * x);
        const float2 st = stats[x * 128 + r];
        mbar_arrive(&bars[ST_EMPTY + x]);
        ++stc[x];
        if (has_pv) {
          mbar_wait_q(&bars[O_FULL + x], oc[x] & 1);
          tc_fence_after();
        }
        if (issuer) stamp(1, k, 2 + 4 * x);
        const float inv = st.x, mu = st.y;
        const uint32_t t_O = tmem_base + lane_off + 256 + x * BN;
        const float4* sh4 = reinterpret_cast<const float4*>(s_shift);
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          if (issuer) bulk_wait_read0();                   // the previous store has finished reading the staging tile
          named_bar_sync(2, 128);
          uint8_t* srow = sO + r * 128;                    // 128B-swizzled [128 rows][64 cols] box, as the TMA map expects
#pragma unroll
          for (int c32 = 0; c32 < 2; ++c32) {
            float o[32];
            if (has_pv) {
              tmem_ld32(t_O + hf * 64 + c32 * 32, o);
              tc_wait_ld();
            } else {
#pragma unroll
              for (int c = 0; c < 32; ++c) o[c] = 0.f;
            }
            if (hf == 1 && c32 == 1 && has_pv) {           // O_x is in registers: the next item's PV_x(0) may overwrite it
              tc_fence_before();
              mbar_arrive(&bars[O_EMPTY + x]);
              if (issuer) stamp(1, k, 3 + 4 * x);
            }
#pragma unroll
            for (int c8 = 0; c8 < 4; ++c8) {
              const float4 s0 = sh4[(hf * 64 + c32 * 32 + c8 * 8) / 4], s1 = sh4[(hf * 64 + c32 * 32 + c8 * 8) / 4 + 1];
              uint32_t pk[4];
              pk[0] = pack_bf16(fmaf(o[8 * c8 + 0], inv, mu * s0.x), fmaf(o[8 * c8 + 1], inv, mu * s0.y));
              pk[1] = pack_bf16(fmaf(o[8 * c8 + 2], inv, mu * s0.z), fmaf(o[8 * c8 + 3], inv, mu * s0.w));
              pk[2] = pack_bf16(fmaf(o[8 * c8 + 4], inv, mu * s1.x), fmaf(o[8 * c8 + 5], inv, mu * s1.y));
              pk[3] = pack_bf16(fmaf(o[8 * c8 + 6], inv, mu * s1.z), fmaf(o[8 * c8 + 7], inv, mu * s1.w));
              const int chunk = c32 * 4 + c8;                // 16-byte chunk index inside the 128-byte row
              *reinterpret_cast<uint4*>(srow + ((chunk ^ (r & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            }
          }
          fence_proxy_async();                             // generic-proxy smem writes -> visible to the TMA (async proxy)
          named_bar_sync(2, 128);
          if (issuer) {
            tma_store_4d(&tm_o, sO, hf * 64, m.t0[x], m.h, m.b);     // rows >= T are clipped by the TMA unit
            bulk_commit();
          }
        }
        if (has_pv) ++oc[x];
        if (issuer) stamp(1, k, 4 + 4 * x);
      }
      sched_release(k);
    }
    if (issuer) bulk_wait0();
  } else {
    // ===================== softmax warpgroup x: one query row per thread =====================
    reg_inc<REG_SOFTMAX>();
    const int x = warp >> 2;
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    const uint32_t t_S = tmem_base + lane_off + x * BN, t_O = tmem_base + lane_off + 256 + x * BN;
    const float scale2 = p.scale * LOG2E;
    const uint8_t* sQx = sQ + x * TILE;
    uint32_t sc = 0, pvc = 0, zc = 0, stc = 0;

    for (uint32_t k = 0;; ++k) {
      const int it = sched_get(k);
      if (it >= n_items) break;
      const FwdItem m = decode(it);
      const int t0 = x ? m.t0[1] : m.t0[0];               // (no runtime-indexed arrays: they would live in local memory)
      if (t0 >= T) {                                      // tile outside the tensor: nothing to load, compute or store
        sched_release(k);
        continue;
      }
      const int nloc = x ? m.n[1] : m.n[0], jd = x ? m.jd[1] : m.jd[0], h = m.h, b = m.b;
      const int t = t0 + r;
      const bool row_ok = t < T;
      const uint32_t* kbits = have_kbits ? kbits_all + b * words : nullptr;

      float z1 = 0.f;
      const bool tr_thread = r == 0;
      if (tr_thread) stamp(2 + x, k, 0);
      float m_c = -INFINITY, l_c = 0.f, m_x = -INFINITY, l_x = 0.f;   // log2 domain
      // Per-step tile class, decided with WARP-UNIFORM arithmetic on the warp's 32 rows [tw0, tw0 + 31] — the mask-free
      // classes never touch a per-thread mask word (the word / vote version of this cost ~430 cycles of every step):
      //   FULL   every key of the tile visible to every row of the warp           -> max, exp2, P        (the bulk)
      //   EXCL   active handle, every key in range and hidden from every row       -> row-sum of exp2 only (above the diagonal)
      //   ZERO   inactive handle, nothing visible to the warp, tile still has a PV -> P = 0
      //   MIXED  diagonal / padded / ragged tiles                                  -> two chunked passes with mask words
      const int tw0 = t0 + quarter * 32;
      const bool rows_all_ok = tw0 + 31 < T;
      const int lim_lo = tw0 + off, lim_hi = min(tw0 + 31, T - 1) + off;   // last visible key of the warp's first / last row
      for (int j = 0; j < nloc; ++j) {
        const int key0 = j * BN;
        const bool in_range_all = key0 + BN <= S;
        bool keys_all_ok = in_range_all;
        if (kbits != nullptr) {
          const uint4 kw = *reinterpret_cast<const uint4*>(kbits + j * 4);
          keys_all_ok = (kw.x & kw.y & kw.z & kw.w) == 0xFFFFFFFFu;
        }
        const bool full_vis = !p.causal || key0 + BN - 1 <= lim_lo;
        const bool none_vis = (p.causal && key0 > lim_hi) || tw0 >= T;
        const bool has_pv = j <= jd;
        const bool c_full = rows_all_ok && keys_all_ok && full_vis;
        const bool c_excl = p.active && none_vis && in_range_all && rows_all_ok;
        const bool c_zero = !p.active && none_vis;
        mbar_wait_q(&bars[S_FULL + x], sc & 1);
        if (tr_thread && j == 0) stamp(2 + x, k, 3);
        if (tr_thread && j == 1) stamp(2 + x, k, 4);
        if (tr_thread && j == nloc - 1) stamp(2 + x, k, 9);
        ++sc;
        tc_fence_after();
        if (c_full) {
          float s[128];
          tmem_ld32(t_S, *reinterpret_cast<float(*)[32]>(s));
          tmem_ld32(t_S + 32, *reinterpret_cast<float(*)[32]>(s + 32));
          tmem_ld32(t_S + 64, *reinterpret_cast<float(*)[32]>(s + 64));
          tmem_ld32(t_S + 96, *reinterpret_cast<float(*)[32]>(s + 96));
          tc_wait_ld();
          const float m_new = fmaxf(m_c, row_max128(s) * scale2);
          const bool need = (m_c != -INFINITY) && (m_new > m_c + RESCALE_THRESHOLD);
          if (m_c == -INFINITY) m_c = m_new;
          if (__any_sync(0xffffffffu, need)) {            // lazy rescale of O and l_c (warp-collective TMEM ops); rare
            const float alpha = need ? ex2(m_c - m_new) : 1.f;
            if (j > 0) {
              mbar_wait_q(&bars[PV_DONE + x], (pvc - 1) & 1);
              tc_fence_after();
#pragma unroll 1
              for (int i = 0; i < 8; ++i) {
                float o[16];
                tmem_ld16(t_O + 16 * i, o);
                tc_wait_ld();
                uint32_t ou[16];
#pragma unroll
                for (int c = 0; c < 16; ++c) ou[c] = __float_as_uint(o[c] * alpha);
                tmem_st16(t_O + 16 * i, ou);
              }
            }
            if (need) {
              l_c *= alpha;
              m_c = m_new;
            }
          }
#pragma unroll
          for (int i = 0; i < 4; ++i) {                          // P in four quarters, each handed to the MMA issuer as it lands
            l_c += exp_pack_store32(*reinterpret_cast<const float(*)[32]>(s + 32 * i), scale2, -m_c, t_S + 16 * i);
            tc_wait_st();
            tc_fence_before();
            mbar_arrive(&bars[P_PART + 4 * x + i]);
          }
        } else if (c_excl) {
          float s[128];
          tmem_ld32(t_S, *reinterpret_cast<float(*)[32]>(s));
          tmem_ld32(t_S + 32, *reinterpret_cast<float(*)[32]>(s + 32));
          tmem_ld32(t_S + 64, *reinterpret_cast<float(*)[32]>(s + 64));
          tmem_ld32(t_S + 96, *reinterpret_cast<float(*)[32]>(s + 96));
          tc_wait_ld();
          if (!has_pv) {
            // The scores are in registers and this step writes no P: S_x is free NOW — QK_x(j+1) runs under this step's
            // exponentials instead of behind them.  (Above the diagonal of an active handle every step is of this kind:
            // 1.8 us per pair of tiles for 0.55 us of MMA before the early release.)
            tc_fence_before();
            mbar_arrive(&bars[SOFT_DONE + x]);
          }
          const float mx_new = fmaxf(m_x, row_max128(s) * scale2);
          const float carry = (m_x == -INFINITY) ? 0.f : l_x * ex2(m_x - mx_new);
          // three eighths of these exponentials go to the FMA pipe (expsum32_poly): measured 0.56 -> 0.58-0.59 of burst at
          // T = 4096 and 0.59-0.60 -> 0.61-0.63 at T = 8192; 1/4: less, 1/2: the same, 5/8: slower than none
          constexpr uint32_t PM = FWDP_EXCL_POLY_MASK;
          l_x = carry + (expsum32_poly<PM>(*reinterpret_cast<const float(*)[32]>(s), scale2, -mx_new) +
                         expsum32_poly<PM>(*reinterpret_cast<const float(*)[32]>(s + 32), scale2, -mx_new)) +
                (expsum32_poly<PM>(*reinterpret_cast<const float(*)[32]>(s + 64), scale2, -mx_new) +
                 expsum32_poly<PM>(*reinterpret_cast<const float(*)[32]>(s + 96), scale2, -mx_new));
          m_x = mx_new;
          if (has_pv) {                                    // (cannot happen above the diagonal of this tile; kept for safety)
            uint32_t z[16];
#pragma unroll
            for (int c = 0; c < 16; ++c) z[c] = 0u;
#pragma unroll
            for (int i = 0; i < 4; ++i) tmem_st16(t_S + 16 * i, z);
            tc_wait_st();
            tc_fence_before();
#pragma unroll
            for (int i = 0; i < 4; ++i) mbar_arrive(&bars[P_PART + 4 * x + i]);
          }
        } else if (c_zero) {
          if (has_pv) {                                    // another warp of the tile has visible keys: this one contributes P = 0
            uint32_t z[16];
#pragma unroll
            for (int c = 0; c < 16; ++c) z[c] = 0u;
#pragma unroll
            for (int i = 0; i < 4; ++i) tmem_st16(t_S + 16 * i, z);
            tc_wait_st();
            tc_fence_before();
#pragma unroll
            for (int i = 0; i < 4; ++i) mbar_arrive(&bars[P_PART + 4 * x + i]);
          }
        } else {
          // ---- MIXED (diagonal / padded / ragged tiles): two passes over 32-column chunks re-read from TMEM, each chunk
          // classified warp-uniformly so that only chunks that really mix visible, excluded and absent keys pay for mask
          // words — on a causal diagonal that is ONE chunk of four per warp:
          //   0 every key visible to every row   1 (active) every key excluded for every row
          //   2 contributes nothing (hidden + inactive, or past S, or no valid row)   3 partial: masked arithmetic
          auto chunk_class = [&](int i, uint32_t& kw) -> int {
            const int c0 = key0 + 32 * i;
            const bool none_in = c0 >= S, in_all = c0 + 32 <= S;
            kw = kbits ? kbits[j * 4 + i] : low_bitsp(S - c0);
            if (rows_all_ok && kw == 0xFFFFFFFFu && (!p.causal || c0 + 31 <= lim_lo)) return 0;
            const bool hid_all = tw0 >= T || none_in || (p.causal && c0 > lim_hi);
            if (hid_all) {
              if (none_in || tw0 >= T || !p.active) return 2;
              if (in_all && rows_all_ok) return 1;
            }
            return 3;
          };
          float mv = -INFINITY, mx = -INFINITY;
#pragma unroll 1
          for (int i = 0; i < 4; ++i) {
            uint32_t kw;
            const int cls = chunk_class(i, kw);
            if (cls == 2) continue;
            float s[32];
            tmem_ld32(t_S + 32 * i, s);
            tc_wait_ld();
            if (cls == 0) {
              mv = fmaxf(mv, max32(s));
            } else if (cls == 1) {
              mx = fmaxf(mx, max32(s));
            } else {
              const uint32_t inr = low_bitsp(S - key0 - 32 * i);
              const uint32_t cz = p.causal ? low_bitsp(t + off - key0 - 32 * i + 1) : 0xFFFFFFFFu;
              const uint32_t vw = row_ok ? (kw & cz) : 0u;
              const uint32_t xw = (p.active && row_ok) ? (inr & ~vw) : 0u;
#pragma unroll
              for (int c = 0; c < 32; ++c) {
                mv = ((vw >> c) & 1u) ? fmaxf(mv, s[c]) : mv;
                mx = ((xw >> c) & 1u) ? fmaxf(mx, s[c]) : mx;
              }
            }
          }
          mv *= scale2;
          mx *= scale2;
          if (has_pv) {
            const float m_new = fmaxf(m_c, mv);
            const bool need = (m_c != -INFINITY) && (m_new > m_c + RESCALE_THRESHOLD);
            if (m_c == -INFINITY) m_c = m_new;
            if (__any_sync(0xffffffffu, need)) {
              const float alpha = need ? ex2(m_c - m_new) : 1.f;
              if (j > 0) {
                mbar_wait_q(&bars[PV_DONE + x], (pvc - 1) & 1);
                tc_fence_after();
#pragma unroll 1
                for (int i = 0; i < 8; ++i) {
                  float o[16];
                  tmem_ld16(t_O + 16 * i, o);
                  tc_wait_ld();
                  uint32_t ou[16];
#pragma unroll
                  for (int c = 0; c < 16; ++c) ou[c] = __float_as_uint(o[c] * alpha);
                  tmem_st16(t_O + 16 * i, ou);
                }
              }
              if (need) {
                l_c *= alpha;
                m_c = m_new;
              }
            }
          }
          const float mx_new = fmaxf(m_x, mx);
          if (mx_new != -INFINITY) {
            l_x = (m_x == -INFINITY) ? 0.f : l_x * ex2(m_x - mx_new);
            m_x = mx_new;
          }
          const float ref_v = (m_c == -INFINITY) ? 0.f : -m_c, ref_x = (m_x == -INFINITY) ? 0.f : -m_x;
          float rowsum = 0.f, es = 0.f;
#pragma unroll 1
          for (int i = 0; i < 4; ++i) {
            uint32_t kw;
            const int cls = chunk_class(i, kw);
            if (cls == 2) {
              if (has_pv) {                                  // P = 0 for this chunk (another chunk / warp of the tile has a PV)
                uint32_t z[16];
#pragma unroll
                for (int c = 0; c < 16; ++c) z[c] = 0u;
                tmem_st16(t_S + 16 * i, z);
                tc_wait_st();
                tc_fence_before();
                mbar_arrive(&bars[P_PART + 4 * x + i]);
              }
              continue;
            }
            // P chunk i (16 packed columns) lies inside score chunks <= i: always LOAD the scores of chunk i before storing P
            float s[32];
            tmem_ld32(t_S + 32 * i, s);
            tc_wait_ld();
            if (cls == 1 && has_pv) {
              uint32_t z[16];
#pragma unroll
              for (int c = 0; c < 16; ++c) z[c] = 0u;
              tmem_st16(t_S + 16 * i, z);
            }
            if (cls == 0) {
              if (has_pv) rowsum += exp_pack_store32(s, scale2, ref_v, t_S + 16 * i);
            } else if (cls == 1) {
              es += expsum32(s, scale2, ref_x);
            } else {
              const uint32_t inr = low_bitsp(S - key0 - 32 * i);
              const uint32_t cz = p.causal ? low_bitsp(t + off - key0 - 32 * i + 1) : 0xFFFFFFFFu;
              const uint32_t vw = row_ok ? (kw & cz) : 0u;
              const uint32_t xw = (p.active && row_ok) ? (inr & ~vw) : 0u;
              uint32_t pk[16];
#pragma unroll
              for (int c2 = 0; c2 < 16; ++c2) {
                const bool v0 = (vw >> (2 * c2)) & 1u, v1 = (vw >> (2 * c2 + 1)) & 1u;
                const bool x0 = (xw >> (2 * c2)) & 1u, x1 = (xw >> (2 * c2 + 1)) & 1u;
                const float e0 = ex2(fmaf(s[2 * c2], scale2, v0 ? ref_v : ref_x));
                const float e1 = ex2(fmaf(s[2 * c2 + 1], scale2, v1 ? ref_v : ref_x));
                const float p0 = v0 ? e0 : 0.f, p1 = v1 ? e1 : 0.f;
                rowsum += p0 + p1;
                es += (x0 ? e0 : 0.f) + (x1 ? e1 : 0.f);
                pk[c2] = pack_bf16(p0, p1);
              }
              if (has_pv) tmem_st16(t_S + 16 * i, pk);
            }
            if (has_pv) {                                    // quarter i of P is in place (stored above whatever its class)
              tc_wait_st();
              tc_fence_before();
              mbar_arrive(&bars[P_PART + 4 * x + i]);
            }
          }
          l_c += rowsum;
          l_x += es;
        }
        if (has_pv) {
          ++pvc;                                             // the four P_PART arrivals of the step stand for SOFT_DONE
        } else if (!c_excl || c_full) {                      // (an exponent-sum-only step released S_x right after loading it)
          tc_fence_before();
          mbar_arrive(&bars[SOFT_DONE + x]);
        }
      }

      // ---- item end: row statistics -> epilogue warpgroup (shared memory) and global ----
      if (tr_thread) stamp(2 + x, k, 5);
      const float inv = l_c > 0.f ? 1.f / l_c : 0.f;
      const float lse_c = l_c > 0.f ? (m_c + log2f(l_c)) * LN2 : -INFINITY;
      float mu = 0.f, lse_full = 0.f;
      if (p.active && nloc > 0) {                          // z1 of this row, formed by the epilogue warpgroup at the item's start
        mbar_wait_q(&bars[Z_FULL + x], zc & 1);
        ++zc;
        z1 = stats[x * 128 + r].y;
      }
      if (p.active) {
        const float lse_x = l_x > 0.f ? (m_x + log2f(l_x)) * LN2 : -INFINITY;
        const float hi = fmaxf(lse_c, lse_x), lo = fminf(lse_c, lse_x);
        lse_full = (lo == -INFINITY) ? hi : hi + log1pf(__expf(lo - hi));
        mu = 1.f / (1.f + __expf(lse_full - z1));
      }
      mbar_wait_q(&bars[ST_EMPTY + x], (stc & 1) ^ 1);   // the epilogue has read the previous item's statistics
      if (tr_thread) stamp(2 + x, k, 6);
      stats[x * 128 + r] = make_float2(inv, mu);
      mbar_arrive(&bars[ST_FULL + x]);
      ++stc;
      if (row_ok) {
        const int64_t ridx = ((int64_t)b * p.H + h) * T + t;
        p.lse_c[ridx] = lse_c;
        if (p.active) {
          p.lse_full[ridx] = lse_full;
          p.mu[ridx] = mu;
        }
      }
      if (tr_thread) stamp(2 + x, k, 7);
      sched_release(k);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 13) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

bool fwdp_supports(const AttnParams& p, bool forced) {
  using namespace fp;
  // Measured on B200 (scripts/ab_fwd.py, family 6 vs 3, TFLOP/s, B = 16384 / T, H = 32): active handle 372 / 395 at T = 512,
  // 563 / 546 at 1024, 748 / 726 at 2048, 893 / 857 at 4096, 963 / 894 at 8192; inactive 491 / 534 at 1024, 684 / 707 at 2048,
  // 871 / 855 at 4096, 1052 / 957 at 8192, 756 / 727 at the Idefics2 context pass (S = 3200, GQA 32:8, no mask).  Few key
  // tiles per item: the item hand-over weighs more than the 128-key MMAs save, and an inactive handle has half the tiles.
  if (p.T <= BM || (!forced && p.T < (p.active ? 1024 : 3072))) return false;   // (one query tile: the single-tile kernels)
  if (p.key_mask != nullptr && (size_t)p.B * kbit_words(p.S) * 4 > (size_t)KBITS_MAX) return false;
  return true;
}

int fwdp(const AttnParams& p, cudaStream_t st) {
  using namespace fp;
  CUtensorMap tq, tk, tv, to;
  int rc;
  if ((rc = make_map(&tq, p.q, p.B, p.H, p.T, p.q_sb, p.q_sh, p.q_st, BM))) return rc;
  if ((rc = make_map(&to, p.out, p.B, p.H, p.T, p.o_sb, p.o_sh, p.o_st, BM))) return rc;
  if ((rc = make_map(&tk, p.k, p.B, p.Hkv, p.S, p.k_sb, p.k_sh, p.k_ss, BN))) return rc;
  if ((rc = make_map(&tv, p.v, p.B, p.Hkv, p.S, p.v_sb, p.v_sh, p.v_ss, BN))) return rc;
  const size_t smem = OFF_KBITS + (p.key_mask ? (size_t)p.B * kbit_words(p.S) * 4 : 0);
  static size_t configured = 0;
  if (smem > configured) {
    MIMIC_CUDA_OK(cudaFuncSetAttribute(attn_fwd_tcp_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_MAX));
    MIMIC_CUDA_OK(cudaFuncSetAttribute(attn_fwd_tcp_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_MAX));
    configured = SMEM_MAX;
  }
  // section = (head, batch) streams whose K + V fit the L2 budget together (GQA groups kept whole)
  const int G = p.H / p.Hkv, HB = p.H * p.B;
  const long long kv_per_hb = 2LL * p.S * HD * 2 / G;
  long long sec = l2_budget_bytes() / (kv_per_hb > 0 ? kv_per_hb : 1);
  sec = sec / G * G;
  if (sec < G) sec = G;
  if (sec > HB) sec = HB;
  const int n_items = ((p.T + 2 * BM - 1) / (2 * BM)) * HB;
  const int grid = n_items < num_sms() ? n_items : num_sms();
  // ticket counters: a ring of self-resetting (next item, CTAs done) pairs in device memory, one pair per launch in
  // flight — launches on different streams never share a pair; a captured launch keeps replaying its own
  static std::atomic<unsigned> seq{0};
  unsigned int* tickets = nullptr;
  MIMIC_CUDA_OK(cudaGetSymbolAddress(reinterpret_cast<void**>(&tickets), g_fwdp_tickets));
  unsigned int* ticket = tickets + 2 * (seq.fetch_add(1) % N_TICKETS);
  if (debug_trace_capacity() >= 4 * 16 * 16) {
    AttnParams pp = p;
    pp.trace = debug_trace_buffer();
    attn_fwd_tcp_kernel<true><<<grid, THREADS, smem, st>>>(tq, tk, tv, to, pp, (int)sec, ticket);
  } else {
    attn_fwd_tcp_kernel<false><<<grid, THREADS, smem, st>>>(tq, tk, tv, to, p, (int)sec, ticket);
  }
  MIMIC_LAUNCH_OK("attn_fwd_tcp_kernel");
  return MIMIC_OK;
}

}  // namespace tc
}  // namespace mimic
