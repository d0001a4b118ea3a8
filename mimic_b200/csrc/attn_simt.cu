// Shifted self-attention, SIMT fp32-math kernels (forward + backward).
//
// Role: (1) the fp32-I/O path whose tolerance against the oracle is 1e-5, (2) the path for
// head dims / dtypes the tcgen05 kernels do not take (d != 128), (3) the on-device cross-check
// for the tcgen05 kernels.  Math: SURVEY.md Appendix A; reference call sites
// src/shift_encoder.py:352-359 (SDPA) and 682-732 (do_shift).
//
// One statistic pair per DISJOINT key set: (m_c,l_c) over keys visible to the row (the SDPA
// softmax) and (m_x,l_x) over the excluded keys (future / masked).  The reference's unmasked
// log_Z2 (lines 697-703) is logaddexp(lse_c, lse_x): every score is exponentiated exactly once.
#include <algorithm>

#include "attn_common.cuh"

namespace mimic {
namespace simt {

constexpr int BM = 64;   // query rows per CTA
constexpr int BN = 64;   // keys per tile
constexpr int NT = 256;  // threads: 16 x 16, thread (ty,tx) owns rows ty+16i, cols tx+16j
constexpr int PLD = BN + 1;

template <typename E, int D>
__device__ __forceinline__ void load_tile(float* dst, const E* src, int64_t row_stride, int row0,
                                          int row_limit, int tid) {
  constexpr int LD = D + 4;
  constexpr int CH = D / 8;
  for (int idx = tid; idx < 64 * CH; idx += NT) {
    int r = idx / CH, c8 = idx % CH;
    float f[8];
    if (row0 + r < row_limit) {
      load8(src + (int64_t)(row0 + r) * row_stride + c8 * 8, f);
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) f[i] = 0.f;
    }
    float* d = dst + r * LD + c8 * 8;
    *reinterpret_cast<float4*>(d) = make_float4(f[0], f[1], f[2], f[3]);
    *reinterpret_cast<float4*>(d + 4) = make_float4(f[4], f[5], f[6], f[7]);
  }
}

// acc[i][j] += sum_k A[(ty+16i)][k] * Bm[(tx+16j)][k]
template <int D>
__device__ __forceinline__ void tile_abt(const float* A, const float* Bm, int ty, int tx, float (&acc)[4][4]) {
  constexpr int LD = D + 4;
#pragma unroll 4
  for (int kk = 0; kk < D; kk += 4) {
    float4 a[4], b[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) a[i] = *reinterpret_cast<const float4*>(A + (ty + 16 * i) * LD + kk);
#pragma unroll
    for (int j = 0; j < 4; ++j) b[j] = *reinterpret_cast<const float4*>(Bm + (tx + 16 * j) * LD + kk);
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j)
        acc[i][j] += a[i].x * b[j].x + a[i].y * b[j].y + a[i].z * b[j].z + a[i].w * b[j].w;
  }
}

__device__ __forceinline__ float group16_max(float v) {
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float group16_sum(float v) {
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <typename E, int D>
__global__ void __launch_bounds__(NT) attn_fwd_kernel(AttnParams p) {
  extern __shared__ __align__(16) float smem[];
  constexpr int LD = D + 4;
  constexpr int DC = D / 16;
  float* Qs = smem;
  float* Ks = Qs + BM * LD;
  float* Vs = Ks + BN * LD;
  float* Ps = Vs + BN * LD;

  const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int kvh = h / (p.H / p.Hkv);
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int T = p.T, S = p.S, off = p.S - p.T;
  const int t0 = qt * BM;
  const E* qb = reinterpret_cast<const E*>(p.q) + b * p.q_sb + h * p.q_sh;
  const E* kb = reinterpret_cast<const E*>(p.k) + b * p.k_sb + kvh * p.k_sh;
  const E* vb = reinterpret_cast<const E*>(p.v) + b * p.v_sb + kvh * p.v_sh;
  const uint8_t* km = p.key_mask ? p.key_mask + (int64_t)b * S : nullptr;

  load_tile<E, D>(Qs, qb, p.q_st, t0, T, tid);

  float m_c[4], l_c[4], m_x[4], l_x[4], o[4][DC];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    m_c[i] = m_x[i] = -INFINITY;
    l_c[i] = l_x[i] = 0.f;
#pragma unroll
    for (int c = 0; c < DC; ++c) o[i][c] = 0.f;
  }

  const int n_tiles = (S + BN - 1) / BN;
  int last_vis_tile = n_tiles - 1;
  if (p.causal) {
    int smax = min(t0 + BM - 1, T - 1) + off;  // largest visible key index of this q tile
    last_vis_tile = smax < 0 ? -1 : min(n_tiles - 1, smax / BN);
  }
  const int end_tile = p.active ? n_tiles : last_vis_tile + 1;

  for (int j = 0; j < end_tile; ++j) {
    const bool need_v = j <= last_vis_tile;
    __syncthreads();
    load_tile<E, D>(Ks, kb, p.k_ss, j * BN, S, tid);
    if (need_v) load_tile<E, D>(Vs, vb, p.v_ss, j * BN, S, tid);
    __syncthreads();

    float s[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) s[i][jj] = 0.f;
    tile_abt<D>(Qs, Ks, ty, tx, s);

    bool kin[4], kok[4];
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
      int sidx = j * BN + tx + 16 * jj;
      kin[jj] = sidx < S;
      kok[jj] = kin[jj] && (km == nullptr || km[sidx] != 0);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int t = t0 + ty + 16 * i;
      bool vis[4], exc[4];
      float mv = -INFINITY, mx = -INFINITY;
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        int sidx = j * BN + tx + 16 * jj;
        s[i][jj] *= p.scale;
        vis[jj] = kok[jj] && t < T && (!p.causal || sidx <= t + off);
        exc[jj] = p.active && kin[jj] && t < T && !vis[jj];
        if (vis[jj]) mv = fmaxf(mv, s[i][jj]);
        if (exc[jj]) mx = fmaxf(mx, s[i][jj]);
      }
      mv = group16_max(mv);
      mx = group16_max(mx);
      const float mc_new = fmaxf(m_c[i], mv), mx_new = fmaxf(m_x[i], mx);
      const float a_c = (mc_new == -INFINITY) ? 1.f : expf(m_c[i] - mc_new);
      const float a_x = (mx_new == -INFINITY) ? 1.f : expf(m_x[i] - mx_new);
      float ps = 0.f, es = 0.f;
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        float pv = vis[jj] ? expf(s[i][jj] - mc_new) : 0.f;
        float ev = exc[jj] ? expf(s[i][jj] - mx_new) : 0.f;
        ps += pv;
        es += ev;
        if (need_v) Ps[(ty + 16 * i) * PLD + tx + 16 * jj] = pv;
      }
      ps = group16_sum(ps);
      es = group16_sum(es);
      l_c[i] = l_c[i] * a_c + ps;
      l_x[i] = l_x[i] * a_x + es;
      m_c[i] = mc_new;
      m_x[i] = mx_new;
#pragma unroll
      for (int c = 0; c < DC; ++c) o[i][c] *= a_c;
    }
    if (need_v) {
      __syncthreads();
#pragma unroll 4
      for (int ss = 0; ss < BN; ++ss) {
        float pv[4], vv[DC];
#pragma unroll
        for (int i = 0; i < 4; ++i) pv[i] = Ps[(ty + 16 * i) * PLD + ss];
#pragma unroll
        for (int c = 0; c < DC; ++c) vv[c] = Vs[ss * LD + tx + 16 * c];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int c = 0; c < DC; ++c) o[i][c] += pv[i] * vv[c];
      }
    }
  }

  E* ob = reinterpret_cast<E*>(p.out) + b * p.o_sb + h * p.o_sh;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int t = t0 + ty + 16 * i;
    const float inv = l_c[i] > 0.f ? 1.f / l_c[i] : 0.f;
    const float lse_c = l_c[i] > 0.f ? m_c[i] + logf(l_c[i]) : -INFINITY;
    float mu = 0.f, lse_full = 0.f;
    if (p.active) {
      const float lse_x = l_x[i] > 0.f ? m_x[i] + logf(l_x[i]) : -INFINITY;
      const float hi = fmaxf(lse_c, lse_x), lo = fminf(lse_c, lse_x);
      lse_full = (lo == -INFINITY) ? hi : hi + log1pf(expf(lo - hi));
      float z = 0.f;
#pragma unroll
      for (int c = 0; c < DC; ++c) z += Qs[(ty + 16 * i) * LD + tx + 16 * c] * p.w[h * D + tx + 16 * c];
      z = group16_sum(z) + p.bias[h];
      mu = 1.f / (1.f + expf(lse_full - z));
    }
    if (t < T) {
#pragma unroll
      for (int c = 0; c < DC; ++c) {
        float val = o[i][c] * inv;
        if (p.active) val += mu * p.shift[h * D + tx + 16 * c];
        ob[(int64_t)t * p.o_st + tx + 16 * c] = from_f32<E>(val);
      }
      if (tx == 0) {
        const int64_t r = ((int64_t)b * p.H + h) * T + t;
        p.lse_c[r] = lse_c;
        if (p.active) {
          p.lse_full[r] = lse_full;
          p.mu[r] = mu;
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Backward, step 1: per-row scalars, dq seed, parameter gradients.
//   dmu = g.shift[h]   Delta = g.SA = g.out - mu*dmu   dz = mu(1-mu)dmu
//   Delta' = Delta + dz*exp(lse_c - lse_full)   (folds the visible part of -dz*P^f into P^c(dP-Delta'))
//   dq_acc = dz * w[h]      dshift[h] += sum_t mu_t g_t    dw[h] += sum_t dz_t q_t    db[h] += sum_t dz_t
// ONE pass over g, out and q: a warp owns a query row at a time (a lane its EPL consecutive columns, vector loads, four
// rows in flight), reduces the two row dots by shuffles, and — because every lane ends up with the row's mu and dz — adds
// its columns of mu*g and dz*q to per-lane column accumulators on the spot.  (The first version made a second pass over g
// and q, column-wise per thread, for these sums: 0.27 ms at the T = 4096 sweep shape, a tenth of the whole backward, against
// ~0.1 ms for its bytes.)  Per CTA: the four warps' column sums meet in shared memory, ONE atomicAdd per (CTA, h, c).
// (Tried in round 2 and measured slower, 0.85 vs 0.80 ms at T = 512 and 3.01 vs 2.82 ms at T = 4096 for the whole backward:
// prefetching the next group of four rows into a second set of registers — it needs ~120 registers, i.e. 4 instead of 6
// CTAs per SM, and still spills.)
// ROWS query rows per CTA: 64 for long sequences (few atomics per parameter), 16 when the whole problem is a few thousand
// rows (a MimIC query pass: T = 128) and the launch is pure latency.
template <typename E, int EPL>
__device__ __forceinline__ void load_epl(const E* p, float (&f)[EPL]) {
  if constexpr (sizeof(E) == 2 && EPL == 4) {
    const uint2 u = *reinterpret_cast<const uint2*>(p);
    const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.x));
    const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.y));
    f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y;
  } else if constexpr (sizeof(E) == 2 && EPL == 2) {
    const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(p));
    f[0] = a.x; f[1] = a.y;
  } else if constexpr (sizeof(E) == 4 && EPL == 4) {
    const float4 a = *reinterpret_cast<const float4*>(p);
    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w;
  } else if constexpr (sizeof(E) == 4 && EPL == 2) {
    const float2 a = *reinterpret_cast<const float2*>(p);
    f[0] = a.x; f[1] = a.y;
  } else {
#pragma unroll
    for (int e = 0; e < EPL; ++e) f[e] = to_f32(p[e]);
  }
}

// EPL consecutive elements of a row kept in their storage format until use (bf16: half the registers of fp32 — the four
// rows x three arrays in flight per lane are what decides this kernel's occupancy)
template <typename E, int EPL>
struct RawEpl {
  E v[EPL];
  __device__ __forceinline__ void load(const E* p) {
    if constexpr (sizeof(E) * EPL == 8) *reinterpret_cast<uint2*>(v) = *reinterpret_cast<const uint2*>(p);
    else if constexpr (sizeof(E) * EPL == 16) *reinterpret_cast<uint4*>(v) = *reinterpret_cast<const uint4*>(p);
    else if constexpr (sizeof(E) * EPL == 4) *reinterpret_cast<uint32_t*>(v) = *reinterpret_cast<const uint32_t*>(p);
    else {
#pragma unroll
      for (int e = 0; e < EPL; ++e) v[e] = p[e];
    }
  }
  __device__ __forceinline__ float get(int e) const { return to_f32(v[e]); }
};

template <typename E, int D, int ROWS>
__global__ void __launch_bounds__(128, 6) attn_bwd_prep_kernel(AttnParams p) {
  constexpr int EPL = D / 32;  // elements per lane
  __shared__ float col_s[4][D], col_w[4][D];
  __shared__ float bias_s[4];
  const int t0 = blockIdx.x * ROWS, h = blockIdx.y, b = blockIdx.z;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int T = p.T;
  const int c0 = lane * EPL;
  const E* gb = reinterpret_cast<const E*>(p.d_out) + b * p.go_sb + h * p.go_sh + c0;
  const E* ob = reinterpret_cast<const E*>(p.out) + b * p.o_sb + h * p.o_sh + c0;
  const E* qb = reinterpret_cast<const E*>(p.q) + b * p.q_sb + h * p.q_sh + c0;
  const int64_t row0 = ((int64_t)b * p.H + h) * T;
  float shv[EPL], wv[EPL], acc_s[EPL], acc_w[EPL];
#pragma unroll
  for (int e = 0; e < EPL; ++e) {
    shv[e] = p.active ? p.shift[h * D + c0 + e] : 0.f;
    wv[e] = p.active ? p.w[h * D + c0 + e] * p.dq_seed_mul : 0.f;
    acc_s[e] = 0.f;
    acc_w[e] = 0.f;
  }
  float acc_b = 0.f;
  // rows r = warp + 4 i of the CTA's block, four at a time: all their loads are issued before the first reduction
  for (int r4 = 0; r4 < ROWS / 4; r4 += 4) {
    RawEpl<E, EPL> gv[4], ov[4], qv[4];
    float mu4[4], rho4[4], dl4[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int t = t0 + warp + 4 * (r4 + i);
      const bool in = t < T;
      const int tc = in ? t : T - 1;                       // clamped: no branch around the loads, results masked below
      gv[i].load(gb + (int64_t)tc * p.go_st);
      ov[i].load(ob + (int64_t)tc * p.o_st);
      if (p.active) {
        qv[i].load(qb + (int64_t)tc * p.q_st);
        mu4[i] = p.mu[row0 + tc];                          // (the same address in every lane: one broadcast load)
        rho4[i] = p.lse_c[row0 + tc] - p.lse_full[row0 + tc];
        dl4[i] = p.d_lse_full ? p.d_lse_full[row0 + tc] : 0.f;
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int t = t0 + warp + 4 * (r4 + i);
      const bool in = t < T;
      float dmu = 0.f, gout = 0.f;
#pragma unroll
      for (int e = 0; e < EPL; ++e) {
        gout += gv[i].get(e) * ov[i].get(e);
        dmu += gv[i].get(e) * shv[e];
      }
      dmu = warp_sum(dmu);                                 // (xor shuffles: every lane holds the row's sums)
      gout = warp_sum(gout);
      float mu = 0.f, dz = 0.f, dl = gout, dz_all = 0.f;
      if (p.active) {
        mu = mu4[i];
        dz = mu * (1.f - mu) * dmu;
        // a gradient g arriving at lse_full itself adds g * P^f to dS, i.e. acts like dz = -g on the score path only
        // (it has no z1 = q.W + b behind it, so the parameter / dq-seed sums below keep the op's own dz)
        dz_all = dz - dl4[i];
        const float rho = (rho4[i] != rho4[i] || rho4[i] == -INFINITY) ? 0.f : expf(rho4[i]);   // lse_c = -inf: no visible key
        dl = gout - mu * dmu + dz_all * rho;
      }
      if (in) {
        if (lane == 0) {
          p.dz_ws[row0 + t] = dz_all;
          p.delta_ws[row0 + t] = dl;
        }
        if (!p.dq_direct) {                                  // (one key tile: the main kernel adds dz * w itself and writes bf16 dq)
          float* dqa = p.dq_acc + (row0 + t) * D + c0;
#pragma unroll
          for (int e = 0; e < EPL; ++e) dqa[e] = dz * wv[e];   // dq seed (zero for an inactive handle)
        }
        if (p.active) {
#pragma unroll
          for (int e = 0; e < EPL; ++e) {
            acc_s[e] = fmaf(mu, gv[i].get(e), acc_s[e]);
            acc_w[e] = fmaf(dz, qv[i].get(e), acc_w[e]);
          }
          acc_b += dz;
        }
      }
    }
  }
  if (p.active) {                                          // (kernel-uniform)
#pragma unroll
    for (int e = 0; e < EPL; ++e) {
      col_s[warp][c0 + e] = acc_s[e];
      col_w[warp][c0 + e] = acc_w[e];
    }
    if (lane == 0) bias_s[warp] = acc_b;
    __syncthreads();
    if (tid < D) {
      atomicAdd(p.dshift + h * D + tid, (col_s[0][tid] + col_s[1][tid]) + (col_s[2][tid] + col_s[3][tid]));
      atomicAdd(p.dw + h * D + tid, (col_w[0][tid] + col_w[1][tid]) + (col_w[2][tid] + col_w[3][tid]));
    }
    if (tid == 127) atomicAdd(p.dbias + h, (bias_s[0] + bias_s[1]) + (bias_s[2] + bias_s[3]));
  }
}

// Backward, step 2: one CTA per (key tile, kv head, batch); loops over the query heads of the
// GQA group and over query tiles; dK/dV live in registers, dQ goes out by fp32 atomics.
template <typename E, int D>
__global__ void __launch_bounds__(NT) attn_bwd_kernel(AttnParams p) {
  extern __shared__ __align__(16) float smem[];
  constexpr int LD = D + 4;
  constexpr int DC = D / 16;
  float* Ks = smem;
  float* Vs = Ks + BN * LD;
  float* Qs = Vs + BN * LD;
  float* Gs = Qs + BM * LD;
  float* Ps = Gs + BM * LD;
  float* dSs = Ps + BM * PLD;
  float* lse_c_s = dSs + BM * PLD;
  float* lse_f_s = lse_c_s + BM;
  float* delta_s = lse_f_s + BM;
  float* dz_s = delta_s + BM;

  const int kt = blockIdx.x, kvh = blockIdx.y, b = blockIdx.z;
  const int G = p.H / p.Hkv;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int T = p.T, S = p.S, off = p.S - p.T;
  const int s0 = kt * BN;
  const E* kb = reinterpret_cast<const E*>(p.k) + b * p.k_sb + kvh * p.k_sh;
  const E* vb = reinterpret_cast<const E*>(p.v) + b * p.v_sb + kvh * p.v_sh;
  const uint8_t* km = p.key_mask ? p.key_mask + (int64_t)b * S : nullptr;

  load_tile<E, D>(Ks, kb, p.k_ss, s0, S, tid);
  load_tile<E, D>(Vs, vb, p.v_ss, s0, S, tid);

  float dk[4][DC], dv[4][DC];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int c = 0; c < DC; ++c) dk[i][c] = dv[i][c] = 0.f;

  bool kin[4], kok[4];
#pragma unroll
  for (int jj = 0; jj < 4; ++jj) {
    int sidx = s0 + tx + 16 * jj;
    kin[jj] = sidx < S;
    kok[jj] = kin[jj] && (km == nullptr || km[sidx] != 0);
  }

  const int n_qt = (T + BM - 1) / BM;
  int first_qt = 0;
  if (!p.active && p.causal) {
    int tmin = s0 - off;  // first query row that can see key s0
    first_qt = tmin <= 0 ? 0 : tmin / BM;
  }

  for (int g = 0; g < G; ++g) {
    const int h = kvh * G + g;
    const E* qb = reinterpret_cast<const E*>(p.q) + b * p.q_sb + h * p.q_sh;
    const E* gb = reinterpret_cast<const E*>(p.d_out) + b * p.go_sb + h * p.go_sh;
    const int64_t row0 = ((int64_t)b * p.H + h) * T;
    for (int qt = first_qt; qt < n_qt; ++qt) {
      const int t0 = qt * BM;
      const bool need_c = !p.causal || (min(t0 + BM - 1, T - 1) + off >= s0);
      __syncthreads();
      load_tile<E, D>(Qs, qb, p.q_st, t0, T, tid);
      if (need_c) load_tile<E, D>(Gs, gb, p.go_st, t0, T, tid);
      if (tid < BM) {
        const int t = t0 + tid;
        const bool ok = t < T;
        lse_c_s[tid] = ok ? p.lse_c[row0 + t] : 0.f;
        lse_f_s[tid] = (ok && p.active) ? p.lse_full[row0 + t] : 0.f;
        delta_s[tid] = ok ? p.delta_ws[row0 + t] : 0.f;
        dz_s[tid] = ok ? p.dz_ws[row0 + t] : 0.f;
      }
      __syncthreads();

      float s[4][4], dp[4][4];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) s[i][jj] = dp[i][jj] = 0.f;
      tile_abt<D>(Qs, Ks, ty, tx, s);
      if (need_c) tile_abt<D>(Gs, Vs, ty, tx, dp);

#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int r = ty + 16 * i, t = t0 + r;
        const float lc = lse_c_s[r], lf = lse_f_s[r], dl = delta_s[r], dz = dz_s[r];
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          const int sidx = s0 + tx + 16 * jj;
          const float val = s[i][jj] * p.scale;
          const bool vis = kok[jj] && t < T && (!p.causal || sidx <= t + off);
          const bool exc = p.active && kin[jj] && t < T && !vis;
          float pv = 0.f, ds = 0.f;
          if (vis) {
            pv = expf(val - lc);
            ds = pv * (dp[i][jj] - dl);
          } else if (exc) {
            ds = -dz * expf(val - lf);
          }
          Ps[r * PLD + tx + 16 * jj] = pv;
          dSs[r * PLD + tx + 16 * jj] = ds * p.scale;
        }
      }
      __syncthreads();

      // dV[s][c] += sum_t P[t][s] g[t][c];  dK[s][c] += sum_t dS[t][s] q[t][c]   (thread: keys ty+16i, cols tx+16c)
#pragma unroll 2
      for (int t = 0; t < BM; ++t) {
        float pv[4], dsv[4], gv[DC], qv[DC];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          pv[i] = Ps[t * PLD + ty + 16 * i];
          dsv[i] = dSs[t * PLD + ty + 16 * i];
        }
#pragma unroll
        for (int c = 0; c < DC; ++c) {
          gv[c] = need_c ? Gs[t * LD + tx + 16 * c] : 0.f;
          qv[c] = Qs[t * LD + tx + 16 * c];
        }
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int c = 0; c < DC; ++c) {
            dv[i][c] += pv[i] * gv[c];
            dk[i][c] += dsv[i] * qv[c];
          }
      }
      // dQ[t][c] += sum_s dS[t][s] K[s][c]   (thread: rows ty+16i, cols tx+16c)
      float dq[4][DC];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int c = 0; c < DC; ++c) dq[i][c] = 0.f;
#pragma unroll 2
      for (int ss = 0; ss < BN; ++ss) {
        float dsv[4], kv[DC];
#pragma unroll
        for (int i = 0; i < 4; ++i) dsv[i] = dSs[(ty + 16 * i) * PLD + ss];
#pragma unroll
        for (int c = 0; c < DC; ++c) kv[c] = Ks[ss * LD + tx + 16 * c];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int c = 0; c < DC; ++c) dq[i][c] += dsv[i] * kv[c];
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int t = t0 + ty + 16 * i;
        if (t < T) {
#pragma unroll
          for (int c = 0; c < DC; ++c) atomicAdd(p.dq_acc + (row0 + t) * D + tx + 16 * c, dq[i][c]);
        }
      }
    }
  }

  E* dkb = reinterpret_cast<E*>(p.dk) + (((int64_t)b * p.Hkv + kvh) * S) * D;
  E* dvb = reinterpret_cast<E*>(p.dv) + (((int64_t)b * p.Hkv + kvh) * S) * D;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int sidx = s0 + ty + 16 * i;
    if (sidx < S) {
#pragma unroll
      for (int c = 0; c < DC; ++c) {
        dkb[(int64_t)sidx * D + tx + 16 * c] = from_f32<E>(dk[i][c]);
        dvb[(int64_t)sidx * D + tx + 16 * c] = from_f32<E>(dv[i][c]);
      }
    }
  }
}

template <typename E>
__global__ void convert_dq_kernel(const float* __restrict__ src, E* __restrict__ dst, int64_t n8, float mul) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n8; i += (int64_t)gridDim.x * blockDim.x) {
    float f[8];
    load8(src + i * 8, f);
#pragma unroll
    for (int e = 0; e < 8; ++e) f[e] *= mul;
    store8(dst + i * 8, f);
  }
}

template <typename E, int D>
int launch_fwd(const AttnParams& p, cudaStream_t st) {
  constexpr size_t smem = (size_t)(BM + 2 * BN) * (D + 4) * 4 + (size_t)BM * PLD * 4;
  MIMIC_CUDA_OK(cudaFuncSetAttribute(attn_fwd_kernel<E, D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((p.T + BM - 1) / BM, p.H, p.B);
  attn_fwd_kernel<E, D><<<grid, NT, smem, st>>>(p);
  MIMIC_LAUNCH_OK("attn_fwd_kernel(simt)");
  return MIMIC_OK;
}

template <typename E, int D>
int launch_bwd(const AttnParams& p, cudaStream_t st) {
  {
    if ((int64_t)p.B * p.H * p.T <= 65536) {
      dim3 grid((p.T + 15) / 16, p.H, p.B);
      attn_bwd_prep_kernel<E, D, 16><<<grid, 128, 0, st>>>(p);
    } else {
      dim3 grid((p.T + 63) / 64, p.H, p.B);
      attn_bwd_prep_kernel<E, D, 64><<<grid, 128, 0, st>>>(p);
    }
    MIMIC_LAUNCH_OK("attn_bwd_prep_kernel(simt)");
  }
  constexpr size_t smem = (size_t)(2 * BM + 2 * BN) * (D + 4) * 4 + (size_t)2 * BM * PLD * 4 + 4 * BM * 4;
  MIMIC_CUDA_OK(cudaFuncSetAttribute(attn_bwd_kernel<E, D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((p.S + BN - 1) / BN, p.Hkv, p.B);
  attn_bwd_kernel<E, D><<<grid, NT, smem, st>>>(p);
  MIMIC_LAUNCH_OK("attn_bwd_kernel(simt)");
  const int64_t n8 = (int64_t)p.B * p.H * p.T * D / 8;
  int blocks = (int)std::min<int64_t>((n8 + 255) / 256, (int64_t)num_sms() * 8);
  convert_dq_kernel<E><<<blocks, 256, 0, st>>>(p.dq_acc, reinterpret_cast<E*>(p.dq), n8, 1.f);
  MIMIC_LAUNCH_OK("convert_dq_kernel");
  return MIMIC_OK;
}

#define MIMIC_DISPATCH_SIMT(FN)                                                              \
  if (p.dtype == MIMIC_BF16) {                                                               \
    if (p.d == 32) return FN<__nv_bfloat16, 32>(p, st);                                      \
    if (p.d == 64) return FN<__nv_bfloat16, 64>(p, st);                                      \
    if (p.d == 128) return FN<__nv_bfloat16, 128>(p, st);                                    \
  } else {                                                                                   \
    if (p.d == 32) return FN<float, 32>(p, st);                                              \
    if (p.d == 64) return FN<float, 64>(p, st);                                              \
    if (p.d == 128) return FN<float, 128>(p, st);                                            \
  }                                                                                          \
  set_error("shifted attention (simt): head_dim %d unsupported (32, 64, 128)", p.d);         \
  return MIMIC_ERR_UNSUPPORTED;

// pieces of the backward that the tcgen05 path shares (bf16, d = 128)
int bwd_prep_bf16_d128(const AttnParams& p, cudaStream_t st) {
  if ((int64_t)p.B * p.H * p.T <= 65536) {
    dim3 grid((p.T + 15) / 16, p.H, p.B);
    attn_bwd_prep_kernel<__nv_bfloat16, 128, 16><<<grid, 128, 0, st>>>(p);
  } else {
    dim3 grid((p.T + 63) / 64, p.H, p.B);
    attn_bwd_prep_kernel<__nv_bfloat16, 128, 64><<<grid, 128, 0, st>>>(p);
  }
  MIMIC_LAUNCH_OK("attn_bwd_prep_kernel");
  return MIMIC_OK;
}
int convert_dq_bf16(const AttnParams& p, cudaStream_t st) {
  const int64_t n8 = (int64_t)p.B * p.H * p.T * p.d / 8;
  int blocks = (int)std::min<int64_t>((n8 + 255) / 256, (int64_t)num_sms() * 8);
  convert_dq_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>(p.dq_acc, reinterpret_cast<__nv_bfloat16*>(p.dq), n8, p.dq_out_mul);
  MIMIC_LAUNCH_OK("convert_dq_kernel");
  return MIMIC_OK;
}

int fwd(const AttnParams& p, cudaStream_t st) { MIMIC_DISPATCH_SIMT(launch_fwd) }
int bwd(const AttnParams& p, cudaStream_t st) { MIMIC_DISPATCH_SIMT(launch_bwd) }

}  // namespace simt
}  // namespace mimic
