// HBM-bound kernels of the MimIC hot path: record gather (R), fused layer-wise MSE forward+backward
// (M), fused LM-loss cross-entropy forward+backward (C), and the on-device clip + AdamW step.
// All are single launches over all layers / rows, 16-byte vectorised, streaming (L1 no-allocate)
// loads and stores, persistent grid-stride over rows with grids sized from the SM count.
#include <algorithm>
#include <cstdlib>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace mimic {

// ---------------------------------------------------------------------------------------------
// 8-element chunk I/O with streaming hints
template <typename T>
struct Chunk;
template <>
struct Chunk<__nv_bfloat16> {
  static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&f)[8]) {
    uint4 u = ldg_stream(p);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float2 t = __bfloat1622float2(h[i]);
      f[2 * i] = t.x;
      f[2 * i + 1] = t.y;
    }
  }
  static __device__ __forceinline__ void store(__nv_bfloat16* p, const float (&f)[8]) {
    uint4 u;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
    stg_stream(p, u);
  }
};
template <>
struct Chunk<float> {
  static __device__ __forceinline__ void load(const float* p, float (&f)[8]) {
    uint4 a = ldg_stream(p), b = ldg_stream(p + 4);
    f[0] = __uint_as_float(a.x); f[1] = __uint_as_float(a.y); f[2] = __uint_as_float(a.z); f[3] = __uint_as_float(a.w);
    f[4] = __uint_as_float(b.x); f[5] = __uint_as_float(b.y); f[6] = __uint_as_float(b.z); f[7] = __uint_as_float(b.w);
  }
  static __device__ __forceinline__ void store(float* p, const float (&f)[8]) {
    stg_stream(p, make_uint4(__float_as_uint(f[0]), __float_as_uint(f[1]), __float_as_uint(f[2]), __float_as_uint(f[3])));
    stg_stream(p + 4, make_uint4(__float_as_uint(f[4]), __float_as_uint(f[5]), __float_as_uint(f[6]), __float_as_uint(f[7])));
  }
};

// Raw 8-element chunks kept in their storage format until use: 4 chunks in flight cost 16 registers of bf16 instead of 32
// of fp32 (the CE kernel fits two 512-thread CTAs per SM with them).  Default-cached loads: pass 2 wants the lines back.
template <typename T>
struct Raw8;
template <>
struct Raw8<__nv_bfloat16> {
  uint4 u;
  __device__ __forceinline__ void load(const __nv_bfloat16* p) { u = *reinterpret_cast<const uint4*>(p); }
  __device__ __forceinline__ void load(const __nv_bfloat16* p, uint64_t policy) { u = ldg_hint(p, policy); }
  static __device__ __forceinline__ void store(__nv_bfloat16* p, const float (&f)[8], uint64_t policy) {
    uint4 v;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&v);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
    stg_hint(p, v, policy);
  }
  __device__ __forceinline__ void unpack(float (&f)[8]) const {
    // one ALU instruction per element (shift / mask); __bfloat1622float2 costs PRMT + shift for the high half
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      f[2 * i] = __uint_as_float(w[i] << 16);
      f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
  }
};
template <>
struct Raw8<float> {
  float4 a, b;
  __device__ __forceinline__ void load(const float* p) {
    a = *reinterpret_cast<const float4*>(p);
    b = *reinterpret_cast<const float4*>(p + 4);
  }
  __device__ __forceinline__ void load(const float* p, uint64_t policy) {
    const uint4 x = ldg_hint(p, policy), y = ldg_hint(p + 4, policy);
    a = make_float4(__uint_as_float(x.x), __uint_as_float(x.y), __uint_as_float(x.z), __uint_as_float(x.w));
    b = make_float4(__uint_as_float(y.x), __uint_as_float(y.y), __uint_as_float(y.z), __uint_as_float(y.w));
  }
  static __device__ __forceinline__ void store(float* p, const float (&f)[8], uint64_t policy) {
    stg_hint(p, make_uint4(__float_as_uint(f[0]), __float_as_uint(f[1]), __float_as_uint(f[2]), __float_as_uint(f[3])), policy);
    stg_hint(p + 4, make_uint4(__float_as_uint(f[4]), __float_as_uint(f[5]), __float_as_uint(f[6]), __float_as_uint(f[7])), policy);
  }
  __device__ __forceinline__ void unpack(float (&f)[8]) const {
    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w;
    f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
  }
};

__device__ __forceinline__ float block_sum(float v, float* sh /*[32]*/) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  __syncthreads();
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  v = (threadIdx.x < nw) ? sh[threadIdx.x] : 0.f;
  if (warp == 0) v = warp_sum(v);
  return v;  // valid in warp 0
}

// ---------------------------------------------------------------------------------------------
// R: record gather
template <typename TS, typename TD>
__global__ void __launch_bounds__(256) record_gather_kernel(const TS* __restrict__ src, int64_t src_row_stride,
                                                            const int32_t* __restrict__ src_row,
                                                            TD* __restrict__ dst, int N, int D) {
  const int chunks = D / 8;
  for (int n = blockIdx.x; n < N; n += gridDim.x) {
    const int sr = src_row[n];           // -1 = unused slot of the fixed-capacity buffer: zero it
    const TS* s = src + (int64_t)(sr < 0 ? 0 : sr) * src_row_stride;
    TD* d = dst + (int64_t)n * D;
    for (int c = threadIdx.x; c < chunks; c += blockDim.x) {
      float f[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      if (sr >= 0) Chunk<TS>::load(s + c * 8, f);
      Chunk<TD>::store(d + c * 8, f);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// M: fused layer-wise MSE forward + backward.  Work item = one (layer, row); a CTA walks items
// grid-stride; per item each thread moves 8-element chunks of H, H' and dH.
template <typename T, int UNROLL>
__global__ void __launch_bounds__(256) mse_fwd_bwd_kernel(const void* const* __restrict__ h_ptrs, int64_t h_row_stride,
                                                          const T* __restrict__ hp,
                                                          const int32_t* __restrict__ row_to_compact,
                                                          const float* __restrict__ row_w,
                                                          void* const* __restrict__ dh_ptrs, float grad_scale, int L,
                                                          int rows, int N, int D, float* __restrict__ partial,
                                                          unsigned int* __restrict__ counter, float* __restrict__ loss) {
  __shared__ float red[32];
  __shared__ bool is_last;
  const int chunks = D / 8;
  const int64_t items = (int64_t)L * rows;
  float acc = 0.f;
  for (int64_t it = blockIdx.x; it < items; it += gridDim.x) {
    const int l = (int)(it / rows), r = (int)(it % rows);
    const int c = row_to_compact[r];
    T* dh = dh_ptrs ? reinterpret_cast<T*>(dh_ptrs[l]) + (int64_t)r * D : nullptr;
    if (c < 0) {
      if (dh) {
        const float z[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        for (int ch = threadIdx.x; ch < chunks; ch += blockDim.x) Chunk<T>::store(dh + ch * 8, z);
      }
      continue;
    }
    const T* h = reinterpret_cast<const T*>(h_ptrs[l]) + (int64_t)r * h_row_stride;
    const T* t = hp + ((int64_t)l * N + c) * D;
    const float w = row_w[c];
    const float gw = 2.f * w * grad_scale;
    for (int ch0 = threadIdx.x; ch0 < chunks; ch0 += blockDim.x * UNROLL) {
      float a[UNROLL][8], b[UNROLL][8];
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        const int ch = ch0 + u * blockDim.x;
        if (ch < chunks) {
          Chunk<T>::load(h + ch * 8, a[u]);
          Chunk<T>::load(t + ch * 8, b[u]);
        }
      }
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        const int ch = ch0 + u * blockDim.x;
        if (ch < chunks) {
          float g[8];
          float s = 0.f;
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const float df = a[u][e] - b[u][e];
            s += df * df;
            g[e] = gw * df;
          }
          acc += w * s;
          if (dh) Chunk<T>::store(dh + ch * 8, g);
        }
      }
    }
  }
  acc = block_sum(acc, red);
  // the last CTA to finish sums the per-CTA partials in a fixed order (deterministic) and re-arms the counter
  if (threadIdx.x == 0) {
    partial[blockIdx.x] = acc;
    __threadfence();
    is_last = atomicAdd(counter, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (is_last) {
    __threadfence();
    float tot = 0.f;
    for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) tot += __ldcg(partial + i);
    tot = block_sum(tot, red);
    if (threadIdx.x == 0) {
      loss[0] = tot;
      *counter = 0u;
    }
  }
}

// Pipelined variant for rows of <= 512 chunks (D <= 4096: every BASELINE backbone), bf16.  The generic kernel above runs,
// per item, a chain of three dependent global loads (row -> compact index -> weight, layer -> pointer) before the first
// byte of the row is requested, and requests the next item's rows only after the current gradient is stored.  Here the
// compact index is fetched TWO items ahead, the weight and the rows ONE item ahead (raw bf16 chunks, 16 registers per
// item in flight), and the layer pointers sit in shared memory: while an item is reduced and its gradient stored the next
// item's 16 KB are already on their way.
__global__ void __launch_bounds__(256) mse_fwd_bwd_pipe_kernel(const void* const* __restrict__ h_ptrs, int64_t h_row_stride,
                                                               const __nv_bfloat16* __restrict__ hp,
                                                               const int32_t* __restrict__ row_to_compact,
                                                               const float* __restrict__ row_w,
                                                               void* const* __restrict__ dh_ptrs, float grad_scale, int L,
                                                               int rows, int N, int D, float* __restrict__ partial,
                                                               unsigned int* __restrict__ counter, float* __restrict__ loss) {
  using T = __nv_bfloat16;
  constexpr int MAXL = 128;
  __shared__ float red[32];
  __shared__ bool is_last;
  __shared__ const T* s_h[MAXL];
  __shared__ T* s_dh[MAXL];
  for (int l = threadIdx.x; l < L; l += blockDim.x) {
    s_h[l] = reinterpret_cast<const T*>(h_ptrs[l]);
    s_dh[l] = dh_ptrs ? reinterpret_cast<T*>(dh_ptrs[l]) : nullptr;
  }
  __syncthreads();
  const int chunks = D / 8;
  const int64_t items = (int64_t)L * rows, g = gridDim.x;
  const int ch_a = threadIdx.x, ch_b = threadIdx.x + 256;
  const bool has_a = ch_a < chunks, has_b = ch_b < chunks;
  auto compact_of = [&](int64_t it) { return it < items ? row_to_compact[(int)(it % rows)] : -1; };
  struct Rows {
    uint4 ha, hb, ta, tb;
  };
  auto request = [&](int64_t it, int c, Rows& x) {      // the item's rows of H and H' (nothing for an unselected row)
    if (c < 0) return;
    const int l = (int)(it / rows), r = (int)(it % rows);
    const T* h = s_h[l] + (int64_t)r * h_row_stride;
    const T* t = hp + ((int64_t)l * N + c) * D;
    if (has_a) {
      x.ha = ldg_stream(h + ch_a * 8);
      x.ta = ldg_stream(t + ch_a * 8);
    }
    if (has_b) {
      x.hb = ldg_stream(h + ch_b * 8);
      x.tb = ldg_stream(t + ch_b * 8);
    }
  };
  auto diff_chunk = [&](const uint4& hv, const uint4& tv, float gw, float& ssq, uint4& out) {
    Raw8<T> a, b;
    a.u = hv;
    b.u = tv;
    float fa[8], fb[8], gch[8];
    a.unpack(fa);
    b.unpack(fb);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float df = fa[e] - fb[e];
      ssq = fmaf(df, df, ssq);
      gch[e] = gw * df;
    }
    __nv_bfloat162* o = reinterpret_cast<__nv_bfloat162*>(&out);
#pragma unroll
    for (int e = 0; e < 4; ++e) o[e] = __floats2bfloat162_rn(gch[2 * e], gch[2 * e + 1]);
  };
  float acc = 0.f;
  int64_t it = blockIdx.x;
  int c0 = compact_of(it), c1 = compact_of(it + g);
  float w0 = c0 >= 0 ? row_w[c0] : 0.f;
  Rows x0 = {}, x1 = {};
  request(it, c0, x0);
  for (; it < items; it += g) {
    const int c2 = compact_of(it + 2 * g);               // in flight during this item
    const float w1 = c1 >= 0 ? row_w[c1] : 0.f;          // c1 arrived an item ago
    request(it + g, c1, x1);                             // the next item's rows travel while this one is worked on
    const int l = (int)(it / rows), r = (int)(it % rows);
    T* dh = s_dh[l] ? s_dh[l] + (int64_t)r * D : nullptr;
    if (c0 < 0) {
      if (dh) {
        const uint4 z = make_uint4(0u, 0u, 0u, 0u);
        if (has_a) stg_stream(dh + ch_a * 8, z);
        if (has_b) stg_stream(dh + ch_b * 8, z);
      }
    } else {
      const float gw = 2.f * w0 * grad_scale;
      float ssq = 0.f;
      uint4 ga, gb;
      if (has_a) {
        diff_chunk(x0.ha, x0.ta, gw, ssq, ga);
        if (dh) stg_stream(dh + ch_a * 8, ga);
      }
      if (has_b) {
        diff_chunk(x0.hb, x0.tb, gw, ssq, gb);
        if (dh) stg_stream(dh + ch_b * 8, gb);
      }
      acc = fmaf(w0, ssq, acc);
    }
    c0 = c1;
    w0 = w1;
    x0 = x1;
    c1 = c2;
  }
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) {
    partial[blockIdx.x] = acc;
    __threadfence();
    is_last = atomicAdd(counter, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (is_last) {
    __threadfence();
    float tot = 0.f;
    for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) tot += __ldcg(partial + i);
    tot = block_sum(tot, red);
    if (threadIdx.x == 0) {
      loss[0] = tot;
      *counter = 0u;
    }
  }
}

static int mse_grid(int L, int rows) {
  const int64_t items = (int64_t)L * rows;
  const int64_t cap = (int64_t)num_sms() * 8;
  return (int)(items < cap ? items : cap);
}

// ---------------------------------------------------------------------------------------------
// C: fused cross-entropy forward + backward, one CTA per row (grid-stride).  Pass 1 reads the row
// once from HBM keeping an online (max, sum); pass 2 re-reads it (L2-resident: the same CTA read
// it microseconds earlier) and writes the gradient.  Rows need not be 16-byte aligned (V = 32003):
// a scalar head/tail brackets the vector body.
template <typename T>
__device__ __forceinline__ int head_elems(const T* p, int V) {
  const uintptr_t a = reinterpret_cast<uintptr_t>(p);
  constexpr uintptr_t AL = (sizeof(T) == 2) ? 16 : 32;   // one 8-element chunk
  int h = (int)(((AL - (a % AL)) % AL) / sizeof(T));
  return h < V ? h : V;
}

__device__ __forceinline__ void online_add(float& m, float& s, float x) {
  if (x > m) {
    s = s * __expf(m - x) + 1.f;
    m = x;
  } else {
    s += __expf(x - m);
  }
}

template <typename T>
__global__ void __launch_bounds__(512, sizeof(T) == 2 ? 2 : 1) ce_fwd_bwd_kernel(const T* __restrict__ logits, int64_t row_stride,
                                                         const int64_t* __restrict__ labels, T* dlogits,
                                                         float grad_scale, int R, int V, int32_t* __restrict__ n_kept,
                                                         float* __restrict__ loss, float* __restrict__ row_loss,
                                                         unsigned int* __restrict__ counter) {
  __shared__ float red[32];
  __shared__ float bc[2];
  __shared__ bool is_last;
  const uint64_t keep = l2_keep_policy(), stream = l2_stream_policy();
  // every CTA counts the kept rows itself (R labels: a few KB from L2) — no separate counting launch
  float cnt = 0.f;
  for (int i = threadIdx.x; i < R; i += blockDim.x) cnt += labels[i] >= 0 ? 1.f : 0.f;
  cnt = block_sum(cnt, red);
  if (threadIdx.x == 0) bc[0] = cnt;
  __syncthreads();
  const float n_keep = bc[0];
  __syncthreads();
  const float inv_n = n_keep > 0.f ? 1.f / n_keep : 0.f;
  for (int r = blockIdx.x; r < R; r += gridDim.x) {
    const T* x = logits + (int64_t)r * row_stride;
    T* dx = dlogits ? dlogits + (int64_t)r * row_stride : nullptr;
    const int64_t label = labels[r];
    const int head = head_elems(x, V);
    const int body = (V - head) / 8;
    const int tail0 = head + body * 8;
    if (label < 0) {
      if (threadIdx.x == 0) row_loss[r] = 0.f;
      if (dx) {
        const float z[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        for (int i = threadIdx.x; i < head; i += blockDim.x) dx[i] = from_f32<T>(0.f);
        for (int c = threadIdx.x; c < body; c += blockDim.x) Chunk<T>::store(dx + head + c * 8, z);
        for (int i = tail0 + threadIdx.x; i < V; i += blockDim.x) dx[i] = from_f32<T>(0.f);
      }
      continue;
    }
    // pass 1: online max / sum
    float m = -INFINITY, s = 0.f;
    for (int i = threadIdx.x; i < head; i += blockDim.x) online_add(m, s, to_f32(x[i]));
    constexpr float L2E = 1.4426950408889634f;
    for (int c0 = threadIdx.x; c0 < body; c0 += blockDim.x * 4) {
      Raw8<T> raw[4];                  // 4 x 16-byte loads in flight per thread, unpacked one chunk at a time
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int c = c0 + u * blockDim.x;
        if (c < body) raw[u].load(x + head + c * 8, keep);      // pass 2 wants these lines back from L2
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int c = c0 + u * blockDim.x;
        if (c < body) {
          float f[8];
          raw[u].unpack(f);
          float cm = f[0];
#pragma unroll
          for (int e = 1; e < 8; ++e) cm = fmaxf(cm, f[e]);
          if (cm > m) {
            s *= __expf(m - cm);
            m = cm;
          }
          const float nm = -m * L2E;
#pragma unroll
          for (int e = 0; e < 8; ++e) s += fast_ex2(fmaf(f[e], L2E, nm));
        }
      }
    }
    for (int i = tail0 + threadIdx.x; i < V; i += blockDim.x) online_add(m, s, to_f32(x[i]));
    // block combine of (m, s)
    float gm = m;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) gm = fmaxf(gm, __shfl_xor_sync(0xffffffffu, gm, o));
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = gm;
    __syncthreads();
    gm = red[0];
    for (int wi = 1; wi < (int)(blockDim.x >> 5); ++wi) gm = fmaxf(gm, red[wi]);
    float ss = (m == -INFINITY) ? 0.f : s * __expf(m - gm);
    ss = block_sum(ss, red);
    if (threadIdx.x == 0) {
      bc[0] = gm;
      bc[1] = ss;
    }
    __syncthreads();
    gm = bc[0];
    const float lse = gm + __logf(bc[1]);
    const float inv_sum = 1.f / bc[1];
    if (threadIdx.x == 0) row_loss[r] = lse - to_f32(x[label]);
    if (dx) {
      const float gs = grad_scale * inv_n;
      // pass 2: gradient.  x may alias dx: every element is read before it is written by the
      // same thread, and no other thread touches it.
      for (int i = threadIdx.x; i < head; i += blockDim.x) {
        float pv = __expf(to_f32(x[i]) - gm) * inv_sum;
        dx[i] = from_f32<T>(gs * (pv - (i == label ? 1.f : 0.f)));
      }
      const float nlse2 = -(gm + __logf(bc[1])) * L2E;         // p = exp2(x * log2e - lse * log2e): one FFMA + one MUFU
      const int lab_c = ((int)label - head) >> 3;              // chunk that holds the label (only it pays the compare)
      for (int c0 = threadIdx.x; c0 < body; c0 += blockDim.x * 4) {
        Raw8<T> raw[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int c = c0 + u * blockDim.x;
          if (c < body) raw[u].load(x + head + c * 8, stream);  // last use of the row
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int c = c0 + u * blockDim.x;
          if (c < body) {
            const int base = head + c * 8;
            float f[8];
            raw[u].unpack(f);
#pragma unroll
            for (int e = 0; e < 8; ++e) f[e] = gs * fast_ex2(fmaf(f[e], L2E, nlse2));
            if (c == lab_c && (int)label >= head) {
#pragma unroll
              for (int e = 0; e < 8; ++e)
                if (base + e == (int)label) f[e] -= gs;
            }
            Raw8<T>::store(dx + base, f, stream);                 // the gradient streams out: it must not push rows out of L2
          }
        }
      }
      for (int i = tail0 + threadIdx.x; i < V; i += blockDim.x) {
        float pv = __expf(to_f32(x[i]) - gm) * inv_sum;
        dx[i] = from_f32<T>(gs * (pv - (i == label ? 1.f : 0.f)));
      }
    }
    __syncthreads();
  }
  // last CTA: deterministic sum of the row losses, mean over kept rows
  if (threadIdx.x == 0) {
    __threadfence();
    is_last = atomicAdd(counter, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (is_last) {
    __threadfence();
    float tot = 0.f;
    for (int i = threadIdx.x; i < R; i += blockDim.x) tot += __ldcg(row_loss + i);
    tot = block_sum(tot, red);
    if (threadIdx.x == 0) {
      loss[0] = tot * inv_n;
      n_kept[0] = (int32_t)(n_keep + 0.5f);
      *counter = 0u;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// C, register-resident variant (bf16): the logits row never leaves the chip between the statistics and the gradient.
// A thread-block CLUSTER of CL CTAs x 256 threads owns a row; CTA `rank` owns a contiguous run of its 16-byte chunks.
// While the cluster works on row r, ONE thread per CTA has the CTA's run of the cluster's NEXT row brought into a 40 KB
// shared-memory stage by a bulk async copy (cp.async.bulk -> mbarrier complete_tx): HBM streams during the exponentials
// instead of waiting for them (the CTAs of an SM start in lock-step and stay there, so co-residency alone did not overlap
// the load phase of one row with the arithmetic of another: 0.64 of the HBM peak before the stage).  A row is then
// stage -> registers (raw bf16: 4 registers per chunk, <= MAXC chunks per thread), (max, sum) reduced inside the CTA, the
// CTAs of the cluster exchange their two floats through distributed shared memory (one cluster barrier per row), and the
// gradient is formed from the same registers and streamed out: dram read == algorithmic (the two-pass kernel re-read
// 0.36 GB of the 0.47 GB LLaVA logits from HBM).  CL is the smallest of 1 / 2 / 4 / 8 whose registers hold the row
// (V <= 20 480 / 40 960 / 81 920 / 163 840); longer rows and fp32 logits take the two-pass kernel above.
template <int CL>
__device__ __forceinline__ void cluster_sync_all() {
  if constexpr (CL > 1) {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  } else {
    __syncthreads();
  }
}
__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
// global -> this CTA's shared memory, `bytes` (a multiple of 16, both addresses 16-byte aligned), completion on `bar`
__device__ __forceinline__ void bulk_load_to_stage(void* sdst, const void* gsrc, uint32_t bytes, uint64_t* bar, uint64_t policy) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
               ::"r"(ptx::smem_u32(sdst)), "l"(gsrc), "r"(bytes), "r"(ptx::smem_u32(bar)), "l"(policy)
               : "memory");
}
__device__ __forceinline__ void stage_wait(uint64_t* bar, uint32_t parity) {
  while (!ptx::mbar_try_wait(bar, parity)) {
  }
}

constexpr int CE_MAXC = 10;

// the run of 16-byte chunks CTA `rank` of a CL-cluster owns in a row of `body` chunks: [first, first + count)
template <int CL>
__device__ __forceinline__ void cta_run(int body, uint32_t rank, int& first, int& count) {
  const int per = (body + CL - 1) / CL;
  first = (int)rank * per;
  count = body - first;
  count = count < 0 ? 0 : (count > per ? per : count);
}

// (max, sum) partials travel between the CTAs of a cluster as asynchronous remote shared-memory stores that complete a
// transaction count on the RECEIVER's mbarrier (st.async ... mbarrier::complete_tx::bytes): no cluster barrier per row.
// (barrier.cluster.arrive.release made every row wait until the previous row's gradient stores had drained — `membar`
// was the second largest stall reason of the kernel — and the partials were then read with 8 remote loads.)
__device__ __forceinline__ uint32_t map_to_cta(const void* local, uint32_t cta) {
  uint32_t ra;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(ptx::smem_u32(local)), "r"(cta));
  return ra;
}
__device__ __forceinline__ void st_async_f2(uint32_t remote_dst, float a, float b, uint32_t remote_bar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.f32 [%0], {%1, %2}, [%3];"
               ::"r"(remote_dst), "f"(a), "f"(b), "r"(remote_bar)
               : "memory");
}
// combine N <= 32 (max, sum) pairs of a softmax (sum relative to its max) held in shared memory: lane i takes pair i, the
// maximum and then the rescaled sum go through warp shuffles — every warp does it redundantly, no registers for the N
// pairs, no data-dependent selects; (-inf, 0) entries drop out (2^-inf = 0)
template <int N>
__device__ __forceinline__ void ms_combine(const float2* v, int lane, float& M, float& S) {
  constexpr float L2E = 1.4426950408889634f;
  const float2 t = lane < N ? v[lane] : make_float2(-INFINITY, 0.f);
  M = t.x;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) M = fmaxf(M, __shfl_xor_sync(0xffffffffu, M, o));
  const float nM = (M == -INFINITY) ? 0.f : -M * L2E;
  S = warp_sum(t.y * fast_ex2(fmaf(t.x, L2E, nM)));
}

template <int CL, int NT>
__global__ void __launch_bounds__(NT, 65536 / (NT * 64)) ce_fwd_bwd_reg_kernel(const __nv_bfloat16* __restrict__ logits, int64_t row_stride,
                                                                const int64_t* __restrict__ labels, __nv_bfloat16* dlogits,
                                                                float grad_scale, int R, int V, int32_t* __restrict__ n_kept,
                                                                float* __restrict__ loss, float* __restrict__ row_loss,
                                                                unsigned int* __restrict__ counter) {
  using T = __nv_bfloat16;
  constexpr int NW = NT / 32;
  extern __shared__ __align__(128) unsigned char ce_stage[];   // CE_MAXC * NT * 16 bytes: this CTA's run of the next row
  __shared__ float red[32];
  __shared__ float bc[2];
  __shared__ float2 wred[2][NW];           // per-warp (max, sum) of a row, double-buffered by processed-row parity
  __shared__ __align__(8) float2 slots[2][CL];   // (max, sum) of every CTA of the cluster for a row, same double buffer
  __shared__ uint64_t full_bar;            // the stage holds the row it was armed for
  __shared__ uint64_t xbar[2];             // slots[par] are complete
  __shared__ bool is_last;
  constexpr float L2E = 1.4426950408889634f;
  const uint32_t rank = CL > 1 ? cluster_rank() : 0u;
  const int cluster_id = blockIdx.x / CL, n_clusters = gridDim.x / CL;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint64_t stream = l2_stream_policy();
  if (tid == 0) {
    ptx::mbar_init(&full_bar, 1);
    ptx::mbar_init(&xbar[0], 1);
    ptx::mbar_init(&xbar[1], 1);
    ptx::fence_barrier_init();
  }
  // thread 0: have this CTA's run of row r brought into the stage (rows without a label are never staged)
  auto stage_row = [&](int r) {
    if (r >= R || labels[r] < 0) return;
    const T* x = logits + (int64_t)r * row_stride;
    const int head = head_elems(x, V);
    int first, count;
    cta_run<CL>((V - head) / 8, rank, first, count);
    if (count > 0) {
      ptx::mbar_expect_tx(&full_bar, (uint32_t)count * 16u);
      bulk_load_to_stage(ce_stage, x + head + first * 8, (uint32_t)count * 16u, &full_bar, stream);
    }
  };
  // the <= 7 + 7 elements in front of / behind the 16-byte aligned body (one each for the first threads of CTA 0) and the
  // logit at the label (thread 64 of CTA 0) are fetched ONE ROW AHEAD, when the stage is armed: nobody waits for them
  auto edge_index = [&](int head, int tail0) {
    return rank != 0 ? -1 : (tid < head ? tid : (tid >= 32 && tid < 40 && tail0 + (tid - 32) < V ? tail0 + tid - 32 : -1));
  };
  auto edge_fetch = [&](int r) -> float {
    if (rank != 0 || r >= R) return -INFINITY;
    const int64_t lab = labels[r];
    if (lab < 0) return -INFINITY;
    const T* x = logits + (int64_t)r * row_stride;
    if (tid == 64) return to_f32(x[lab]);
    const int head = head_elems(x, V);
    const int ei = edge_index(head, head + ((V - head) / 8) * 8);
    return ei >= 0 ? to_f32(x[ei]) : -INFINITY;
  };
  float cnt = 0.f;
  for (int i = tid; i < R; i += blockDim.x) cnt += labels[i] >= 0 ? 1.f : 0.f;
  cnt = block_sum(cnt, red);                // (its barriers also publish the mbarrier init inside the CTA)
  if (tid == 0) {
    bc[0] = cnt;
    stage_row(cluster_id);
  }
  float edge = edge_fetch(cluster_id);
  cluster_sync_all<CL>();                   // every CTA's exchange barriers exist before anyone sends to them
  const float n_keep = bc[0];
  const float inv_n = n_keep > 0.f ? 1.f / n_keep : 0.f;
  const float gs = grad_scale * inv_n;
  uint32_t done = 0;                       // rows this cluster has put through the exchange (slot / barrier parity)
  uint32_t staged = 0;                     // rows this CTA has taken out of its stage (mbarrier phase parity)
  for (int r = cluster_id; r < R; r += n_clusters) {
    const T* x = logits + (int64_t)r * row_stride;
    T* dx = dlogits ? dlogits + (int64_t)r * row_stride : nullptr;
    const int64_t label = labels[r];
    const int head = head_elems(x, V);
    const int body = (V - head) / 8;
    const int tail0 = head + body * 8;
    int first, count;
    cta_run<CL>(body, rank, first, count);
    if (label < 0) {                       // (uniform over the cluster: no exchange on this path)
      if (tid == 0) stage_row(r + n_clusters);      // the stage is free: the previous row left it before its arithmetic
      edge = edge_fetch(r + n_clusters);
      if (rank == 0 && tid == 0) row_loss[r] = 0.f;
      if (dx) {
        const float z[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        const int gtid = rank * NT + tid;
        constexpr int GT = CL * NT;
        for (int i = gtid; i < head; i += GT) dx[i] = from_f32<T>(0.f);
        for (int c = gtid; c < body; c += GT) Chunk<T>::store(dx + head + c * 8, z);
        for (int i = tail0 + gtid; i < V; i += GT) dx[i] = from_f32<T>(0.f);
      }
      continue;
    }
    // ---- the row, once: stage -> registers (<= MAXC chunks per thread), then the stage goes to the next row ----
    Raw8<T> raw[CE_MAXC];
    if (count > 0) {
      stage_wait(&full_bar, staged & 1);
      ++staged;
    }
#pragma unroll
    for (int u = 0; u < CE_MAXC; ++u) {
      const int c = tid + u * NT;
      if (c < count) raw[u].u = *reinterpret_cast<const uint4*>(ce_stage + (size_t)c * 16);
    }
    const int ei = edge_index(head, tail0);
    const float ex = ei >= 0 ? edge : -INFINITY;
    const float x_label = edge;            // (thread 64 of CTA 0)
    __syncthreads();
    if (tid == 0) stage_row(r + n_clusters);
    edge = edge_fetch(r + n_clusters);
    // thread maximum on the packed pairs (VHMNMX.BF16_V2: a quarter of an instruction per element, no unpacking)
    float m = ex;
    {
      __nv_bfloat162 mm = __halves2bfloat162(__ushort_as_bfloat16(0xFF80u), __ushort_as_bfloat16(0xFF80u));   // (-inf, -inf)
#pragma unroll
      for (int u = 0; u < CE_MAXC; ++u) {
        if (tid + u * NT < count) {
          const __nv_bfloat162* hh = reinterpret_cast<const __nv_bfloat162*>(&raw[u].u);
          mm = __hmax2(__hmax2(mm, hh[0]), __hmax2(hh[1], __hmax2(hh[2], hh[3])));
        }
      }
      const float2 m2 = __bfloat1622float2(mm);
      m = fmaxf(m, fmaxf(m2.x, m2.y));
    }
    // e = exp2((x - m_thread) log2e): summed in fp32, then written back over the raw chunk as bf16 pairs — the gradient
    // pass multiplies it by a per-thread factor instead of exponentiating again (ONE MUFU per element, the kernel was
    // issue / MUFU bound at two; the extra rounding is below the bf16 rounding of the gradient itself)
    float s = 0.f, ee = 0.f;
    {
      const float nm = (m == -INFINITY) ? 0.f : -m * L2E;
      if (ei >= 0) {
        ee = fast_ex2(fmaf(ex, L2E, nm));
        s += ee;
      }
#pragma unroll
      for (int u = 0; u < CE_MAXC; ++u) {
        if (tid + u * NT < count) {
          float f[8];
          raw[u].unpack(f);
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            f[e] = fast_ex2(fmaf(f[e], L2E, nm));
            s += f[e];
          }
          __nv_bfloat162* hh = reinterpret_cast<__nv_bfloat162*>(&raw[u].u);
#pragma unroll
          for (int e = 0; e < 4; ++e) hh[e] = __floats2bfloat162_rn(f[2 * e], f[2 * e + 1]);
        }
      }
    }
    // ---- warp (max, sum) -> shared; ONE block barrier; every thread merges the 8 warps; the cluster's by st.async ----
    const int par = done & 1;
    {
      float wm = m;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) wm = fmaxf(wm, __shfl_xor_sync(0xffffffffu, wm, o));
      float ws = (m == -INFINITY) ? 0.f : s * fast_ex2((m - wm) * L2E);
      ws = warp_sum(ws);
      if (lane == 0) wred[par][warp] = make_float2(wm, ws);
    }
    __syncthreads();
    float M, Ssum;
    ms_combine<NW>(wred[par], lane, M, Ssum);
    if constexpr (CL > 1) {
      if (tid == 0) ptx::mbar_expect_tx(&xbar[par], CL * 8u);
      if (tid < CL) st_async_f2(map_to_cta(&slots[par][rank], (uint32_t)tid), M, Ssum, map_to_cta(&xbar[par], (uint32_t)tid));
      stage_wait(&xbar[par], (done >> 1) & 1);
      ms_combine<CL>(slots[par], lane, M, Ssum);
    }
    ++done;
    const float lse = M + __logf(Ssum);
    if (rank == 0 && tid == 64) row_loss[r] = lse - x_label;
    // ---- gradient from the same registers: p = e * exp2((m_thread - lse) log2e) on the packed pairs ----
    // fac = fh + fl (two bf16 terms: fac to 2^-17), g = e * fh + e * fl with HFMA2.BF16 — one rounding to bf16 as before,
    // a quarter of the instructions of unpack / FMUL / pack (the kernel is issue-bound: ~1950 instructions per thread and
    // row before this, 0.62 of the HBM peak)
    const float fac = (m == -INFINITY) ? 0.f : gs * fast_ex2((m - lse) * L2E);
    if (dx) {
      if (ei >= 0) dx[ei] = from_f32<T>(ee * fac - (ei == (int)label ? gs : 0.f));
      const __nv_bfloat16 fh = __float2bfloat16_rn(fac);
      const __nv_bfloat16 fl = __float2bfloat16_rn(fac - __bfloat162float(fh));
      const __nv_bfloat162 fh2 = __halves2bfloat162(fh, fh), fl2 = __halves2bfloat162(fl, fl);
      const int lab_e = (int)label - head - first * 8;      // position of the label inside this CTA's run (may be outside it)
      const int lab_c = lab_e >= 0 ? (lab_e >> 3) : -1;     // the chunk of the run that holds it: only that one is patched
      T* dxp = dx + head + (size_t)(first + tid) * 8;
#pragma unroll
      for (int u = 0; u < CE_MAXC; ++u) {
        const int c = tid + u * NT;
        if (c < count) {
          uint4 g;
          __nv_bfloat162* gg = reinterpret_cast<__nv_bfloat162*>(&g);
          const __nv_bfloat162* hh = reinterpret_cast<const __nv_bfloat162*>(&raw[u].u);
#pragma unroll
          for (int e = 0; e < 4; ++e) gg[e] = __hfma2(hh[e], fh2, __hmul2(hh[e], fl2));
          if (c == lab_c) {                                 // p - 1 at the label, in fp32 like every other element was before
            const int k = lab_e & 7;
            const uint32_t w = k < 2 ? raw[u].u.x : (k < 4 ? raw[u].u.y : (k < 6 ? raw[u].u.z : raw[u].u.w));
            const float e_lab = __uint_as_float((k & 1) ? (w & 0xffff0000u) : (w << 16));
            const uint32_t pv = (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(e_lab * fac - gs));
            const uint32_t keep = (k & 1) ? 0x0000ffffu : 0xffff0000u, val = (k & 1) ? (pv << 16) : pv;
            const int ws = k >> 1;                          // (selects, not a reference: g stays in registers)
            g.x = ws == 0 ? ((g.x & keep) | val) : g.x;
            g.y = ws == 1 ? ((g.y & keep) | val) : g.y;
            g.z = ws == 2 ? ((g.z & keep) | val) : g.z;
            g.w = ws == 3 ? ((g.w & keep) | val) : g.w;
          }
          stg_stream(dxp + (size_t)u * NT * 8, g);
        }
      }
    }
  }
  cluster_sync_all<CL>();                   // no CTA leaves while a neighbour may still send to its slots
  if (tid == 0) {
    __threadfence();
    is_last = atomicAdd(counter, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (is_last) {
    __threadfence();
    float tot = 0.f;
    for (int i = tid; i < R; i += blockDim.x) tot += __ldcg(row_loss + i);
    tot = block_sum(tot, red);
    if (tid == 0) {
      loss[0] = tot * inv_n;
      n_kept[0] = (int32_t)(n_keep + 0.5f);
      *counter = 0u;
    }
  }
}

template <int CL, int NT>
static int launch_ce_reg(const __nv_bfloat16* logits, int64_t row_stride, const int64_t* labels, __nv_bfloat16* dlogits,
                         float grad_scale, int R, int V, int32_t* n_kept, float* loss, float* row_loss, unsigned int* counter,
                         cudaStream_t st) {
  cudaLaunchConfig_t cfg{};
  cfg.blockDim = dim3(NT);
  cfg.dynamicSmemBytes = CE_MAXC * NT * 16;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  // persistent clusters: exactly as many as the device co-schedules (a cluster that has to wait for a second wave would
  // start its static share of the rows when the others are done — measured: 74 launched, 71 resident, SMs 33 % idle)
  static int max_clusters = 0;
  if (max_clusters == 0) {
    MIMIC_CUDA_OK(cudaFuncSetAttribute(ce_fwd_bwd_reg_kernel<CL, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, CE_MAXC * NT * 16));
    MIMIC_CUDA_OK(cudaFuncSetAttribute(ce_fwd_bwd_reg_kernel<CL, NT>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                       (int)cudaSharedmemCarveoutMaxShared));
    cfg.gridDim = dim3(CL * num_sms() * 4);
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, ce_fwd_bwd_reg_kernel<CL, NT>, &cfg) != cudaSuccess || n <= 0) n = num_sms() * (65536 / (NT * 64)) / CL - 4;
    max_clusters = n > 0 ? n : 1;
  }
  int n_clusters = max_clusters < R ? max_clusters : R;
  cfg.gridDim = dim3(n_clusters * CL);
  MIMIC_CUDA_OK(cudaLaunchKernelEx(&cfg, ce_fwd_bwd_reg_kernel<CL, NT>, logits, row_stride, labels, dlogits, grad_scale, R, V,
                                   n_kept, loss, row_loss, counter));
  return MIMIC_OK;
}

// ---------------------------------------------------------------------------------------------
// Layer-wise cosine-similarity alignment (Strategy.LAYER_WISE_COS_SIM, src/shift_model.py:152-160), forward +
// backward in one launch.  One CTA per (layer, row) item keeps the row of H and H' in registers between the
// reduction (h.h', |h|^2, |h'|^2) and the gradient, so HBM is read once:  3 * N * sizeof(T) algorithmic bytes, as M.
//   loss = sum_{l,n} row_w[n] * (1 - cos_ln),  cos = h.h' / (max(|h|, eps) * max(|h'|, eps))   (F.cosine_similarity, eps 1e-8)
//   dH   = -grad_scale * row_w * ( h' / (|h||h'|) - cos * h / |h|^2 )
template <typename T, int MAXU>
__global__ void __launch_bounds__(256) cos_fwd_bwd_kernel(const void* const* __restrict__ h_ptrs, int64_t h_row_stride,
                                                          const T* __restrict__ hp,
                                                          const int32_t* __restrict__ row_to_compact,
                                                          const float* __restrict__ row_w,
                                                          void* const* __restrict__ dh_ptrs, float grad_scale, int L,
                                                          int rows, int N, int D, float* __restrict__ partial,
                                                          unsigned int* __restrict__ counter, float* __restrict__ loss) {
  __shared__ float red[3][32];
  __shared__ float bc[3];
  __shared__ bool is_last;
  const int chunks = D / 8;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int64_t items = (int64_t)L * rows;
  float acc = 0.f;
  for (int64_t it = blockIdx.x; it < items; it += gridDim.x) {
    const int l = (int)(it / rows), r = (int)(it % rows);
    const int c = row_to_compact[r];
    T* dh = dh_ptrs ? reinterpret_cast<T*>(dh_ptrs[l]) + (int64_t)r * D : nullptr;
    if (c < 0) {
      if (dh) {
        const float z[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        for (int ch = threadIdx.x; ch < chunks; ch += blockDim.x) Chunk<T>::store(dh + ch * 8, z);
      }
      continue;
    }
    const T* h = reinterpret_cast<const T*>(h_ptrs[l]) + (int64_t)r * h_row_stride;
    const T* t = hp + ((int64_t)l * N + c) * D;
    // (a register-prefetch pipeline of the next item was measured: 158 registers, occupancy loss, 2.4 TB/s vs 4.3 TB/s)
    float a[MAXU][8], b[MAXU][8];
    float dot = 0.f, na = 0.f, nb = 0.f;
#pragma unroll
    for (int u = 0; u < MAXU; ++u) {
      const int ch = threadIdx.x + u * blockDim.x;
      if (ch < chunks) {
        Chunk<T>::load(h + ch * 8, a[u]);
        Chunk<T>::load(t + ch * 8, b[u]);
      }
    }
#pragma unroll
    for (int u = 0; u < MAXU; ++u) {
      const int ch = threadIdx.x + u * blockDim.x;
      if (ch < chunks) {
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          dot = fmaf(a[u][e], b[u][e], dot);
          na = fmaf(a[u][e], a[u][e], na);
          nb = fmaf(b[u][e], b[u][e], nb);
        }
      }
    }
    dot = warp_sum(dot);
    na = warp_sum(na);
    nb = warp_sum(nb);
    __syncthreads();                                   // previous item's bc[] reads are done
    if (lane == 0) {
      red[0][warp] = dot;
      red[1][warp] = na;
      red[2][warp] = nb;
    }
    __syncthreads();
    if (warp == 0) {
      float v0 = lane < nw ? red[0][lane] : 0.f, v1 = lane < nw ? red[1][lane] : 0.f, v2 = lane < nw ? red[2][lane] : 0.f;
      v0 = warp_sum(v0);
      v1 = warp_sum(v1);
      v2 = warp_sum(v2);
      if (lane == 0) {
        bc[0] = v0;
        bc[1] = v1;
        bc[2] = v2;
      }
    }
    __syncthreads();
    const float ha = fmaxf(sqrtf(bc[1]), 1e-8f), hb = fmaxf(sqrtf(bc[2]), 1e-8f);
    const float inv = 1.f / (ha * hb);
    const float cs = bc[0] * inv;
    const float w = row_w[c];
    if (threadIdx.x == 0) acc += w * (1.f - cs);
    if (dh) {
      const float gw = -grad_scale * w;
      const float k1 = gw * inv, k2 = gw * cs / (ha * ha);
#pragma unroll
      for (int u = 0; u < MAXU; ++u) {
        const int ch = threadIdx.x + u * blockDim.x;
        if (ch < chunks) {
          float g[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) g[e] = k1 * b[u][e] - k2 * a[u][e];
          Chunk<T>::store(dh + ch * 8, g);
        }
      }
    }
  }
  acc = block_sum(acc, red[0]);
  if (threadIdx.x == 0) {
    partial[blockIdx.x] = acc;
    __threadfence();
    is_last = atomicAdd(counter, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (is_last) {
    __threadfence();
    float tot = 0.f;
    for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) tot += __ldcg(partial + i);
    tot = block_sum(tot, red[0]);
    if (threadIdx.x == 0) {
      loss[0] = tot;
      *counter = 0u;
    }
  }
}

// Variant with an asynchronous two-stage pipeline through shared memory (cp.async, 16 bytes per copy): the rows of H
// and H' of item i+1 stream into stage (i+1)&1 while item i is reduced and its gradient written out of stage i&1, so
// the block-wide barriers of the reduction no longer leave the memory system idle (the register version loads, waits,
// synchronises three times, stores — 4.25 TB/s; a register prefetch of the next item costs too many registers).
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <typename T, int NS>
__global__ void __launch_bounds__(256) cos_fwd_bwd_pipe_kernel(const void* const* __restrict__ h_ptrs, int64_t h_row_stride,
                                                               const T* __restrict__ hp,
                                                               const int32_t* __restrict__ row_to_compact,
                                                               const float* __restrict__ row_w,
                                                               void* const* __restrict__ dh_ptrs, float grad_scale, int L,
                                                               int rows, int N, int D, float* __restrict__ partial,
                                                               unsigned int* __restrict__ counter, float* __restrict__ loss) {
  extern __shared__ __align__(16) uint8_t cos_smem[];    // [NS stages][H row | H' row], D * sizeof(T) bytes each
  __shared__ float red[3][32];
  __shared__ float bc[3];
  __shared__ bool is_last;
  const int chunks = D / 8;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int64_t items = (int64_t)L * rows;
  const size_t row_bytes = (size_t)D * sizeof(T);
  constexpr int VPC = sizeof(T) == 2 ? 1 : 2;            // 16-byte vectors per 8-element chunk
  auto stage_h = [&](int st) { return reinterpret_cast<T*>(cos_smem + (size_t)st * 2 * row_bytes); };
  auto stage_t = [&](int st) { return reinterpret_cast<T*>(cos_smem + (size_t)st * 2 * row_bytes + row_bytes); };
  auto issue = [&](int64_t it, int st) -> int {          // returns the compact row (or -1: nothing to load)
    int c = -1;
    if (it < items) {
      const int l = (int)(it / rows), r = (int)(it % rows);
      c = row_to_compact[r];
      if (c >= 0) {
        const T* h = reinterpret_cast<const T*>(h_ptrs[l]) + (int64_t)r * h_row_stride;
        const T* t = hp + ((int64_t)l * N + c) * D;
        T* sh = stage_h(st);
        T* stt = stage_t(st);
        for (int v = threadIdx.x; v < chunks * VPC; v += blockDim.x) {
          cp_async16(reinterpret_cast<uint4*>(sh) + v, reinterpret_cast<const uint4*>(h) + v);
          cp_async16(reinterpret_cast<uint4*>(stt) + v, reinterpret_cast<const uint4*>(t) + v);
        }
      }
    }
    cp_async_commit();                                   // always one group per item: the wait counts stay regular
    return c;
  };
  float acc = 0.f;
  int cq[NS];                                            // compact rows of the items in flight (item k lives in stage k % NS)
#pragma unroll
  for (int j = 0; j < NS - 1; ++j) cq[j] = issue(blockIdx.x + (int64_t)j * gridDim.x, j);
  int k = 0;
  for (int64_t it = blockIdx.x; it < items; it += gridDim.x, ++k) {
    const int st = k % NS;
    // item k + NS - 1 goes into the stage that the barrier ending item k - 1 released
    cq[NS - 1] = issue(it + (int64_t)(NS - 1) * gridDim.x, (k + NS - 1) % NS);
    cp_async_wait<NS - 1>();                             // all but the newest NS - 1 groups have landed: item k is in stage st
    __syncthreads();
    const int l = (int)(it / rows), r = (int)(it % rows);
    const int c = cq[0];
#pragma unroll
    for (int j = 0; j < NS - 1; ++j) cq[j] = cq[j + 1];
    T* dh = dh_ptrs ? reinterpret_cast<T*>(dh_ptrs[l]) + (int64_t)r * D : nullptr;
    if (c < 0) {
      if (dh) {
        const float z[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        for (int ch = threadIdx.x; ch < chunks; ch += blockDim.x) Chunk<T>::store(dh + ch * 8, z);
      }
      __syncthreads();
      continue;
    }
    const T* sh = stage_h(st);
    const T* stt = stage_t(st);
    float dot = 0.f, na = 0.f, nb = 0.f;
    for (int ch = threadIdx.x; ch < chunks; ch += blockDim.x) {
      float a[8], b[8];
      load8(sh + ch * 8, a);
      load8(stt + ch * 8, b);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        dot = fmaf(a[e], b[e], dot);
        na = fmaf(a[e], a[e], na);
        nb = fmaf(b[e], b[e], nb);
      }
    }
    dot = warp_sum(dot);
    na = warp_sum(na);
    nb = warp_sum(nb);
    if (lane == 0) {
      red[0][warp] = dot;
      red[1][warp] = na;
      red[2][warp] = nb;
    }
    __syncthreads();
    if (warp == 0) {
      float v0 = lane < nw ? red[0][lane] : 0.f, v1 = lane < nw ? red[1][lane] : 0.f, v2 = lane < nw ? red[2][lane] : 0.f;
      v0 = warp_sum(v0);
      v1 = warp_sum(v1);
      v2 = warp_sum(v2);
      if (lane == 0) {
        bc[0] = v0;
        bc[1] = v1;
        bc[2] = v2;
      }
    }
    __syncthreads();
    const float ha = fmaxf(sqrtf(bc[1]), 1e-8f), hb = fmaxf(sqrtf(bc[2]), 1e-8f);
    const float inv = 1.f / (ha * hb);
    const float cs = bc[0] * inv;
    const float w = row_w[c];
    if (threadIdx.x == 0) acc += w * (1.f - cs);
    if (dh) {
      const float gw = -grad_scale * w;
      const float k1 = gw * inv, k2 = gw * cs / (ha * ha);
      for (int ch = threadIdx.x; ch < chunks; ch += blockDim.x) {
        float a[8], b[8], g[8];
        load8(sh + ch * 8, a);
        load8(stt + ch * 8, b);
#pragma unroll
        for (int e = 0; e < 8; ++e) g[e] = k1 * b[e] - k2 * a[e];
        Chunk<T>::store(dh + ch * 8, g);
      }
    }
    __syncthreads();                                     // stage st and bc[] are free again
  }
  cp_async_wait<0>();
  acc = block_sum(acc, red[0]);
  if (threadIdx.x == 0) {
    partial[blockIdx.x] = acc;
    __threadfence();
    is_last = atomicAdd(counter, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (is_last) {
    __threadfence();
    float tot = 0.f;
    for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) tot += __ldcg(partial + i);
    tot = block_sum(tot, red[0]);
    if (threadIdx.x == 0) {
      loss[0] = tot;
      *counter = 0u;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Logits KL divergence (Strategy.LOGITS_KL_DIV, src/shift_model.py:181-191), forward + backward in one launch:
//   loss = (1/R) sum_r KL( softmax(y_r) || softmax(z_r) ),   dz = grad_scale / R * (softmax(z) - softmax(y))
// z = query-pass ("shift") logits row i, y = context-pass ("prefix") logits row y_row[i]; y_row[i] < 0 = row not selected
// (gradient row zeroed).  Rows of the two tensors start at unrelated alignments (V is odd for the real vocabularies), so the
// 8-element chunks are aligned for z (and dz) and y comes in through 16-byte windows (two aligned loads + funnel shift;
// measured 0.95 TB/s with element-wise loads instead).  Pass 1 keeps
// online (max, sum) for both rows plus A = sum e^{y-m}(y - z);  KL_r = A / S_y - lse_y + lse_z.  Pass 2 re-reads both rows
// (L2) and writes the gradient:  3 * R * V * sizeof(T) algorithmic bytes.
struct KlAcc {
  float mz, sz, my, sy, ay;
};
__device__ __forceinline__ void kl_add(KlAcc& k, float z, float y) {
  if (z > k.mz) {
    k.sz *= __expf(k.mz - z);
    k.mz = z;
  }
  k.sz += __expf(z - k.mz);
  if (y > k.my) {
    const float f = __expf(k.my - y);
    k.sy *= f;
    k.ay *= f;
    k.my = y;
  }
  const float e = __expf(y - k.my);
  k.sy += e;
  k.ay = fmaf(e, y - z, k.ay);
}

// 16 bytes starting at an address that is only element-aligned: two aligned 16-byte loads (neighbouring lanes' windows
// overlap, so the second one hits L1) and a byte-granular funnel shift.  The second vector is only touched when the
// window really straddles it, so no byte outside the 16-byte granules holding valid data is read.
__device__ __forceinline__ uint4 ldg_window_part(const uint4* p, uint64_t policy) {
  // default L1 caching on purpose: neighbouring lanes' windows overlap, the second aligned vector is an L1 hit
  uint4 r;
  asm volatile("ld.global.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p), "l"(policy));
  return r;
}
// bytes [o, o + 16) of the 32 bytes (lo, hi), 0 < o < 16
__device__ __forceinline__ uint4 window_combine(const uint4& lo, const uint4& hi, unsigned o) {
  const unsigned long long q0 = ((unsigned long long)lo.y << 32) | lo.x, q1 = ((unsigned long long)lo.w << 32) | lo.z;
  const unsigned long long q2 = ((unsigned long long)hi.y << 32) | hi.x, q3 = ((unsigned long long)hi.w << 32) | hi.z;
  const unsigned long long x0 = o < 8 ? q0 : q1, x1 = o < 8 ? q1 : q2, x2 = o < 8 ? q2 : q3;
  const unsigned sh = (o & 7) * 8;
  const unsigned long long r0 = sh ? (x0 >> sh) | (x1 << (64 - sh)) : x0;
  const unsigned long long r1 = sh ? (x1 >> sh) | (x2 << (64 - sh)) : x1;
  return make_uint4((unsigned)r0, (unsigned)(r0 >> 32), (unsigned)r1, (unsigned)(r1 >> 32));
}
__device__ __forceinline__ uint4 load16_window(const void* p, uint64_t policy) {
  const uintptr_t a = reinterpret_cast<uintptr_t>(p);
  const uint4* base = reinterpret_cast<const uint4*>(a & ~uintptr_t(15));
  const unsigned o = (unsigned)(a & 15);
  const uint4 lo = ldg_window_part(base, policy);
  if (o == 0) return lo;
  return window_combine(lo, ldg_window_part(base + 1, policy), o);
}
__device__ __forceinline__ void load8_window(const __nv_bfloat16* p, float (&f)[8], uint64_t policy) {
  const uint4 u = load16_window(p, policy);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = __bfloat1622float2(h[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ void load8_window(const float* p, float (&f)[8], uint64_t policy) {
  const uint4 a = load16_window(p, policy), b = load16_window(p + 4, policy);
  f[0] = __uint_as_float(a.x); f[1] = __uint_as_float(a.y); f[2] = __uint_as_float(a.z); f[3] = __uint_as_float(a.w);
  f[4] = __uint_as_float(b.x); f[5] = __uint_as_float(b.y); f[6] = __uint_as_float(b.z); f[7] = __uint_as_float(b.w);
}

template <typename T>
__global__ void __launch_bounds__(512, sizeof(T) == 2 ? 2 : 1) kl_fwd_bwd_kernel(const T* __restrict__ zl, int64_t z_stride, const T* __restrict__ yl,
                                                         int64_t y_stride, const int32_t* __restrict__ y_row, T* dz,
                                                         float grad_scale, int R, int V, int32_t* __restrict__ n_rows,
                                                         float* __restrict__ loss, float* __restrict__ row_loss,
                                                         unsigned int* __restrict__ counter) {
  __shared__ float red[32];
  __shared__ float bc[5];
  __shared__ bool is_last;
  const uint64_t keep = l2_keep_policy(), stream = l2_stream_policy();
  float cnt = 0.f;
  for (int i = threadIdx.x; i < R; i += blockDim.x) cnt += y_row[i] >= 0 ? 1.f : 0.f;
  cnt = block_sum(cnt, red);
  if (threadIdx.x == 0) bc[0] = cnt;
  __syncthreads();
  const float n_sel = bc[0];
  __syncthreads();
  const float inv_n = n_sel > 0.f ? 1.f / n_sel : 0.f;
  constexpr int U = 2;                 // 2 x (16 B of z + 16-byte window of y) in flight per thread
  for (int r = blockIdx.x; r < R; r += gridDim.x) {
    const T* z = zl + (int64_t)r * z_stride;
    T* d = dz ? dz + (int64_t)r * z_stride : nullptr;
    const int yr = y_row[r];
    const int head = head_elems(z, V);           // chunks are aligned for z (and dz); y is read through windows
    const int body = (V - head) / 8;
    const int tail0 = head + body * 8;
    if (yr < 0) {
      if (threadIdx.x == 0) row_loss[r] = 0.f;
      if (d) {
        const float zero[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        for (int i = threadIdx.x; i < head; i += blockDim.x) d[i] = from_f32<T>(0.f);
        for (int c = threadIdx.x; c < body; c += blockDim.x) Chunk<T>::store(d + head + c * 8, zero);
        for (int i = tail0 + threadIdx.x; i < V; i += blockDim.x) d[i] = from_f32<T>(0.f);
      }
      continue;
    }
    const T* y = yl + (int64_t)yr * y_stride;
    KlAcc k = {-INFINITY, 0.f, -INFINITY, 0.f, 0.f};
    for (int i = threadIdx.x; i < head; i += blockDim.x) kl_add(k, to_f32(z[i]), to_f32(y[i]));
    for (int c0 = threadIdx.x; c0 < body; c0 += blockDim.x * U) {
      float fz[U][8], fy[U][8];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int c = c0 + u * blockDim.x;
        if (c < body) {
          Raw8<T> rz;
          rz.load(z + head + c * 8, keep);                    // both rows come back in pass 2: keep them in L2
          rz.unpack(fz[u]);
          load8_window(y + head + c * 8, fy[u], keep);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int c = c0 + u * blockDim.x;
        if (c < body) {
          // chunk-level rescale: one max / one exp per 8 elements for each running sum
          float cz = fz[u][0], cy = fy[u][0];
#pragma unroll
          for (int e = 1; e < 8; ++e) {
            cz = fmaxf(cz, fz[u][e]);
            cy = fmaxf(cy, fy[u][e]);
          }
          if (cz > k.mz) {
            k.sz *= __expf(k.mz - cz);
            k.mz = cz;
          }
          if (cy > k.my) {
            const float f = __expf(k.my - cy);
            k.sy *= f;
            k.ay *= f;
            k.my = cy;
          }
          constexpr float L2E = 1.4426950408889634f;
          const float nmz = -k.mz * L2E, nmy = -k.my * L2E;     // exp(x - m) = 2^(x log2e - m log2e): one FFMA + one MUFU
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            k.sz += fast_ex2(fmaf(fz[u][e], L2E, nmz));
            const float ey = fast_ex2(fmaf(fy[u][e], L2E, nmy));
            k.sy += ey;
            k.ay = fmaf(ey, fy[u][e] - fz[u][e], k.ay);
          }
        }
      }
    }
    for (int i = tail0 + threadIdx.x; i < V; i += blockDim.x) kl_add(k, to_f32(z[i]), to_f32(y[i]));
    // block combine: maxima first, then the rescaled sums
    float gz = k.mz, gy = k.my;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      gz = fmaxf(gz, __shfl_xor_sync(0xffffffffu, gz, o));
      gy = fmaxf(gy, __shfl_xor_sync(0xffffffffu, gy, o));
    }
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = gz;
    __syncthreads();
    gz = red[0];
    for (int wi = 1; wi < (int)(blockDim.x >> 5); ++wi) gz = fmaxf(gz, red[wi]);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = gy;
    __syncthreads();
    gy = red[0];
    for (int wi = 1; wi < (int)(blockDim.x >> 5); ++wi) gy = fmaxf(gy, red[wi]);
    const float fzs = (k.mz == -INFINITY) ? 0.f : __expf(k.mz - gz), fys = (k.my == -INFINITY) ? 0.f : __expf(k.my - gy);
    const float sz = block_sum(k.sz * fzs, red);
    if (threadIdx.x == 0) bc[1] = sz;
    const float sy = block_sum(k.sy * fys, red);
    if (threadIdx.x == 0) bc[2] = sy;
    const float ay = block_sum(k.ay * fys, red);
    if (threadIdx.x == 0) bc[3] = ay;
    __syncthreads();
    const float Sz = bc[1], Sy = bc[2], Ay = bc[3];
    const float lse_z = gz + __logf(Sz), lse_y = gy + __logf(Sy);
    if (threadIdx.x == 0) row_loss[r] = Ay / Sy - lse_y + lse_z;
    if (d) {
      const float gs = grad_scale * inv_n;
      const float nlz2 = -lse_z * 1.4426950408889634f, nly2 = -lse_y * 1.4426950408889634f;
      // pass 2 (rows come back from L2).  z may alias dz: every element is read before the same thread writes it.
      for (int i = threadIdx.x; i < head; i += blockDim.x)
        d[i] = from_f32<T>(gs * (__expf(to_f32(z[i]) - lse_z) - __expf(to_f32(y[i]) - lse_y)));
      for (int c0 = threadIdx.x; c0 < body; c0 += blockDim.x * U) {
        float fz[U][8], fy[U][8];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int c = c0 + u * blockDim.x;
          if (c < body) {
            Raw8<T> rz;
            rz.load(z + head + c * 8, stream);
            rz.unpack(fz[u]);
            load8_window(y + head + c * 8, fy[u], stream);
          }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int c = c0 + u * blockDim.x;
          if (c < body) {
#pragma unroll
            for (int e = 0; e < 8; ++e)
              fz[u][e] = gs * (fast_ex2(fmaf(fz[u][e], 1.4426950408889634f, nlz2)) - fast_ex2(fmaf(fy[u][e], 1.4426950408889634f, nly2)));
            Raw8<T>::store(d + head + c * 8, fz[u], stream);
          }
        }
      }
      for (int i = tail0 + threadIdx.x; i < V; i += blockDim.x)
        d[i] = from_f32<T>(gs * (__expf(to_f32(z[i]) - lse_z) - __expf(to_f32(y[i]) - lse_y)));
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    __threadfence();
    is_last = atomicAdd(counter, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (is_last) {
    __threadfence();
    float tot = 0.f;
    for (int i = threadIdx.x; i < R; i += blockDim.x) tot += __ldcg(row_loss + i);
    tot = block_sum(tot, red);
    if (threadIdx.x == 0) {
      loss[0] = tot * inv_n;
      n_rows[0] = (int32_t)(n_sel + 0.5f);
      *counter = 0u;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// KL, register-resident variant (bf16) — the construction of ce_fwd_bwd_reg_kernel with two rows per item: a cluster of
// CL CTAs x 256 threads owns a pair (z row, matching y row); each CTA's contiguous run of the NEXT pair is brought into two
// 40 KB shared-memory stages by bulk async copies while the current pair is being worked on.  z chunks are 16-byte
// aligned; the y row starts at an unrelated alignment (V is odd for the real vocabularies), so its stage holds the aligned
// superset of the run and a thread assembles its chunk from two neighbouring 16-byte vectors (byte funnel shift).  Stage ->
// registers (raw bf16, 2 x 4 registers per chunk, <= KL_MAXC chunks per thread, 128 registers per thread, two CTAs per
// SM), (m_z, s_z, m_y, s_y, A) reduced inside the CTA, five floats exchanged through distributed shared memory, gradient
// from the same registers.  Both exponentials are evaluated again in the gradient pass (softmax(z) - softmax(y) tends to
// zero as training converges: a bf16-rounded e would put a 2^-9 relative error of EACH term into the difference), so the
// kernel pays 4 MUFU per element pair: at 16 / clk / SM that alone is 6.7 TB/s of algorithmic bytes, i.e. the MUFU pipe
// and HBM are co-limiting.  dram read == algorithmic: 2 rows in, 1 row out.  CL = smallest of 1 / 2 / 4 / 8 with V <= CL * 20 480.
constexpr int KL_MAXC = 10;
constexpr int KL_NT = 256;
constexpr int KL_STAGE_Z = KL_MAXC * KL_NT * 16, KL_STAGE_Y = KL_STAGE_Z + 16;

struct __align__(16) KlPart {
  float mz, sz, my, sy, ay, pad0, pad1, pad2;
};
__device__ __forceinline__ void st_async_f4(uint32_t remote_dst, float a, float b, float c, float d, uint32_t remote_bar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.f32 [%0], {%1, %2, %3, %4}, [%5];"
               ::"r"(remote_dst), "f"(a), "f"(b), "f"(c), "f"(d), "r"(remote_bar)
               : "memory");
}
__device__ __forceinline__ void st_async_f1(uint32_t remote_dst, float a, uint32_t remote_bar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.f32 [%0], %1, [%2];"
               ::"r"(remote_dst), "f"(a), "r"(remote_bar)
               : "memory");
}
// combine N <= 32 partials of a row pair (s_z relative to m_z; s_y and A relative to m_y) held in shared memory: lane i
// takes partial i, maxima and rescaled sums go through warp shuffles (see ms_combine)
template <int N>
__device__ __forceinline__ void kl_combine(const KlPart* v, int lane, KlPart& t) {
  constexpr float L2E = 1.4426950408889634f;
  float4 a = make_float4(-INFINITY, 0.f, -INFINITY, 0.f);
  float ay = 0.f;
  if (lane < N) {
    a = *reinterpret_cast<const float4*>(&v[lane]);
    ay = v[lane].ay;
  }
  float Mz = a.x, My = a.z;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    Mz = fmaxf(Mz, __shfl_xor_sync(0xffffffffu, Mz, o));
    My = fmaxf(My, __shfl_xor_sync(0xffffffffu, My, o));
  }
  const float nMz = (Mz == -INFINITY) ? 0.f : -Mz * L2E, nMy = (My == -INFINITY) ? 0.f : -My * L2E;
  const float f = fast_ex2(fmaf(a.z, L2E, nMy));
  t.mz = Mz;
  t.my = My;
  t.sz = warp_sum(a.y * fast_ex2(fmaf(a.x, L2E, nMz)));
  t.sy = warp_sum(a.w * f);
  t.ay = warp_sum(ay * f);
}

template <int CL>
__global__ void __launch_bounds__(KL_NT, 2) kl_fwd_bwd_reg_kernel(const __nv_bfloat16* __restrict__ zl, int64_t z_stride,
                                                                  const __nv_bfloat16* __restrict__ yl, int64_t y_stride,
                                                                  const int32_t* __restrict__ y_row, __nv_bfloat16* dz,
                                                                  float grad_scale, int R, int V, int32_t* __restrict__ n_rows,
                                                                  float* __restrict__ loss, float* __restrict__ row_loss,
                                                                  unsigned int* __restrict__ counter) {
  using T = __nv_bfloat16;
  constexpr int NW = KL_NT / 32;
  extern __shared__ __align__(128) unsigned char kl_stage[];   // [KL_STAGE_Z] z run | [KL_STAGE_Y] aligned superset of the y run
  unsigned char* stage_z = kl_stage;
  unsigned char* stage_y = kl_stage + KL_STAGE_Z;
  __shared__ float red[32];
  __shared__ float bc[1];
  __shared__ KlPart wred[2][NW];           // per-warp partials of a row pair, double-buffered by processed-row parity
  __shared__ KlPart slots[2][CL];          // partials of every CTA of the cluster, same double buffer (filled by st.async)
  __shared__ uint64_t full_bar;            // both stages hold the pair they were armed for
  __shared__ uint64_t xbar[2];             // slots[par] are complete
  __shared__ bool is_last;
  constexpr float L2E = 1.4426950408889634f;
  const uint64_t stream = l2_stream_policy();
  const uint32_t rank = CL > 1 ? cluster_rank() : 0u;
  const int cluster_id = blockIdx.x / CL, n_clusters = gridDim.x / CL;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) {
    ptx::mbar_init(&full_bar, 1);
    ptx::mbar_init(&xbar[0], 1);
    ptx::mbar_init(&xbar[1], 1);
    ptx::fence_barrier_init();
  }
  // thread 0: have this CTA's run of pair r brought into the stages (unselected rows are never staged)
  auto stage_pair = [&](int r) {
    if (r >= R) return;
    const int yr = y_row[r];
    if (yr < 0) return;
    const T* z = zl + (int64_t)r * z_stride;
    const T* y = yl + (int64_t)yr * y_stride;
    const int head = head_elems(z, V);
    int first, count;
    cta_run<CL>((V - head) / 8, rank, first, count);
    if (count > 0) {
      const uintptr_t ya = reinterpret_cast<uintptr_t>(y + head + first * 8);
      const uint32_t yo = (uint32_t)(ya & 15), zb = (uint32_t)count * 16u, yb = zb + (yo ? 16u : 0u);
      ptx::mbar_expect_tx(&full_bar, zb + yb);
      bulk_load_to_stage(stage_z, z + head + first * 8, zb, &full_bar, stream);
      bulk_load_to_stage(stage_y, reinterpret_cast<const void*>(ya - yo), yb, &full_bar, stream);
    }
  };
  // the <= 7 + 7 element pairs around the aligned body (one each for the first threads of CTA 0), fetched one pair ahead
  auto edge_index = [&](int head, int tail0) {
    return rank != 0 ? -1 : (tid < head ? tid : (tid >= 32 && tid < 40 && tail0 + (tid - 32) < V ? tail0 + tid - 32 : -1));
  };
  auto edge_fetch = [&](int r, float& ez, float& ey) {
    ez = ey = -INFINITY;
    if (rank != 0 || r >= R) return;
    const int yr = y_row[r];
    if (yr < 0) return;
    const T* z = zl + (int64_t)r * z_stride;
    const int head = head_elems(z, V);
    const int ei = edge_index(head, head + ((V - head) / 8) * 8);
    if (ei >= 0) {
      ez = to_f32(z[ei]);
      ey = to_f32(yl[(int64_t)yr * y_stride + ei]);
    }
  };
  float cnt = 0.f;
  for (int i = tid; i < R; i += blockDim.x) cnt += y_row[i] >= 0 ? 1.f : 0.f;
  cnt = block_sum(cnt, red);                // (its barriers also publish the mbarrier init inside the CTA)
  if (tid == 0) {
    bc[0] = cnt;
    stage_pair(cluster_id);
  }
  float edge_z, edge_y;
  edge_fetch(cluster_id, edge_z, edge_y);
  cluster_sync_all<CL>();                   // every CTA's exchange barriers exist before anyone sends to them
  const float n_sel = bc[0];
  const float inv_n = n_sel > 0.f ? 1.f / n_sel : 0.f;
  const float gs = grad_scale * inv_n;
  uint32_t done = 0, staged = 0;
  for (int r = cluster_id; r < R; r += n_clusters) {
    const T* z = zl + (int64_t)r * z_stride;
    T* d = dz ? dz + (int64_t)r * z_stride : nullptr;
    const int yr = y_row[r];
    const int head = head_elems(z, V);           // chunks are aligned for z (and dz); y is read through windows
    const int body = (V - head) / 8;
    const int tail0 = head + body * 8;
    int first, count;
    cta_run<CL>(body, rank, first, count);
    if (yr < 0) {                                // (uniform over the cluster: no exchange on this path)
      if (tid == 0) stage_pair(r + n_clusters);  // the stages are free: the previous pair left them before its arithmetic
      edge_fetch(r + n_clusters, edge_z, edge_y);
      if (rank == 0 && tid == 0) row_loss[r] = 0.f;
      if (d) {
        const float zero[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        const int gtid = rank * KL_NT + tid;
        constexpr int GT = CL * KL_NT;
        for (int i = gtid; i < head; i += GT) d[i] = from_f32<T>(0.f);
        for (int c = gtid; c < body; c += GT) Chunk<T>::store(d + head + c * 8, zero);
        for (int i = tail0 + gtid; i < V; i += GT) d[i] = from_f32<T>(0.f);
      }
      continue;
    }
    const unsigned yo = (unsigned)(reinterpret_cast<uintptr_t>(yl + (int64_t)yr * y_stride + head + first * 8) & 15);
    // ---- both rows, once: stages -> registers, then the stages go to the next pair ----
    Raw8<T> rz[KL_MAXC], ry[KL_MAXC];
    if (count > 0) {
      stage_wait(&full_bar, staged & 1);
      ++staged;
    }
#pragma unroll
    for (int u = 0; u < KL_MAXC; ++u) {
      const int c = tid + u * KL_NT;
      if (c < count) {
        rz[u].u = *reinterpret_cast<const uint4*>(stage_z + (size_t)c * 16);
        const uint4 lo = *reinterpret_cast<const uint4*>(stage_y + (size_t)c * 16);
        ry[u].u = yo ? window_combine(lo, *reinterpret_cast<const uint4*>(stage_y + (size_t)c * 16 + 16), yo) : lo;
      }
    }
    const int ei = edge_index(head, tail0);
    const float ez = ei >= 0 ? edge_z : -INFINITY, ey = ei >= 0 ? edge_y : -INFINITY;
    __syncthreads();
    if (tid == 0) stage_pair(r + n_clusters);
    edge_fetch(r + n_clusters, edge_z, edge_y);
    float mz = ez, my = ey;
    {
      const __nv_bfloat162 ninf = __halves2bfloat162(__ushort_as_bfloat16(0xFF80u), __ushort_as_bfloat16(0xFF80u));
      __nv_bfloat162 pz = ninf, py = ninf;         // thread maxima on the packed pairs (no unpacking)
#pragma unroll
      for (int u = 0; u < KL_MAXC; ++u) {
        if (tid + u * KL_NT < count) {
          const __nv_bfloat162* hz = reinterpret_cast<const __nv_bfloat162*>(&rz[u].u);
          const __nv_bfloat162* hy = reinterpret_cast<const __nv_bfloat162*>(&ry[u].u);
          pz = __hmax2(__hmax2(pz, hz[0]), __hmax2(hz[1], __hmax2(hz[2], hz[3])));
          py = __hmax2(__hmax2(py, hy[0]), __hmax2(hy[1], __hmax2(hy[2], hy[3])));
        }
      }
      const float2 a = __bfloat1622float2(pz), b = __bfloat1622float2(py);
      mz = fmaxf(mz, fmaxf(a.x, a.y));
      my = fmaxf(my, fmaxf(b.x, b.y));
    }
    float sz = 0.f, sy = 0.f, ay = 0.f;
    {
      const float nmz = (mz == -INFINITY) ? 0.f : -mz * L2E, nmy = (my == -INFINITY) ? 0.f : -my * L2E;
      if (ei >= 0) {
        sz += fast_ex2(fmaf(ez, L2E, nmz));
        const float e = fast_ex2(fmaf(ey, L2E, nmy));
        sy += e;
        ay = fmaf(e, ey - ez, ay);
      }
#pragma unroll
      for (int u = 0; u < KL_MAXC; ++u) {
        if (tid + u * KL_NT < count) {
          float fz[8], fy[8];
          rz[u].unpack(fz);
          ry[u].unpack(fy);
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            sz += fast_ex2(fmaf(fz[e], L2E, nmz));
            const float t = fast_ex2(fmaf(fy[e], L2E, nmy));
            sy += t;
            ay = fmaf(t, fy[e] - fz[e], ay);
          }
        }
      }
    }
    // ---- warp partials -> shared; ONE block barrier; every thread merges the 8 warps; the cluster's by st.async ----
    const int par = done & 1;
    {
      float wz = mz, wy = my;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        wz = fmaxf(wz, __shfl_xor_sync(0xffffffffu, wz, o));
        wy = fmaxf(wy, __shfl_xor_sync(0xffffffffu, wy, o));
      }
      const float fzs = (mz == -INFINITY) ? 0.f : fast_ex2((mz - wz) * L2E), fys = (my == -INFINITY) ? 0.f : fast_ex2((my - wy) * L2E);
      const float a = warp_sum(sz * fzs), b = warp_sum(sy * fys), c = warp_sum(ay * fys);
      if (lane == 0) wred[par][warp] = KlPart{wz, a, wy, b, c, 0.f, 0.f, 0.f};
    }
    __syncthreads();
    KlPart tot;
    kl_combine<NW>(wred[par], lane, tot);
    if constexpr (CL > 1) {
      if (tid == 0) ptx::mbar_expect_tx(&xbar[par], CL * 20u);
      if (tid < CL) {
        const uint32_t dst = map_to_cta(&slots[par][rank], (uint32_t)tid), bar = map_to_cta(&xbar[par], (uint32_t)tid);
        st_async_f4(dst, tot.mz, tot.sz, tot.my, tot.sy, bar);
        st_async_f1(dst + 16, tot.ay, bar);
      }
      stage_wait(&xbar[par], (done >> 1) & 1);
      kl_combine<CL>(slots[par], lane, tot);
    }
    ++done;
    const float lse_z = tot.mz + __logf(tot.sz), lse_y = tot.my + __logf(tot.sy);
    if (rank == 0 && tid == 0) row_loss[r] = tot.ay / tot.sy - lse_y + lse_z;
    // ---- gradient from the same registers ----
    if (d) {
      const float nlz2 = -lse_z * L2E, nly2 = -lse_y * L2E;
      if (ei >= 0) d[ei] = from_f32<T>(gs * (fast_ex2(fmaf(ez, L2E, nlz2)) - fast_ex2(fmaf(ey, L2E, nly2))));
      T* dp = d + head + (size_t)(first + tid) * 8;
#pragma unroll
      for (int u = 0; u < KL_MAXC; ++u) {
        const int c = tid + u * KL_NT;
        if (c < count) {
          float fz[8], fy[8];
          rz[u].unpack(fz);
          ry[u].unpack(fy);
#pragma unroll
          for (int e = 0; e < 8; ++e) fz[e] = gs * (fast_ex2(fmaf(fz[e], L2E, nlz2)) - fast_ex2(fmaf(fy[e], L2E, nly2)));
          Chunk<T>::store(dp + (size_t)u * KL_NT * 8, fz);
        }
      }
    }
  }
  cluster_sync_all<CL>();                   // no CTA leaves while a neighbour may still send to its slots
  if (tid == 0) {
    __threadfence();
    is_last = atomicAdd(counter, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (is_last) {
    __threadfence();
    float tot = 0.f;
    for (int i = tid; i < R; i += blockDim.x) tot += __ldcg(row_loss + i);
    tot = block_sum(tot, red);
    if (tid == 0) {
      loss[0] = tot * inv_n;
      n_rows[0] = (int32_t)(n_sel + 0.5f);
      *counter = 0u;
    }
  }
}

template <int CL>
static int launch_kl_reg(const __nv_bfloat16* z, int64_t z_stride, const __nv_bfloat16* y, int64_t y_stride, const int32_t* y_row,
                         __nv_bfloat16* dz, float grad_scale, int R, int V, int32_t* n_rows, float* loss, float* row_loss,
                         unsigned int* counter, cudaStream_t st) {
  cudaLaunchConfig_t cfg{};
  cfg.blockDim = dim3(KL_NT);
  cfg.dynamicSmemBytes = KL_STAGE_Z + KL_STAGE_Y;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  static int max_clusters = 0;              // persistent clusters: as many as the device co-schedules (see launch_ce_reg)
  if (max_clusters == 0) {
    MIMIC_CUDA_OK(cudaFuncSetAttribute(kl_fwd_bwd_reg_kernel<CL>, cudaFuncAttributeMaxDynamicSharedMemorySize, KL_STAGE_Z + KL_STAGE_Y));
    MIMIC_CUDA_OK(cudaFuncSetAttribute(kl_fwd_bwd_reg_kernel<CL>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                       (int)cudaSharedmemCarveoutMaxShared));
    cfg.gridDim = dim3(CL * num_sms() * 2);
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, kl_fwd_bwd_reg_kernel<CL>, &cfg) != cudaSuccess || n <= 0) n = num_sms() * 2 / CL - 4;
    max_clusters = n > 0 ? n : 1;
  }
  const int n_clusters = max_clusters < R ? max_clusters : R;
  cfg.gridDim = dim3(n_clusters * CL);
  MIMIC_CUDA_OK(cudaLaunchKernelEx(&cfg, kl_fwd_bwd_reg_kernel<CL>, z, z_stride, y, y_stride, y_row, dz, grad_scale, R, V, n_rows,
                                   loss, row_loss, counter));
  return MIMIC_OK;
}

// ---------------------------------------------------------------------------------------------
// Rotary embedding of q and k in one pass (the RoPE step of the replaced attention forwards,
// src/shift_encoder.py:307-312, 417-422, 495-501: q*cos + rotate_half(q)*sin done by ~10 elementwise launches in
// the HF code).  Reads q [B,T,H,d] and k [B,T,Hkv,d] through their strides (the projection outputs, in place of any
// .contiguous()), writes contiguous outputs:  y_lo = x_lo c_lo - x_hi s_lo,  y_hi = x_hi c_hi + x_lo s_hi.
// inverse = the transpose (backward):        dx_lo = g_lo c_lo + g_hi s_hi,  dx_hi = g_hi c_hi - g_lo s_lo.
// Algorithmic bytes: 2 * (H + Hkv) * d * sizeof(T) per token (+ cos/sin once per token, L1-resident across heads).
template <typename T, typename IDX>
__global__ void __launch_bounds__(256) rope_qk_kernel(const T* __restrict__ q, int64_t q_sb, int64_t q_st, int64_t q_sh,
                                                      const T* __restrict__ k, int64_t k_sb, int64_t k_st, int64_t k_sh,
                                                      const T* __restrict__ cosp, const T* __restrict__ sinp, int64_t c_sb,
                                                      int64_t c_st, T* __restrict__ q_out, T* __restrict__ k_out, int B, int Tn,
                                                      int H, int Hkv, int d, int inverse) {
  const int half = d / 2, cpr = half / 8;              // 8-element chunks per half row
  const int items = (H + Hkv) * cpr;
  const IDX total = (IDX)B * (IDX)Tn * (IDX)items;
  // flat (token, head, chunk) index space: every thread of every CTA has work in every round; IDX is 32-bit whenever the
  // index space allows (64-bit divisions cost ~100 instructions each)
  for (IDX idx = (IDX)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (IDX)gridDim.x * blockDim.x) {
    const IDX tok = idx / (IDX)items;
    const int it = (int)(idx - tok * (IDX)items);
    const int b = (int)(tok / (IDX)Tn), t = (int)(tok - (IDX)b * (IDX)Tn);
    const T* cr = cosp + b * c_sb + t * c_st;
    const T* sr = sinp + b * c_sb + t * c_st;
    {
      const int head = it / cpr, c = it - head * cpr;
      const bool is_q = head < H;
      const int hh = is_q ? head : head - H;
      const T* src = is_q ? q + b * q_sb + t * q_st + hh * q_sh : k + b * k_sb + t * k_st + hh * k_sh;
      T* dst = is_q ? q_out + ((int64_t)tok * H + hh) * d : k_out + ((int64_t)tok * Hkv + hh) * d;
      float lo[8], hi[8], cl[8], ch[8], sl[8], sh[8], ol[8], oh[8];
      load8(src + 8 * c, lo);
      load8(src + half + 8 * c, hi);
      load8(cr + 8 * c, cl);
      load8(cr + half + 8 * c, ch);
      load8(sr + 8 * c, sl);
      load8(sr + half + 8 * c, sh);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        if (inverse) {
          ol[e] = fmaf(lo[e], cl[e], hi[e] * sh[e]);
          oh[e] = fmaf(hi[e], ch[e], -lo[e] * sl[e]);
        } else {
          ol[e] = fmaf(lo[e], cl[e], -hi[e] * sl[e]);
          oh[e] = fmaf(hi[e], ch[e], lo[e] * sh[e]);
        }
      }
      Chunk<T>::store(dst + 8 * c, ol);
      Chunk<T>::store(dst + half + 8 * c, oh);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// clip + AdamW on the flat fp32 parameter buffer
// Sum of squares of the (scaled) gradient, DETERMINISTIC: per-CTA partials (fixed strided order inside a CTA) go to the
// workspace, the last CTA to arrive adds them in index order.  A float atomicAdd per CTA made the clip coefficient depend
// on the arrival order: two data-parallel ranks holding bit-identical all-reduced gradients then clipped by factors one
// ulp apart and their parameters drifted apart (tests/test_gpu_dp_nccl.py caught it on one box of the pool).
//   ws[0] = sum (result)   ws[1] = arrival counter (uint32, self-resetting)   ws[2 ...] = per-CTA partials
__global__ void __launch_bounds__(256) grad_sumsq_kernel(float* __restrict__ grad, int64_t n, float grad_mul, float* __restrict__ ws) {
  __shared__ float red[32];
  __shared__ bool is_last;
  float acc = 0.f;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float g = grad[i] * grad_mul;
    grad[i] = g;
    acc += g * g;
  }
  acc = block_sum(acc, red);
  unsigned int* counter = reinterpret_cast<unsigned int*>(ws + 1);
  if (threadIdx.x == 0) {
    ws[2 + blockIdx.x] = acc;
    __threadfence();
    is_last = atomicAdd(counter, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (is_last) {
    __threadfence();
    float tot = 0.f;
    for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) tot += __ldcg(ws + 2 + i);
    tot = block_sum(tot, red);
    if (threadIdx.x == 0) {
      ws[0] = tot;
      *counter = 0u;
    }
  }
}

__global__ void __launch_bounds__(256) adamw_kernel(float* __restrict__ param, const float* __restrict__ grad, float* __restrict__ m,
                                                    float* __restrict__ v, const float* __restrict__ wd, int64_t n,
                                                    const float* __restrict__ hyper, float beta1, float beta2, float eps,
                                                    float max_norm, const float* __restrict__ sumsq,
                                                    const float* __restrict__ lr_mul) {
  const float lr0 = hyper[0], bc1 = hyper[1], bc2_sqrt = hyper[2];
  float clip = 1.f;
  if (max_norm > 0.f) {
    const float c = max_norm / (sqrtf(sumsq[0]) + 1e-6f);   // torch.nn.utils.clip_grad_norm_
    clip = c < 1.f ? c : 1.f;
  }
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float g = grad[i] * clip;
    const float lr = lr_mul ? lr0 * lr_mul[i] : lr0;   // second learning rate of the reference's scale group
    float p = param[i];
    p *= 1.f - lr * wd[i];
    const float mi = beta1 * m[i] + (1.f - beta1) * g;
    const float vi = beta2 * v[i] + (1.f - beta2) * g * g;
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    param[i] = p - (lr / bc1) * (mi / denom);
  }
}

}  // namespace mimic

using namespace mimic;

extern "C" int mimic_record_gather(const void* src, int64_t src_row_stride, int32_t src_dtype, const int32_t* src_row,
                                   void* dst, int32_t dst_dtype, int32_t N, int32_t D, void* stream) {
  MIMIC_CHECK_ARG(src && src_row && dst, "NULL argument");
  MIMIC_CHECK_ARG(N >= 0 && D > 0 && D % 8 == 0 && src_row_stride % 8 == 0, "D and row stride must be multiples of 8");
  if (N == 0) return MIMIC_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int grid = N < num_sms() * 8 ? N : num_sms() * 8;
  if (src_dtype == MIMIC_BF16 && dst_dtype == MIMIC_BF16)
    record_gather_kernel<__nv_bfloat16, __nv_bfloat16><<<grid, 256, 0, st>>>(
        (const __nv_bfloat16*)src, src_row_stride, src_row, (__nv_bfloat16*)dst, N, D);
  else if (src_dtype == MIMIC_F32 && dst_dtype == MIMIC_F32)
    record_gather_kernel<float, float><<<grid, 256, 0, st>>>((const float*)src, src_row_stride, src_row, (float*)dst, N, D);
  else if (src_dtype == MIMIC_F32 && dst_dtype == MIMIC_BF16)
    record_gather_kernel<float, __nv_bfloat16><<<grid, 256, 0, st>>>((const float*)src, src_row_stride, src_row,
                                                                    (__nv_bfloat16*)dst, N, D);
  else {
    set_error("record_gather: unsupported dtype pair %d -> %d", src_dtype, dst_dtype);
    return MIMIC_ERR_UNSUPPORTED;
  }
  MIMIC_LAUNCH_OK("record_gather_kernel");
  return MIMIC_OK;
}

extern "C" size_t mimic_mse_workspace(int32_t L, int32_t rows, int32_t D) {
  (void)D;
  return (size_t)(mse_grid(L, rows) + 4) * sizeof(float);   // per-CTA partials + the arrival counter
}

extern "C" int mimic_mse_fwd_bwd(const void* const* h_ptrs, int64_t h_row_stride, int32_t h_dtype, const void* hp,
                                 int32_t hp_dtype, const int32_t* row_to_compact, const float* row_w,
                                 void* const* dh_ptrs, float grad_scale, int32_t L, int32_t rows, int32_t N,
                                 int32_t D, float* loss, void* workspace, size_t workspace_bytes, void* stream) {
  MIMIC_CHECK_ARG(h_ptrs && hp && row_to_compact && row_w && loss && workspace, "NULL argument");
  MIMIC_CHECK_ARG(L > 0 && rows > 0 && N >= 0 && D > 0 && D % 8 == 0 && h_row_stride % 8 == 0, "bad shape");
  MIMIC_CHECK_ARG(h_dtype == hp_dtype, "H and H' must have the same storage type (got %d, %d)", h_dtype, hp_dtype);
  MIMIC_CHECK_ARG(workspace_bytes >= mimic_mse_workspace(L, rows, D), "workspace too small");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int grid = mse_grid(L, rows);
  float* partial = static_cast<float*>(workspace) + 4;
  unsigned int* counter = static_cast<unsigned int*>(workspace);   // zero on first use; the kernel leaves it zero
  static const bool no_pipe = getenv("MIMIC_MSE_NO_PIPE") != nullptr;      // A/B only
  if (h_dtype == MIMIC_BF16 && D <= 4096 && L <= 128 && !no_pipe)
    mse_fwd_bwd_pipe_kernel<<<grid, 256, 0, st>>>(h_ptrs, h_row_stride, (const __nv_bfloat16*)hp, row_to_compact, row_w,
                                                  dh_ptrs, grad_scale, L, rows, N, D, partial, counter, loss);
  else if (h_dtype == MIMIC_BF16)
    mse_fwd_bwd_kernel<__nv_bfloat16, 2><<<grid, 256, 0, st>>>(h_ptrs, h_row_stride, (const __nv_bfloat16*)hp, row_to_compact,
                                                               row_w, dh_ptrs, grad_scale, L, rows, N, D, partial, counter, loss);
  else if (h_dtype == MIMIC_F32)
    mse_fwd_bwd_kernel<float, 2><<<grid, 256, 0, st>>>(h_ptrs, h_row_stride, (const float*)hp, row_to_compact, row_w,
                                                       dh_ptrs, grad_scale, L, rows, N, D, partial, counter, loss);
  else {
    set_error("mse: unsupported dtype %d", h_dtype);
    return MIMIC_ERR_UNSUPPORTED;
  }
  MIMIC_LAUNCH_OK("mse_fwd_bwd_kernel");
  return MIMIC_OK;
}

extern "C" size_t mimic_ce_workspace(int32_t R) { return (size_t)(R + 4) * sizeof(float); }

extern "C" int mimic_ce_fwd_bwd(const void* logits, int64_t logits_row_stride, int32_t dtype, const int64_t* labels,
                                void* dlogits, float grad_scale, int32_t R, int32_t V, float* loss, int32_t* n_kept,
                                void* workspace, size_t workspace_bytes, void* stream) {
  MIMIC_CHECK_ARG(logits && labels && loss && n_kept && workspace, "NULL argument");
  MIMIC_CHECK_ARG(R > 0 && V > 0 && logits_row_stride >= V, "bad shape");
  MIMIC_CHECK_ARG(workspace_bytes >= mimic_ce_workspace(R), "workspace too small");
  if (dlogits)
    MIMIC_CHECK_ARG((reinterpret_cast<uintptr_t>(dlogits) & 31u) == (reinterpret_cast<uintptr_t>(logits) & 31u),
                    "logits and dlogits must share their alignment modulo 32 bytes");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  unsigned int* counter = static_cast<unsigned int*>(workspace);   // zero on first use; the kernel leaves it zero
  float* row_loss = static_cast<float*>(workspace) + 4;
  const int grid = R < num_sms() * 4 ? R : num_sms() * 4;   // measured: more CTAs beat a smaller L2 footprint
  static const bool two_pass_only = getenv("MIMIC_CE_TWO_PASS") != nullptr;      // A/B only
  if (dtype == MIMIC_BF16 && !two_pass_only && V <= 8 * 256 * CE_MAXC * 8) {
    // register-resident row: one 256-thread CTA up to V = 20 480; one 512-thread CTA up to 40 960 (the 32 003-token
    // vocabularies: no cluster exchange at all, measured 0.28 vs 0.25 of the HBM peak at 256 rows); beyond that the smallest
    // cluster of 256-thread CTAs whose registers hold the row (0.84 vs 0.81 for 4 x 512 at 1540 x 152 000)
    const __nv_bfloat16* lg = (const __nv_bfloat16*)logits;
    __nv_bfloat16* dl = (__nv_bfloat16*)dlogits;
    constexpr int64_t cap1 = (int64_t)256 * CE_MAXC * 8;
    int rc;
#define MIMIC_CE_LAUNCH(CLV, NTV) launch_ce_reg<CLV, NTV>(lg, logits_row_stride, labels, dl, grad_scale, R, V, n_kept, loss, row_loss, counter, st)
    if (V <= cap1) rc = MIMIC_CE_LAUNCH(1, 256);
    else if (V <= 2 * cap1) rc = MIMIC_CE_LAUNCH(1, 512);
    else if (V <= 4 * cap1) rc = MIMIC_CE_LAUNCH(4, 256);
    else rc = MIMIC_CE_LAUNCH(8, 256);
#undef MIMIC_CE_LAUNCH
    if (rc) return rc;
    MIMIC_LAUNCH_OK("ce_fwd_bwd_reg_kernel");
    return MIMIC_OK;
  }
  if (dtype == MIMIC_BF16)
    ce_fwd_bwd_kernel<__nv_bfloat16><<<grid, 512, 0, st>>>((const __nv_bfloat16*)logits, logits_row_stride, labels,
                                                           (__nv_bfloat16*)dlogits, grad_scale, R, V, n_kept, loss, row_loss, counter);
  else if (dtype == MIMIC_F32)
    ce_fwd_bwd_kernel<float><<<grid, 512, 0, st>>>((const float*)logits, logits_row_stride, labels, (float*)dlogits,
                                                   grad_scale, R, V, n_kept, loss, row_loss, counter);
  else {
    set_error("ce: unsupported dtype %d", dtype);
    return MIMIC_ERR_UNSUPPORTED;
  }
  MIMIC_LAUNCH_OK("ce_fwd_bwd_kernel");
  return MIMIC_OK;
}

extern "C" int mimic_cos_fwd_bwd(const void* const* h_ptrs, int64_t h_row_stride, int32_t h_dtype, const void* hp,
                                 int32_t hp_dtype, const int32_t* row_to_compact, const float* row_w,
                                 void* const* dh_ptrs, float grad_scale, int32_t L, int32_t rows, int32_t N,
                                 int32_t D, float* loss, void* workspace, size_t workspace_bytes, void* stream) {
  MIMIC_CHECK_ARG(h_ptrs && hp && row_to_compact && row_w && loss && workspace, "NULL argument");
  MIMIC_CHECK_ARG(L > 0 && rows > 0 && N >= 0 && D > 0 && D % 8 == 0 && h_row_stride % 8 == 0, "bad shape");
  MIMIC_CHECK_ARG(D <= 256 * 8 * 4, "cos: D up to 8192 (one row of H and H' lives in the CTA's registers), got %d", D);
  MIMIC_CHECK_ARG(h_dtype == hp_dtype, "H and H' must have the same storage type (got %d, %d)", h_dtype, hp_dtype);
  MIMIC_CHECK_ARG(workspace_bytes >= mimic_mse_workspace(L, rows, D), "workspace too small");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int grid = mse_grid(L, rows);
  float* partial = static_cast<float*>(workspace) + 4;
  unsigned int* counter = static_cast<unsigned int*>(workspace);   // zero on first use; the kernel leaves it zero
  // (a warp-per-row variant that re-reads the row from L2 for the gradient pass was measured: 1.9 TB/s against 4.25 TB/s)
  static const int variant = getenv("MIMIC_COS_VARIANT") ? atoi(getenv("MIMIC_COS_VARIANT")) : 1;   // 1: cp.async pipeline, 0: registers
  constexpr int COS_NS = 2;       // measured: 3 stages (48 KB, 4 CTAs/SM) 4.24 TB/s, 2 stages 4.56 TB/s
  const size_t pipe_smem = (size_t)2 * COS_NS * D * (h_dtype == MIMIC_BF16 ? 2 : 4);
  if (variant == 1 && pipe_smem <= 96 * 1024 && (h_dtype == MIMIC_BF16 || h_dtype == MIMIC_F32)) {
    // every CTA resident at once (a partial second wave is a tail): CTAs per SM by shared memory, at most 8
    const int per_sm = (int)std::min<size_t>(8, (size_t)(220 * 1024) / (pipe_smem + 1024));
    const int64_t cap = (int64_t)num_sms() * per_sm, n_items = (int64_t)L * rows;
    const int pgrid = (int)(n_items < cap ? n_items : cap);
    if (h_dtype == MIMIC_BF16) {
      static size_t conf = 0;
      if (pipe_smem > conf) {
        MIMIC_CUDA_OK(cudaFuncSetAttribute(cos_fwd_bwd_pipe_kernel<__nv_bfloat16, COS_NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pipe_smem));
        conf = pipe_smem;
      }
      cos_fwd_bwd_pipe_kernel<__nv_bfloat16, COS_NS><<<pgrid, 256, pipe_smem, st>>>(h_ptrs, h_row_stride, (const __nv_bfloat16*)hp, row_to_compact,
                                                                          row_w, dh_ptrs, grad_scale, L, rows, N, D, partial, counter, loss);
    } else {
      static size_t conf = 0;
      if (pipe_smem > conf) {
        MIMIC_CUDA_OK(cudaFuncSetAttribute(cos_fwd_bwd_pipe_kernel<float, COS_NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pipe_smem));
        conf = pipe_smem;
      }
      cos_fwd_bwd_pipe_kernel<float, COS_NS><<<pgrid, 256, pipe_smem, st>>>(h_ptrs, h_row_stride, (const float*)hp, row_to_compact, row_w,
                                                                   dh_ptrs, grad_scale, L, rows, N, D, partial, counter, loss);
    }
    MIMIC_LAUNCH_OK("cos_fwd_bwd_pipe_kernel");
    return MIMIC_OK;
  }
  const bool small = D <= 256 * 8 * 2;
  if (h_dtype == MIMIC_BF16 && small)
    cos_fwd_bwd_kernel<__nv_bfloat16, 2><<<grid, 256, 0, st>>>(h_ptrs, h_row_stride, (const __nv_bfloat16*)hp, row_to_compact,
                                                               row_w, dh_ptrs, grad_scale, L, rows, N, D, partial, counter, loss);
  else if (h_dtype == MIMIC_BF16)
    cos_fwd_bwd_kernel<__nv_bfloat16, 4><<<grid, 256, 0, st>>>(h_ptrs, h_row_stride, (const __nv_bfloat16*)hp, row_to_compact,
                                                               row_w, dh_ptrs, grad_scale, L, rows, N, D, partial, counter, loss);
  else if (h_dtype == MIMIC_F32 && small)
    cos_fwd_bwd_kernel<float, 2><<<grid, 256, 0, st>>>(h_ptrs, h_row_stride, (const float*)hp, row_to_compact, row_w, dh_ptrs,
                                                       grad_scale, L, rows, N, D, partial, counter, loss);
  else if (h_dtype == MIMIC_F32)
    cos_fwd_bwd_kernel<float, 4><<<grid, 256, 0, st>>>(h_ptrs, h_row_stride, (const float*)hp, row_to_compact, row_w, dh_ptrs,
                                                       grad_scale, L, rows, N, D, partial, counter, loss);
  else {
    set_error("cos: unsupported dtype %d", h_dtype);
    return MIMIC_ERR_UNSUPPORTED;
  }
  MIMIC_LAUNCH_OK("cos_fwd_bwd_kernel");
  return MIMIC_OK;
}

extern "C" int mimic_kl_fwd_bwd(const void* z, int64_t z_row_stride, const void* y, int64_t y_row_stride, int32_t dtype,
                                const int32_t* y_row, void* dz, float grad_scale, int32_t R, int32_t V, float* loss,
                                int32_t* n_rows, void* workspace, size_t workspace_bytes, void* stream) {
  MIMIC_CHECK_ARG(z && y && y_row && loss && n_rows && workspace, "NULL argument");
  MIMIC_CHECK_ARG(R > 0 && V > 0 && z_row_stride >= V && y_row_stride >= V, "bad shape");
  MIMIC_CHECK_ARG(workspace_bytes >= mimic_ce_workspace(R), "workspace too small");
  if (dz)
    MIMIC_CHECK_ARG((reinterpret_cast<uintptr_t>(dz) & 31u) == (reinterpret_cast<uintptr_t>(z) & 31u),
                    "z and dz must share their alignment modulo 32 bytes");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  unsigned int* counter = static_cast<unsigned int*>(workspace);   // zero on first use; the kernel leaves it zero
  float* row_loss = static_cast<float*>(workspace) + 4;
  const int grid = R < num_sms() * 4 ? R : num_sms() * 4;
  static const bool two_pass = getenv("MIMIC_KL_TWO_PASS") != nullptr;     // (A/B switch, read once)
  constexpr int cap1 = KL_NT * KL_MAXC * 8;
  if (dtype == MIMIC_BF16 && V <= 8 * cap1 && !two_pass) {
    const __nv_bfloat16 *zz = (const __nv_bfloat16*)z, *yy = (const __nv_bfloat16*)y;
    __nv_bfloat16* dd = (__nv_bfloat16*)dz;
    int rc;
    if (V <= cap1) rc = launch_kl_reg<1>(zz, z_row_stride, yy, y_row_stride, y_row, dd, grad_scale, R, V, n_rows, loss, row_loss, counter, st);
    else if (V <= 2 * cap1) rc = launch_kl_reg<2>(zz, z_row_stride, yy, y_row_stride, y_row, dd, grad_scale, R, V, n_rows, loss, row_loss, counter, st);
    else if (V <= 4 * cap1) rc = launch_kl_reg<4>(zz, z_row_stride, yy, y_row_stride, y_row, dd, grad_scale, R, V, n_rows, loss, row_loss, counter, st);
    else rc = launch_kl_reg<8>(zz, z_row_stride, yy, y_row_stride, y_row, dd, grad_scale, R, V, n_rows, loss, row_loss, counter, st);
    if (rc != MIMIC_OK) return rc;
    MIMIC_LAUNCH_OK("kl_fwd_bwd_reg_kernel");
    return MIMIC_OK;
  }
  if (dtype == MIMIC_BF16)
    kl_fwd_bwd_kernel<__nv_bfloat16><<<grid, 512, 0, st>>>((const __nv_bfloat16*)z, z_row_stride, (const __nv_bfloat16*)y,
                                                           y_row_stride, y_row, (__nv_bfloat16*)dz, grad_scale, R, V, n_rows,
                                                           loss, row_loss, counter);
  else if (dtype == MIMIC_F32)
    kl_fwd_bwd_kernel<float><<<grid, 512, 0, st>>>((const float*)z, z_row_stride, (const float*)y, y_row_stride, y_row,
                                                   (float*)dz, grad_scale, R, V, n_rows, loss, row_loss, counter);
  else {
    set_error("kl: unsupported dtype %d", dtype);
    return MIMIC_ERR_UNSUPPORTED;
  }
  MIMIC_LAUNCH_OK("kl_fwd_bwd_kernel");
  return MIMIC_OK;
}

extern "C" int mimic_rope_qk(const void* q, int64_t q_sb, int64_t q_st, int64_t q_sh, const void* k, int64_t k_sb, int64_t k_st,
                             int64_t k_sh, const void* cosp, const void* sinp, int64_t c_sb, int64_t c_st, void* q_out,
                             void* k_out, int32_t dtype, int32_t B, int32_t T, int32_t H, int32_t Hkv, int32_t d,
                             int32_t inverse, void* stream) {
  MIMIC_CHECK_ARG(q && k && cosp && sinp && q_out && k_out, "NULL argument");
  MIMIC_CHECK_ARG(B > 0 && T > 0 && H > 0 && Hkv > 0 && d > 0 && d % 16 == 0, "bad shape (head_dim must be a multiple of 16)");
  const int64_t al = dtype == MIMIC_BF16 ? 8 : 4;      // 16-byte rows
  MIMIC_CHECK_ARG(q_sb % al == 0 && q_st % al == 0 && q_sh % al == 0 && k_sb % al == 0 && k_st % al == 0 && k_sh % al == 0 &&
                      c_sb % al == 0 && c_st % al == 0,
                  "strides must keep rows 16-byte aligned");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int64_t work = (int64_t)B * T * (H + Hkv) * (d / 16);
  const int64_t ctas = (work + 255) / 256;
  const int grid = (int)(ctas < (int64_t)num_sms() * 8 ? ctas : (int64_t)num_sms() * 8);
  const bool narrow = work + (int64_t)grid * 256 < (1LL << 31);
#define MIMIC_ROPE(TT, II)                                                                                                   \
  rope_qk_kernel<TT, II><<<grid, 256, 0, st>>>((const TT*)q, q_sb, q_st, q_sh, (const TT*)k, k_sb, k_st, k_sh, (const TT*)cosp, \
                                               (const TT*)sinp, c_sb, c_st, (TT*)q_out, (TT*)k_out, B, T, H, Hkv, d, inverse)
  if (dtype == MIMIC_BF16) {
    if (narrow) MIMIC_ROPE(__nv_bfloat16, uint32_t);
    else MIMIC_ROPE(__nv_bfloat16, int64_t);
  } else if (dtype == MIMIC_F32) {
    if (narrow) MIMIC_ROPE(float, uint32_t);
    else MIMIC_ROPE(float, int64_t);
  }
#undef MIMIC_ROPE
  else {
    set_error("rope: unsupported dtype %d", dtype);
    return MIMIC_ERR_UNSUPPORTED;
  }
  MIMIC_LAUNCH_OK("rope_qk_kernel");
  return MIMIC_OK;
}

extern "C" size_t mimic_adamw_workspace(void) {
  return (size_t)(2 + 4 * num_sms()) * sizeof(float);   // sum, arrival counter, one partial per CTA of the norm kernel
}

extern "C" int mimic_adamw_step(float* param, float* grad, float* exp_avg, float* exp_avg_sq, const float* weight_decay,
                                int64_t n, const float* hyper, float beta1, float beta2, float eps, float grad_mul,
                                float max_norm, float* sumsq, void* stream) {
  return mimic_adamw_step_groups(param, grad, exp_avg, exp_avg_sq, weight_decay, nullptr, n, hyper, beta1, beta2, eps,
                                 grad_mul, max_norm, sumsq, stream);
}

extern "C" int mimic_adamw_step_groups(float* param, float* grad, float* exp_avg, float* exp_avg_sq,
                                       const float* weight_decay, const float* lr_mul, int64_t n, const float* hyper,
                                       float beta1, float beta2, float eps, float grad_mul, float max_norm, float* sumsq,
                                       void* stream) {
  MIMIC_CHECK_ARG(param && grad && exp_avg && exp_avg_sq && weight_decay && hyper && sumsq, "NULL argument");
  MIMIC_CHECK_ARG(n > 0, "bad n");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int grid = (int)((n + 255) / 256 < (int64_t)num_sms() * 4 ? (n + 255) / 256 : (int64_t)num_sms() * 4);
  grad_sumsq_kernel<<<grid, 256, 0, st>>>(grad, n, grad_mul, sumsq);
  MIMIC_LAUNCH_OK("grad_sumsq_kernel");
  adamw_kernel<<<grid, 256, 0, st>>>(param, grad, exp_avg, exp_avg_sq, weight_decay, n, hyper, beta1, beta2, eps, max_norm, sumsq,
                                     lr_mul);
  MIMIC_LAUNCH_OK("adamw_kernel");
  return MIMIC_OK;
}
