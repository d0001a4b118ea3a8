// Norm-preserving vector shift of a module output — the hook body of AttnFFNShift._shift_hook
// (src/shift_encoder.py:222-247; LIVE / licv.yaml) and of AttnApproximator's optional FFN hook (653-672):
//
//     u = h + s * v            v = {attn,ffn}_shift[layer] (fp32 [D]),  s = {attn,ffn}_shift_scale[layer] (or 1)
//     y = u / |u| * |h|        norms over the hidden dimension, one per token row
//
// The reference spends 7 elementwise / reduction launches and 5 full-size temporaries per hooked module on it (and as
// many again in autograd).  Here: one launch forward, one launch backward, each a single pass over the rows — HBM-bound,
// algorithmic bytes 2 * N * D * sizeof(T) forward (read h, write y) and 3 * N * D * sizeof(T) backward (read h, read dy,
// write dh).  One CTA per token row (grid-stride); every thread keeps its column slice of the row in registers between
// the reduction and the write, and its slice of v (and of the column sums of dL/du in the backward) across rows.
#include <algorithm>

#include "common.cuh"

namespace mimic {
namespace {

constexpr int NS_THREADS = 256;

__device__ __forceinline__ void block_sum3(float& a, float& b, float& c, float* red) {
  a = warp_sum(a);
  b = warp_sum(b);
  c = warp_sum(c);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();   // red[] from the previous row is no longer being read
  if (lane == 0) {
    red[warp] = a;
    red[8 + warp] = b;
    red[16 + warp] = c;
  }
  __syncthreads();
  a = b = c = 0.f;
#pragma unroll
  for (int w = 0; w < NS_THREADS / 32; ++w) {
    a += red[w];
    b += red[8 + w];
    c += red[16 + w];
  }
}

__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// 8 consecutive elements as loaded, unpacked to fp32 only when the row is processed: the NEXT row's loads are issued
// before the current row's reduction, so a CTA always has a row in flight (one row per CTA at a time left the backward
// at 0.28 of HBM peak)
template <typename T>
struct Raw8;
template <>
struct Raw8<__nv_bfloat16> {
  uint4 u;
  __device__ __forceinline__ void load(const __nv_bfloat16* p) { u = ldg_stream(p); }
  __device__ __forceinline__ void unpack(float (&f)[8]) const {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float2 t = __bfloat1622float2(h[i]);
      f[2 * i] = t.x;
      f[2 * i + 1] = t.y;
    }
  }
};
template <>
struct Raw8<float> {
  uint4 a, b;
  __device__ __forceinline__ void load(const float* p) {
    a = ldg_stream(p);
    b = ldg_stream(p + 4);
  }
  __device__ __forceinline__ void unpack(float (&f)[8]) const {
    f[0] = __uint_as_float(a.x); f[1] = __uint_as_float(a.y); f[2] = __uint_as_float(a.z); f[3] = __uint_as_float(a.w);
    f[4] = __uint_as_float(b.x); f[5] = __uint_as_float(b.y); f[6] = __uint_as_float(b.z); f[7] = __uint_as_float(b.w);
  }
};

// VEC 8-element vectors per thread: D <= NS_THREADS * 8 * VEC
template <typename T, int VEC>
__global__ void __launch_bounds__(NS_THREADS) norm_shift_fwd_kernel(const T* __restrict__ h, int64_t h_stride,
                                                                    const float* __restrict__ shift,
                                                                    const float* __restrict__ scale, T* __restrict__ y,
                                                                    int N, int D) {
  __shared__ float red[24];
  const int nvec = D / 8;
  const float s = scale ? *scale : 1.f;
  float sv[VEC][8];
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    const int j = threadIdx.x + NS_THREADS * i;
#pragma unroll
    for (int e = 0; e < 8; ++e) sv[i][e] = j < nvec ? s * shift[8 * j + e] : 0.f;
  }
  Raw8<T> nh[VEC];
  if (blockIdx.x < N) {
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      const int j = threadIdx.x + NS_THREADS * i;
      if (j < nvec) nh[i].load(h + (int64_t)blockIdx.x * h_stride + 8 * j);
    }
  }
  for (int row = blockIdx.x; row < N; row += gridDim.x) {
    float u[VEC][8];
    float hh = 0.f, uu = 0.f, unused = 0.f;
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      const int j = threadIdx.x + NS_THREADS * i;
      if (j < nvec) nh[i].unpack(u[i]);
    }
    if (row + (int)gridDim.x < N) {
#pragma unroll
      for (int i = 0; i < VEC; ++i) {
        const int j = threadIdx.x + NS_THREADS * i;
        if (j < nvec) nh[i].load(h + (int64_t)(row + gridDim.x) * h_stride + 8 * j);
      }
    }
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      const int j = threadIdx.x + NS_THREADS * i;
      if (j < nvec) {
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          hh += u[i][e] * u[i][e];
          u[i][e] += sv[i][e];
          uu += u[i][e] * u[i][e];
        }
      }
    }
    block_sum3(hh, uu, unused, red);
    const float ratio = sqrtf(hh) / sqrtf(uu);   // |u| == 0 divides by zero exactly as the reference does
    T* yr = y + (int64_t)row * D;
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      const int j = threadIdx.x + NS_THREADS * i;
      if (j < nvec) {
#pragma unroll
        for (int e = 0; e < 8; ++e) u[i][e] *= ratio;
        store8(yr + 8 * j, u[i]);
      }
    }
  }
}

// dh = g_u + (p / (a c)) h,   g_u = (a / c) dy - (a p / c^3) u,   a = |h|, c = |u|, p = dy . u;   gsum[d] += sum_rows g_u
template <typename T, int VEC>
__global__ void __launch_bounds__(NS_THREADS) norm_shift_bwd_kernel(const T* __restrict__ h, int64_t h_stride,
                                                                    const T* __restrict__ dy, int64_t dy_stride,
                                                                    const float* __restrict__ shift,
                                                                    const float* __restrict__ scale, T* __restrict__ dh,
                                                                    float* __restrict__ gsum, int N, int D) {
  __shared__ float red[24];
  const int nvec = D / 8;
  const float s = scale ? *scale : 1.f;
  float sv[VEC][8], gacc[VEC][8];
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    const int j = threadIdx.x + NS_THREADS * i;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      sv[i][e] = j < nvec ? s * shift[8 * j + e] : 0.f;
      gacc[i][e] = 0.f;
    }
  }
  Raw8<T> nh[VEC], ng[VEC];
  if (blockIdx.x < N) {
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      const int j = threadIdx.x + NS_THREADS * i;
      if (j < nvec) {
        nh[i].load(h + (int64_t)blockIdx.x * h_stride + 8 * j);
        ng[i].load(dy + (int64_t)blockIdx.x * dy_stride + 8 * j);
      }
    }
  }
  for (int row = blockIdx.x; row < N; row += gridDim.x) {
    float hv[VEC][8], gv[VEC][8];
    float hh = 0.f, uu = 0.f, pu = 0.f;
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      const int j = threadIdx.x + NS_THREADS * i;
      if (j < nvec) {
        nh[i].unpack(hv[i]);
        ng[i].unpack(gv[i]);
      }
    }
    if (row + (int)gridDim.x < N) {
#pragma unroll
      for (int i = 0; i < VEC; ++i) {
        const int j = threadIdx.x + NS_THREADS * i;
        if (j < nvec) {
          nh[i].load(h + (int64_t)(row + gridDim.x) * h_stride + 8 * j);
          ng[i].load(dy + (int64_t)(row + gridDim.x) * dy_stride + 8 * j);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      const int j = threadIdx.x + NS_THREADS * i;
      if (j < nvec) {
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float uv = hv[i][e] + sv[i][e];
          hh += hv[i][e] * hv[i][e];
          uu += uv * uv;
          pu += gv[i][e] * uv;
        }
      }
    }
    block_sum3(hh, uu, pu, red);
    const float a = sqrtf(hh), c = sqrtf(uu);
    const float k_dy = a / c, k_u = a * pu / (c * c * c);
    const float k_h = a > 0.f ? pu / (a * c) : 0.f;   // d|h|/dh at h == 0: the zero subgradient torch uses
    T* dr = dh + (int64_t)row * D;
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      const int j = threadIdx.x + NS_THREADS * i;
      if (j < nvec) {
        float o[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float gu = k_dy * gv[i][e] - k_u * (hv[i][e] + sv[i][e]);
          gacc[i][e] += gu;
          o[e] = gu + k_h * hv[i][e];
        }
        store8(dr + 8 * j, o);
      }
    }
  }
  if (blockIdx.x < N) {
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      const int j = threadIdx.x + NS_THREADS * i;
      if (j < nvec) {
        red_add_v4(gsum + 8 * j, gacc[i][0], gacc[i][1], gacc[i][2], gacc[i][3]);
        red_add_v4(gsum + 8 * j + 4, gacc[i][4], gacc[i][5], gacc[i][6], gacc[i][7]);
      }
    }
  }
}

int check_common(const void* h, int64_t h_stride, int32_t dtype, const float* shift, int32_t N, int32_t D) {
  MIMIC_CHECK_ARG(h && shift, "NULL tensor argument");
  MIMIC_CHECK_ARG(dtype == MIMIC_BF16 || dtype == MIMIC_F32, "dtype must be MIMIC_BF16 or MIMIC_F32");
  MIMIC_CHECK_ARG(N >= 0 && D > 0 && D % 8 == 0 && D <= NS_THREADS * 8 * 4,
                  "norm shift: hidden size %d unsupported (multiple of 8, at most %d)", D, NS_THREADS * 8 * 4);
  MIMIC_CHECK_ARG(h_stride >= D && h_stride % 8 == 0, "row stride must be a multiple of 8 and >= D");
  MIMIC_CHECK_ARG((reinterpret_cast<uintptr_t>(h) & 15u) == 0 && (reinterpret_cast<uintptr_t>(shift) & 15u) == 0,
                  "tensors must be 16-byte aligned");
  return MIMIC_OK;
}

template <typename T>
int launch_fwd(const void* h, int64_t hs, const float* shift, const float* scale, void* y, int N, int D, cudaStream_t st) {
  const int grid = std::min(N, 4 * num_sms());
  const T* hp = static_cast<const T*>(h);
  T* yp = static_cast<T*>(y);
  if (D <= NS_THREADS * 8)
    norm_shift_fwd_kernel<T, 1><<<grid, NS_THREADS, 0, st>>>(hp, hs, shift, scale, yp, N, D);
  else if (D <= NS_THREADS * 16)
    norm_shift_fwd_kernel<T, 2><<<grid, NS_THREADS, 0, st>>>(hp, hs, shift, scale, yp, N, D);
  else
    norm_shift_fwd_kernel<T, 4><<<grid, NS_THREADS, 0, st>>>(hp, hs, shift, scale, yp, N, D);
  MIMIC_LAUNCH_OK("norm_shift_fwd_kernel");
  return MIMIC_OK;
}

template <typename T>
int launch_bwd(const void* h, int64_t hs, const void* dy, int64_t gs, const float* shift, const float* scale, void* dh,
               float* gsum, int N, int D, cudaStream_t st) {
  // the column sums leave each CTA as D/4 vector reductions at the end, so the grid is capped at what is resident
  const int grid = std::min(N, 4 * num_sms());
  const T* hp = static_cast<const T*>(h);
  const T* gp = static_cast<const T*>(dy);
  T* dp = static_cast<T*>(dh);
  if (D <= NS_THREADS * 8)
    norm_shift_bwd_kernel<T, 1><<<grid, NS_THREADS, 0, st>>>(hp, hs, gp, gs, shift, scale, dp, gsum, N, D);
  else if (D <= NS_THREADS * 16)
    norm_shift_bwd_kernel<T, 2><<<grid, NS_THREADS, 0, st>>>(hp, hs, gp, gs, shift, scale, dp, gsum, N, D);
  else
    norm_shift_bwd_kernel<T, 4><<<grid, NS_THREADS, 0, st>>>(hp, hs, gp, gs, shift, scale, dp, gsum, N, D);
  MIMIC_LAUNCH_OK("norm_shift_bwd_kernel");
  return MIMIC_OK;
}

}  // namespace
}  // namespace mimic

using namespace mimic;

extern "C" int mimic_norm_shift_fwd(const void* h, int64_t h_row_stride, int32_t dtype, const float* shift,
                                    const float* scale, void* y, int32_t N, int32_t D, void* stream) {
  int rc = check_common(h, h_row_stride, dtype, shift, N, D);
  if (rc) return rc;
  MIMIC_CHECK_ARG(y && (reinterpret_cast<uintptr_t>(y) & 15u) == 0, "y must be a 16-byte aligned device pointer");
  if (N == 0) return MIMIC_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  return dtype == MIMIC_BF16 ? launch_fwd<__nv_bfloat16>(h, h_row_stride, shift, scale, y, N, D, st)
                             : launch_fwd<float>(h, h_row_stride, shift, scale, y, N, D, st);
}

extern "C" int mimic_norm_shift_bwd(const void* h, int64_t h_row_stride, const void* dy, int64_t dy_row_stride,
                                    int32_t dtype, const float* shift, const float* scale, void* dh, float* gsum,
                                    int32_t N, int32_t D, void* stream) {
  int rc = check_common(h, h_row_stride, dtype, shift, N, D);
  if (rc) return rc;
  MIMIC_CHECK_ARG(dy && dh && gsum, "NULL tensor argument");
  MIMIC_CHECK_ARG(dy_row_stride >= D && dy_row_stride % 8 == 0, "dy row stride must be a multiple of 8 and >= D");
  MIMIC_CHECK_ARG((reinterpret_cast<uintptr_t>(dy) & 15u) == 0 && (reinterpret_cast<uintptr_t>(dh) & 15u) == 0 &&
                      (reinterpret_cast<uintptr_t>(gsum) & 15u) == 0,
                  "tensors must be 16-byte aligned");
  if (N == 0) return MIMIC_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  return dtype == MIMIC_BF16
             ? launch_bwd<__nv_bfloat16>(h, h_row_stride, dy, dy_row_stride, shift, scale, dh, gsum, N, D, st)
             : launch_bwd<float>(h, h_row_stride, dy, dy_row_stride, shift, scale, dh, gsum, N, D, st);
}
