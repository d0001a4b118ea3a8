// Host-side TMA tensor-map construction shared by the tcgen05 attention kernels.
#pragma once
#include <cuda.h>

#include "attn_common.cuh"

namespace mimic {
namespace tc {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static inline EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

// [B, Hn, R, 128] bf16 tensor with element strides (sb, sh, sr, 1) -> 4-D map, box = 64 cols x box_rows rows
static inline int make_map(CUtensorMap* m, const void* base, int B, int Hn, int R, int64_t sb, int64_t sh, int64_t sr,
                           int box_rows) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled entry point not available");
    return MIMIC_ERR_CUDA;
  }
  cuuint64_t dims[4] = {(cuuint64_t)128, (cuuint64_t)R, (cuuint64_t)Hn, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)sr * 2, (cuuint64_t)sh * 2, (cuuint64_t)sb * 2};
  // a size-1 dimension may carry any stride; TMA wants non-zero multiples of 16 bytes
  for (int i = 0; i < 3; ++i)
    if (strides[i] == 0) strides[i] = (cuuint64_t)128 * 2;
  cuuint32_t box[4] = {64, (cuuint32_t)box_rows, 1, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (R=%d Hn=%d B=%d strides %lld %lld %lld)", (int)r, R, Hn, B,
              (long long)sr, (long long)sh, (long long)sb);
    return MIMIC_ERR_CUDA;
  }
  return MIMIC_OK;
}

// fp32 [N, R, 128] contiguous accumulator (dq_acc) -> 3-D map, box = 32 cols x box_rows rows, no swizzle: the target of the
// per-warp TMA reduce-adds of the backward (rows past R are clipped, so a ragged last tile cannot touch the next head)
static inline int make_map_f32_rows(CUtensorMap* m, const void* base, int64_t N, int R, int box_rows) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled entry point not available");
    return MIMIC_ERR_CUDA;
  }
  cuuint64_t dims[3] = {(cuuint64_t)128, (cuuint64_t)R, (cuuint64_t)N};
  cuuint64_t strides[2] = {(cuuint64_t)128 * 4, (cuuint64_t)R * 128 * 4};
  cuuint32_t box[3] = {32, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (fp32 accumulator) failed with CUresult %d (N=%lld R=%d)", (int)r, (long long)N, R);
    return MIMIC_ERR_CUDA;
  }
  return MIMIC_OK;
}

}  // namespace tc
}  // namespace mimic
