// Micro-benchmark: cycles per tcgen05.mma (cta_group::1, kind::f16, bf16) for the operand shapes the attention
// kernels use, with and without concurrent LSU shared-memory traffic.  Build: scripts/build_mma_bench.sh; run under gpurun.
// Answers: is an SS MMA with N=64 bound by the shared-memory operand read (A 4 KB + B 2 KB per 32-cycle slot)?
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#include "../mimic_b200/csrc/tc_ptx.cuh"

using namespace mimic::ptx;

struct Cfg {
  int ts;        // 0: A from smem, 1: A from TMEM
  int N;         // MMA N
  int a_mn, b_mn;
  int nsteps;    // K steps per "tile" (different smem offsets), cycled
  int reps;      // tiles
  int hammer;    // 0: idle warps, 1: other warps stream 16-byte smem stores+loads
};

template <int TS, int N, int AMN, int BMN, int NSTEPS>
__global__ void __launch_bounds__(288, 1) mma_bench_kernel(Cfg c, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  __shared__ volatile int stop;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 196608 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_barrier_init();
    stop = 0;
  }
  if (warp == 0) tmem_alloc(&tmem_slot, 512);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (warp == 0) {
    constexpr uint32_t idesc = idesc_bf16(128, N, AMN, BMN);
    constexpr uint32_t HI = desc_hi_sw128(1024);
    const uint32_t a0 = desc_lo(smem_u32(smem), AMN ? 16384 : 16);
    const uint32_t b0 = desc_lo(smem_u32(smem + 65536), BMN ? 16384 : 16);
    long long t0 = 0, t1 = 0;
    if (elect_one()) {
      t0 = clock64();
      for (int r = 0; r < c.reps; ++r) {
        const uint32_t d = tmem + (r & 1) * 256;
#pragma unroll
        for (int k = 0; k < NSTEPS; ++k) {
          const uint32_t ao = AMN ? k * (2048 >> 4) : (((k >> 2) * 16384 + (k & 3) * 32) >> 4);
          const uint32_t bo = BMN ? k * (2048 >> 4) : (((k >> 2) * 16384 + (k & 3) * 32) >> 4);
          if (TS) {
            if (k == 0) mma_ts_c<false>(d, tmem + 128, desc64(HI, b0 + bo), idesc);
            else mma_ts_c<true>(d, tmem + 128 + 8 * k, desc64(HI, b0 + bo), idesc);
          } else {
            if (k == 0) mma_ss_c<false>(d, desc64(HI, a0 + ao), desc64(HI, b0 + bo), idesc);
            else mma_ss_c<true>(d, desc64(HI, a0 + ao), desc64(HI, b0 + bo), idesc);
          }
        }
      }
      tc_commit(&bar);
    }
    __syncwarp();
    mbar_wait(&bar, 0, 1);
    t1 = clock64();
    t0 = __shfl_sync(0xffffffffu, t0, 0);   // elected lane is lane 0 in practice; fall back to max below
    if (lane == 0) {
      stop = 1;
      if (blockIdx.x == 0) out[0] = t1 - t0;
    }
  } else if (c.hammer) {
    // warps 1..8: 16-byte stores + loads over a 32 KB window (what softmax / drain warps do)
    uint4* w = reinterpret_cast<uint4*>(smem + 131072) + (warp - 1) * 256;
    uint4 v = make_uint4(lane, warp, 0, 0);
    long long n = 0;
    while (!stop) {
#pragma unroll
      for (int j = 0; j < 8; ++j) w[j * 32 + lane] = v;
      if (c.hammer > 1) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          uint4 u = w[j * 32 + (lane ^ 1)];
          v.z += u.x;
        }
      }
      ++n;
    }
    if (v.z == 0xdeadbeef) out[1] = n;
    if (blockIdx.x == 0 && lane == 0 && warp == 1) out[2] = n * 8 * 512 * (c.hammer > 1 ? 2 : 1);   // bytes moved by ONE warp
  }
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

template <int TS, int N, int AMN, int BMN, int NSTEPS>
static void launch_t(const Cfg& c, long long* out) {
  cudaFuncSetAttribute(mma_bench_kernel<TS, N, AMN, BMN, NSTEPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 196608 + 1024);
  mma_bench_kernel<TS, N, AMN, BMN, NSTEPS><<<148, 288, 196608 + 1024>>>(c, out);
}
static void launch(const Cfg& c, long long* out) {
#define CASE(TS_, N_, A_, B_, K_) \
  if (c.ts == TS_ && c.N == N_ && c.a_mn == A_ && c.b_mn == B_ && c.nsteps == K_) return launch_t<TS_, N_, A_, B_, K_>(c, out);
  CASE(0, 64, 0, 0, 8) CASE(0, 128, 0, 0, 8) CASE(0, 256, 0, 0, 4) CASE(0, 64, 1, 1, 8) CASE(0, 128, 1, 1, 8)
  CASE(1, 64, 0, 1, 4) CASE(1, 128, 0, 1, 4) CASE(1, 256, 0, 1, 4) CASE(1, 128, 0, 0, 4)
  printf("no instantiation for this case\n");
  exit(1);
}

int main() {
  long long* out;
  cudaMallocManaged(&out, 64);
  struct Named { const char* name; Cfg c; };
  const Named cases[] = {
      {"SS  N=64  A K-major  B K-major   (S^T, dP^T bwd; S fwd64)", {0, 64, 0, 0, 8, 64, 0}},
      {"SS  N=128 A K-major  B K-major   (S fwd128)", {0, 128, 0, 0, 8, 64, 0}},
      {"SS  N=256 A K-major  B K-major", {0, 256, 0, 0, 4, 64, 0}},
      {"SS  N=64  A MN-major B MN-major  (dQ^T bwd)", {0, 64, 1, 1, 8, 64, 0}},
      {"SS  N=128 A MN-major B MN-major", {0, 128, 1, 1, 8, 64, 0}},
      {"TS  N=64  B MN-major", {1, 64, 0, 1, 4, 64, 0}},
      {"TS  N=128 B MN-major             (dV, dK bwd; PV fwd)", {1, 128, 0, 1, 4, 64, 0}},
      {"TS  N=256 B MN-major", {1, 256, 0, 1, 4, 64, 0}},
      {"TS  N=128 B K-major", {1, 128, 0, 0, 4, 64, 0}},
      {"SS  N=64  K/K + 8 warps storing smem", {0, 64, 0, 0, 8, 64, 1}},
      {"SS  N=64  K/K + 8 warps storing+loading smem", {0, 64, 0, 0, 8, 64, 2}},
      {"SS  N=128 K/K + 8 warps storing+loading smem", {0, 128, 0, 0, 8, 64, 2}},
      {"TS  N=128 MN  + 8 warps storing+loading smem", {1, 128, 0, 1, 4, 64, 2}},
  };
  for (const auto& nc : cases) {
    for (int it = 0; it < 2; ++it) {
      out[0] = out[1] = out[2] = 0;
      launch(nc.c, out);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) {
        printf("%s: CUDA error %s\n", nc.name, cudaGetErrorString(e));
        return 1;
      }
    }
    const double n_mma = (double)nc.c.reps * nc.c.nsteps;
    const double cyc = (double)out[0] / n_mma;
    const double floor_c = 128.0 * nc.c.N / 256.0;
    const double smem_bytes = (nc.c.ts ? 0 : 4096) + nc.c.N * 32.0;
    printf("%-62s %7.1f cyc/mma  (floor %5.1f)  operand smem %5.0f B -> %6.1f B/cyc", nc.name, cyc, floor_c, smem_bytes, smem_bytes / cyc);
    if (nc.c.hammer) printf("  | LSU traffic %.1f B/cyc", 8.0 * out[2] / (double)out[0]);
    printf("\n");
  }
  return 0;
}
