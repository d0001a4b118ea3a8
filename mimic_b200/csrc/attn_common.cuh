// Parameter block shared by the shifted-attention kernel families.
#pragma once
#include "common.cuh"

namespace mimic {

struct AttnParams {
  int B, H, Hkv, T, S, d;
  int64_t q_sb, q_sh, q_st;
  int64_t k_sb, k_sh, k_ss;
  int64_t v_sb, v_sh, v_ss;
  int64_t o_sb, o_st, o_sh;
  int64_t go_sb, go_st, go_sh;
  float scale;
  int causal, active, dtype;
  const void *q, *k, *v;
  const uint8_t* key_mask;
  const float *shift, *w, *bias;
  void* out;
  float *lse_c, *lse_full, *mu;
  // backward only
  const void* d_out;
  const float* d_lse_full = nullptr;   // optional gradient arriving at lse_full from outside the op (head-averaged mu variant)
  void *dq, *dk, *dv;
  float *dshift, *dw, *dbias;
  float *dq_acc, *delta_ws, *dz_ws;
  int debug_flags = 0;
  unsigned long long* trace = nullptr;   // optional in-kernel timeline (mimic_debug_set_trace), CTA (0,0,0) only
  float dq_seed_mul = 1.f, dq_out_mul = 1.f;   // dq_acc seed = dz*w*seed_mul; dq = dq_acc*out_mul
  int dq_direct = 0;                           // tcgen05 backward with ONE key tile (S <= 128): dq leaves the main kernel as bf16
};

unsigned long long* debug_trace_buffer();
size_t debug_trace_capacity();

namespace simt {
int fwd(const AttnParams& p, cudaStream_t st);
int bwd(const AttnParams& p, cudaStream_t st);
int bwd_prep_bf16_d128(const AttnParams& p, cudaStream_t st);   // Delta', dz, dq seed, parameter gradients
int convert_dq_bf16(const AttnParams& p, cudaStream_t st);      // fp32 dq accumulator -> bf16 dq
}  // namespace simt

namespace tc {
// tcgen05 / TMEM / TMA kernels: bf16, d == 128
bool supports(const AttnParams& p, bool backward);
int fwd(const AttnParams& p, cudaStream_t st);
int bwd(const AttnParams& p, cudaStream_t st);
void force_fwd_kernel(int k);   // test / A-B hook behind mimic_debug_force_fwd_kernel
}  // namespace tc

}  // namespace mimic
