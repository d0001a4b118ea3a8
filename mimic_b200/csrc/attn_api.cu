// C-ABI entry points of the shifted attention: argument validation and kernel-family choice.
#include "attn_common.cuh"

namespace mimic {

static int fill(const mimic_attn_desc_t* d, AttnParams& p) {
  MIMIC_CHECK_ARG(d != nullptr, "desc is NULL");
  MIMIC_CHECK_ARG(d->B > 0 && d->H > 0 && d->Hkv > 0 && d->T > 0 && d->S > 0 && d->d > 0,
                  "non-positive dimension (B=%d H=%d Hkv=%d T=%d S=%d d=%d)", d->B, d->H, d->Hkv, d->T, d->S, d->d);
  MIMIC_CHECK_ARG(d->H % d->Hkv == 0, "H (%d) must be a multiple of Hkv (%d)", d->H, d->Hkv);
  MIMIC_CHECK_ARG(d->dtype == MIMIC_BF16 || d->dtype == MIMIC_F32, "unknown dtype %d", d->dtype);
  MIMIC_CHECK_ARG(d->d % 8 == 0, "head_dim %d must be a multiple of 8", d->d);
  const int64_t strides[] = {d->q_sb, d->q_sh, d->q_st, d->k_sb, d->k_sh, d->k_ss,
                             d->v_sb, d->v_sh, d->v_ss, d->o_sb, d->o_st, d->o_sh};
  for (int64_t s : strides) MIMIC_CHECK_ARG(s % 8 == 0, "every stride must be a multiple of 8 elements (got %lld)", (long long)s);
  p.B = d->B; p.H = d->H; p.Hkv = d->Hkv; p.T = d->T; p.S = d->S; p.d = d->d;
  p.q_sb = d->q_sb; p.q_sh = d->q_sh; p.q_st = d->q_st;
  p.k_sb = d->k_sb; p.k_sh = d->k_sh; p.k_ss = d->k_ss;
  p.v_sb = d->v_sb; p.v_sh = d->v_sh; p.v_ss = d->v_ss;
  p.o_sb = d->o_sb; p.o_st = d->o_st; p.o_sh = d->o_sh;
  p.go_sb = d->o_sb; p.go_st = d->o_st; p.go_sh = d->o_sh;
  p.scale = d->scale; p.causal = d->causal; p.active = d->active; p.dtype = d->dtype;
  return MIMIC_OK;
}

static bool aligned16(const void* ptr) { return (reinterpret_cast<uintptr_t>(ptr) & 15u) == 0; }

static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

}  // namespace mimic

using namespace mimic;

extern "C" int mimic_shifted_attn_fwd(const mimic_attn_desc_t* desc, const void* q, const void* k, const void* v,
                                      const uint8_t* key_mask, const float* shift, const float* w,
                                      const float* bias, void* out, float* lse_c, float* lse_full, float* mu,
                                      void* stream) {
  AttnParams p{};
  int rc = fill(desc, p);
  if (rc) return rc;
  MIMIC_CHECK_ARG(q && k && v && out && lse_c, "q/k/v/out/lse_c must not be NULL");
  MIMIC_CHECK_ARG(aligned16(q) && aligned16(k) && aligned16(v) && aligned16(out), "q/k/v/out must be 16-byte aligned");
  if (p.active) MIMIC_CHECK_ARG(shift && w && bias && lse_full && mu, "active shift needs shift/w/bias/lse_full/mu");
  p.q = q; p.k = k; p.v = v; p.key_mask = key_mask; p.shift = shift; p.w = w; p.bias = bias;
  p.out = out; p.lse_c = lse_c; p.lse_full = lse_full; p.mu = mu;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const bool tc_ok = tc::supports(p, false);
  if (desc->impl == MIMIC_IMPL_TCGEN05 && !tc_ok) {
    set_error("tcgen05 forward requested but shape unsupported (needs bf16, d == 128, 16B-aligned strides)");
    return MIMIC_ERR_UNSUPPORTED;
  }
  if (tc_ok && desc->impl != MIMIC_IMPL_SIMT) {
    set_attn_impl("tcgen05");
    return tc::fwd(p, st);
  }
  set_attn_impl("simt");
  return simt::fwd(p, st);
}

extern "C" size_t mimic_shifted_attn_bwd_workspace(const mimic_attn_desc_t* d) {
  if (!d) return 0;
  const size_t rows = (size_t)d->B * d->H * d->T;
  return align_up(rows * d->d * sizeof(float), 256) + 2 * align_up(rows * sizeof(float), 256);
}

extern "C" int mimic_shifted_attn_bwd(const mimic_attn_desc_t* desc, const void* q, const void* k, const void* v,
                                      const uint8_t* key_mask, const float* shift, const float* w,
                                      const float* bias, const void* out, const void* d_out, int64_t go_sb,
                                      int64_t go_st, int64_t go_sh, const float* lse_c, const float* lse_full,
                                      const float* mu, void* dq, void* dk, void* dv, float* dshift, float* dw,
                                      float* dbias, void* workspace, size_t workspace_bytes, void* stream) {
  return mimic_shifted_attn_bwd_lse(desc, q, k, v, key_mask, shift, w, bias, out, d_out, go_sb, go_st, go_sh, lse_c,
                                    lse_full, mu, nullptr, dq, dk, dv, dshift, dw, dbias, workspace, workspace_bytes,
                                    stream);
}

extern "C" int mimic_shifted_attn_bwd_lse(const mimic_attn_desc_t* desc, const void* q, const void* k, const void* v,
                                          const uint8_t* key_mask, const float* shift, const float* w,
                                          const float* bias, const void* out, const void* d_out, int64_t go_sb,
                                          int64_t go_st, int64_t go_sh, const float* lse_c, const float* lse_full,
                                          const float* mu, const float* d_lse_full, void* dq, void* dk, void* dv,
                                          float* dshift, float* dw, float* dbias, void* workspace,
                                          size_t workspace_bytes, void* stream) {
  AttnParams p{};
  int rc = fill(desc, p);
  if (rc) return rc;
  MIMIC_CHECK_ARG(q && k && v && out && d_out && lse_c && dq && dk && dv, "NULL tensor argument");
  MIMIC_CHECK_ARG(aligned16(q) && aligned16(k) && aligned16(v) && aligned16(out) && aligned16(d_out) &&
                      aligned16(dq) && aligned16(dk) && aligned16(dv),
                  "tensors must be 16-byte aligned");
  if (p.active)
    MIMIC_CHECK_ARG(shift && w && bias && lse_full && mu && dshift && dw && dbias,
                    "active shift needs shift/w/bias/lse_full/mu and dshift/dw/dbias");
  MIMIC_CHECK_ARG(p.active || !d_lse_full, "d_lse_full needs an active handle (lse_full is only produced then)");
  if (go_sb || go_st || go_sh) {
    MIMIC_CHECK_ARG(go_sb % 8 == 0 && go_st % 8 == 0 && go_sh % 8 == 0, "d_out strides must be multiples of 8");
    p.go_sb = go_sb; p.go_st = go_st; p.go_sh = go_sh;
  }
  const size_t need = mimic_shifted_attn_bwd_workspace(desc);
  MIMIC_CHECK_ARG(workspace && workspace_bytes >= need, "workspace too small: %zu < %zu", workspace_bytes, need);
  MIMIC_CHECK_ARG((reinterpret_cast<uintptr_t>(workspace) & 255u) == 0, "workspace must be 256-byte aligned");
  const size_t rows = (size_t)p.B * p.H * p.T;
  char* ws = static_cast<char*>(workspace);
  p.dq_acc = reinterpret_cast<float*>(ws);
  ws += align_up(rows * p.d * sizeof(float), 256);
  p.delta_ws = reinterpret_cast<float*>(ws);
  ws += align_up(rows * sizeof(float), 256);
  p.dz_ws = reinterpret_cast<float*>(ws);
  p.q = q; p.k = k; p.v = v; p.key_mask = key_mask; p.shift = shift; p.w = w; p.bias = bias;
  p.out = const_cast<void*>(out); p.d_out = d_out;
  p.d_lse_full = d_lse_full;
  p.lse_c = const_cast<float*>(lse_c); p.lse_full = const_cast<float*>(lse_full); p.mu = const_cast<float*>(mu);
  p.dq = dq; p.dk = dk; p.dv = dv; p.dshift = dshift; p.dw = dw; p.dbias = dbias;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const bool tc_ok = tc::supports(p, true);
  if (desc->impl == MIMIC_IMPL_TCGEN05 && !tc_ok) {
    set_error("tcgen05 backward requested but shape unsupported");
    return MIMIC_ERR_UNSUPPORTED;
  }
  if (tc_ok && desc->impl != MIMIC_IMPL_SIMT) {
    set_attn_impl("tcgen05");
    return tc::bwd(p, st);
  }
  set_attn_impl("simt");
  return simt::bwd(p, st);
}
