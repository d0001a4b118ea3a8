/*
 * mimic_b200.h — C ABI of the B200-native MimIC hot path.
 *
 * One shared library (libmimic_b200.so), plain C symbols, raw device pointers and sizes, no
 * torch types.  Every entry point enqueues work on the caller's CUDA stream and returns an int
 * status (0 = ok); `mimic_last_error()` gives the message of the last failure on this thread.
 * Nothing here allocates device memory: scratch comes from a caller-provided workspace whose
 * size the matching `*_workspace()` query returns.
 *
 * The reference (mercurystraw/MimIC) is pure Python and has no FFI of its own; each entry point
 * below names the reference call sites (file:line, relative to the reference tree) it replaces.
 * The binding a maintainer of the reference would add is shown in INTEGRATION.md.
 */
#ifndef MIMIC_B200_H_
#define MIMIC_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MIMIC_ABI_VERSION 2
#define MIMIC_API __attribute__((visibility("default")))

enum { MIMIC_OK = 0, MIMIC_ERR_ARG = 1, MIMIC_ERR_CUDA = 2, MIMIC_ERR_UNSUPPORTED = 3 };

/* storage type of activations; all statistics / parameter gradients are fp32 */
enum { MIMIC_BF16 = 0, MIMIC_F32 = 1 };

/* which implementation `mimic_shifted_attn_*` picks for bf16, d == 128 (both are CUDA):
 * AUTO = tcgen05/TMA kernels when the shape allows, else the SIMT kernels. */
enum { MIMIC_IMPL_AUTO = 0, MIMIC_IMPL_SIMT = 1, MIMIC_IMPL_TCGEN05 = 2 };

MIMIC_API int mimic_abi_version(void);
MIMIC_API const char* mimic_last_error(void);
/* name of the kernel family the last mimic_shifted_attn_* call on this thread launched */
MIMIC_API const char* mimic_last_attn_impl(void);
/* Debug: device buffer of n_entries uint64 that CTA (0,0,0) of the tcgen05 backward kernel fills with
 * (tag << 48 | clock) records — a per-role timeline for tuning.  NULL disables (default). */
MIMIC_API void mimic_debug_set_trace(void* device_u64_buffer, size_t n_entries);
/* Debug / tests: force the forward kernel family of the tcgen05 path for later calls (0 = automatic choice, -1 = the
 * MIMIC_FWD_KERNEL environment variable; 3 = 64-key tiles, 6 = persistent ping-pong, see csrc/attn_tc.cu). */
MIMIC_API void mimic_debug_force_fwd_kernel(int kernel);

/* ------------------------------------------------------------------------------------------
 * Shifted self-attention.
 *
 * Replaces, per decoder layer, F.scaled_dot_product_attention (src/shift_encoder.py:352-359,
 * 449-456, 518-532) followed by AttnApproximator.do_shift (src/shift_encoder.py:682-732):
 *
 *   SA_t      = softmax_{s visible from t}(scale * q_t.k_s) . v_s
 *   lse_full  = log sum_{ALL s in [0,S)} exp(scale * q_t.k_s)          (unmasked, lines 697-703)
 *   z1_t      = q_t . w[h] + bias[h]                                    (MultiheadLinear, 547-556)
 *   mu_t      = exp(z1_t - logaddexp(z1_t, lse_full))                   (line 712)
 *   out_t     = SA_t + mu_t * shift[h]                                  (line 724)
 *
 * visible(b,t,s) = (!causal || s <= t + (S - T)) && (key_mask == NULL || key_mask[b*S+s] != 0).
 * A query row with no visible key gets SA = 0.  `active == 0` is the inactive-handle path (line
 * 685/732): out = SA, and lse_full / mu / shift / w / bias may be NULL.
 *
 * Layouts (element strides; the innermost, head_dim, stride is 1 for every tensor):
 *   q   [B,H,T,d]    strides q_sb,q_sh,q_st      k,v [B,Hkv,S,d]  strides k_*, v_*   (H % Hkv == 0)
 *   out [B,T,H,d]    strides o_sb,o_st,o_sh      (the layout do_shift returns)
 *   lse_c, lse_full, mu : fp32 [B,H,T] contiguous.  shift, w: fp32 [H,d]; bias: fp32 [H].
 * ------------------------------------------------------------------------------------------ */
typedef struct {
  int32_t B, H, Hkv, T, S, d;
  int64_t q_sb, q_sh, q_st;
  int64_t k_sb, k_sh, k_ss;
  int64_t v_sb, v_sh, v_ss;
  int64_t o_sb, o_st, o_sh;
  float scale;        /* 1/sqrt(d) in the reference (line 699) */
  int32_t causal;
  int32_t active;
  int32_t dtype;      /* MIMIC_BF16 | MIMIC_F32 */
  int32_t impl;       /* MIMIC_IMPL_* */
} mimic_attn_desc_t;

MIMIC_API int mimic_shifted_attn_fwd(const mimic_attn_desc_t* desc,
                           const void* q, const void* k, const void* v,
                           const uint8_t* key_mask,
                           const float* shift, const float* w, const float* bias,
                           void* out, float* lse_c, float* lse_full, float* mu,
                           void* stream);

/* Backward of the above (replaces the autograd graph of SDPA + do_shift).
 *   d_out [B,T,H,d] with the strides of `out` in desc (go_* below override when non-zero).
 *   dq [B,H,T,d], dk, dv [B,Hkv,S,d]: contiguous, dtype of desc, fully overwritten.
 *   dshift, dw [H,d], dbias [H]: fp32, ACCUMULATED INTO (+=) — they point into the flat
 *   parameter-gradient buffer that the data-parallel all-reduce later sums (src/config/train.yaml:14).
 * workspace: mimic_shifted_attn_bwd_workspace(desc) bytes, 256-byte aligned. */
MIMIC_API size_t mimic_shifted_attn_bwd_workspace(const mimic_attn_desc_t* desc);
MIMIC_API int mimic_shifted_attn_bwd(const mimic_attn_desc_t* desc,
                           const void* q, const void* k, const void* v,
                           const uint8_t* key_mask,
                           const float* shift, const float* w, const float* bias,
                           const void* out, const void* d_out,
                           int64_t go_sb, int64_t go_st, int64_t go_sh,
                           const float* lse_c, const float* lse_full, const float* mu,
                           void* dq, void* dk, void* dv,
                           float* dshift, float* dw, float* dbias,
                           void* workspace, size_t workspace_bytes, void* stream);

/* Same backward with one more incoming gradient: d_lse_full fp32 [B,H,T] (or NULL) = dL/d lse_full from
 * OUTSIDE the op.  The non-MULTI_HEAD variant of do_shift (src/shift_encoder.py:689-692, 705-707) averages
 * log_Z2 over the heads before the sigmoid, so mu is formed by the host from the lse_full this op returns
 * (forward run with shift = w = bias = 0: out = SA) and its gradient re-enters here: dS += d_lse_full * P^f. */
MIMIC_API int mimic_shifted_attn_bwd_lse(const mimic_attn_desc_t* desc,
                           const void* q, const void* k, const void* v,
                           const uint8_t* key_mask,
                           const float* shift, const float* w, const float* bias,
                           const void* out, const void* d_out,
                           int64_t go_sb, int64_t go_st, int64_t go_sh,
                           const float* lse_c, const float* lse_full, const float* mu,
                           const float* d_lse_full,
                           void* dq, void* dk, void* dv,
                           float* dshift, float* dw, float* dbias,
                           void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * Norm-preserving vector shift of a module output: y = u / |u| * |h|, u = h + scale * shift.
 * Replaces the hook body of AttnFFNShift._shift_hook (src/shift_encoder.py:222-247; the LIVE baseline,
 * config/encoder/licv.yaml) and of AttnApproximator's optional FFN hook (src/shift_encoder.py:653-672), and
 * their autograd.
 *   h [N, D] rows of stride h_row_stride elements; shift fp32 [D] ({attn,ffn}_shift[layer]); scale: device
 *   pointer to ONE fp32 ({attn,ffn}_shift_scale[layer]) or NULL for 1; y [N, D] contiguous, dtype of h.
 *   Backward: dh [N, D] contiguous is overwritten; gsum fp32 [D] is ACCUMULATED INTO with the column sums of
 *   dL/du, from which the caller forms  dshift = scale * gsum  and  dscale = gsum . shift.
 *   D: multiple of 8, at most 8192.  Norms and all arithmetic in fp32.
 * ------------------------------------------------------------------------------------------ */
MIMIC_API int mimic_norm_shift_fwd(const void* h, int64_t h_row_stride, int32_t dtype,
                         const float* shift, const float* scale, void* y,
                         int32_t N, int32_t D, void* stream);
MIMIC_API int mimic_norm_shift_bwd(const void* h, int64_t h_row_stride, const void* dy, int64_t dy_row_stride,
                         int32_t dtype, const float* shift, const float* scale,
                         void* dh, float* gsum, int32_t N, int32_t D, void* stream);

/* ------------------------------------------------------------------------------------------
 * Record path: gather the selected token rows of one layer's recorded module output straight
 * into the preallocated [L, N, D] buffer.  Replaces record_hook (src/shift_encoder.py:115-120)
 * + torch.stack + masked_select in ShiftModel.get_hidden_states (src/shift_model.py:131-136).
 *   src [rows, D] (row stride src_row_stride elements), src_row[n] = source row of compact row n.
 *   dst [N, D] contiguous (the layer's slice), dst_dtype may differ from src_dtype.
 * ------------------------------------------------------------------------------------------ */
MIMIC_API int mimic_record_gather(const void* src, int64_t src_row_stride, int32_t src_dtype,
                        const int32_t* src_row, void* dst, int32_t dst_dtype,
                        int32_t N, int32_t D, void* stream);

/* ------------------------------------------------------------------------------------------
 * Row plan: the integer side of one MimIC pass in ONE launch, from the token ids [B, S] (int64):
 *   attn_mask  [B,S] u8   ids != pad_id                                   (src/shift_model.py:212, 222)
 *   label_mask [B,S] u8   ShiftModel.generate_label_mask(inputs, num_separator, keep_bos) (75-110): tokens after
 *                         the num_separator-th PAD that is not leading padding (left_padding != 0: pads before the first
 *                         non-pad token are padding, not separators), minus pads
 *   selection  select_mode 0: the label mask; 1: non-pad and != bos_id (the query rows of the alignment loss, 253-257)
 *   rows [cap] i32        flat index b*S+s of the k-th selected token, ascending (the order masked_select / boolean
 *                         indexing enumerate, 131-141), -1 behind the last one
 *   row_to_compact [B*S]  k for a selected token, else -1
 *   row_w [cap] f32       w_scale / max(N_b, 1) for the compact rows of sample b (N_b = its selected tokens), 0 behind
 *   lm_labels [B,S] i64   ids[b,s+1] where that token is not a pad, else -100: the shifted labels of the in-model LM
 *                         loss of transformers 4.45 (call site 238-250)
 *   paired [B*S] i32      pair_src[k] for the k-th selected token (k < pair_cap), else -1: pairs it with the k-th
 *                         row of another plan (boolean-mask indexing on both sides, 185-186)
 *   count [1], row_counts [B], sep_counts [B] i32: selected tokens in all / per sample; separators per sample
 *                         (sep_counts[b] < num_separator is the reference's ValueError, 97-100: host-side check)
 * Every output pointer may be NULL.  Replaces the comparisons / cumsums / nonzero / .item() loop of the reference
 * and their ~40 eager launches per step.
 * ------------------------------------------------------------------------------------------ */
MIMIC_API int mimic_row_plan(const int64_t* ids, int32_t B, int32_t S, int64_t pad_id, int64_t bos_id,
                   int32_t left_padding, int32_t num_separator, int32_t select_mode, int32_t keep_bos, float w_scale,
                   uint8_t* attn_mask, uint8_t* label_mask, int32_t* rows, int32_t cap, int32_t* row_to_compact,
                   float* row_w, int64_t* lm_labels, const int32_t* pair_src, int32_t pair_cap, int32_t* paired,
                   int32_t* count, int32_t* row_counts, int32_t* sep_counts, void* stream);

/* ------------------------------------------------------------------------------------------
 * Layer-wise alignment loss, Strategy.LAYER_WISE_MSE, forward + backward in one pass.
 * Replaces ShiftModel.calculate_layer_wise_loss (src/shift_model.py:145-179) and its autograd:
 *   loss = mean_i mean_{l,t,d} (H_i - H'_i)^2 = sum_{l,n,d} row_w[n] * (H[l][row[n]] - H'[l][n])^2
 * with row_w[n] = 1 / (n_samples * L * Nq_{sample(n)} * D) supplied by the caller.
 *   h_ptrs   device array of L pointers; layer l's query-pass module output [rows, D], h_row_stride
 *   hp       H' compact [L, N, D] (written by mimic_record_gather), hp_dtype
 *   row_to_compact [rows]: compact index n of row r, or -1 when the row is not selected
 *   dh_ptrs  device array of L pointers to [rows, D] contiguous gradients in h dtype, fully
 *            written (zeros on unselected rows): dH = grad_scale * 2 * row_w * (H - H').
 *            NULL = forward only.
 *   loss     fp32 [1], overwritten.   workspace: mimic_mse_workspace() bytes, ZERO on first use (the kernel
 *            keeps an arrival counter there and leaves it zero): one launch, deterministic final sum.
 * ------------------------------------------------------------------------------------------ */
MIMIC_API size_t mimic_mse_workspace(int32_t L, int32_t rows, int32_t D);
MIMIC_API int mimic_mse_fwd_bwd(const void* const* h_ptrs, int64_t h_row_stride, int32_t h_dtype,
                      const void* hp, int32_t hp_dtype,
                      const int32_t* row_to_compact, const float* row_w,
                      void* const* dh_ptrs, float grad_scale,
                      int32_t L, int32_t rows, int32_t N, int32_t D,
                      float* loss, void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * LM-loss cross-entropy, forward + backward in one launch, gradient optionally in place.
 * Replaces the HF in-model loss invoked at src/shift_model.py:238-250 (log_softmax + nll_loss on
 * fp32-upcast logits, mean over kept positions).
 *   logits [R, V] (row stride logits_row_stride), labels int64 [R]; label < 0 = ignored row.
 *   dlogits [R, V] contiguous-or-same-stride as logits (may alias logits); NULL = forward only:
 *            dlogits = grad_scale * (softmax(logits) - onehot(label)) / n_kept, 0 on ignored rows.
 *   loss fp32 [1] overwritten with the mean over kept rows (0 when none); n_kept int32 [1].
 *   workspace: mimic_ce_workspace(R) bytes, ZERO on first use (arrival counter + per-row losses).
 * ------------------------------------------------------------------------------------------ */
MIMIC_API size_t mimic_ce_workspace(int32_t R);
MIMIC_API int mimic_ce_fwd_bwd(const void* logits, int64_t logits_row_stride, int32_t dtype,
                     const int64_t* labels, void* dlogits, float grad_scale,
                     int32_t R, int32_t V, float* loss, int32_t* n_kept,
                     void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * Layer-wise alignment loss, Strategy.LAYER_WISE_COS_SIM (src/shift_model.py:152-160), forward + backward in
 * one pass.  Same arguments and workspace as mimic_mse_fwd_bwd (mimic_mse_workspace), D <= 8192:
 *   loss = mean_i mean_l (1 - mean_t cos(H_i[l,t], H'_i[l,t])) = sum_{l,n} row_w[n] * (1 - cos_ln)
 * with row_w[n] = 1 / (n_samples * L * Nq_{sample(n)}) supplied by the caller; cos as F.cosine_similarity
 * (each norm clamped at 1e-8).  dH = -grad_scale * row_w * (h'/(|h||h'|) - cos * h/|h|^2), zeros on unselected rows.
 * ------------------------------------------------------------------------------------------ */
MIMIC_API int mimic_cos_fwd_bwd(const void* const* h_ptrs, int64_t h_row_stride, int32_t h_dtype,
                      const void* hp, int32_t hp_dtype,
                      const int32_t* row_to_compact, const float* row_w,
                      void* const* dh_ptrs, float grad_scale,
                      int32_t L, int32_t rows, int32_t N, int32_t D,
                      float* loss, void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * Logits KL divergence, Strategy.LOGITS_KL_DIV, forward + backward in one launch.  Replaces
 * ShiftModel.calculate_logits_kl_loss (src/shift_model.py:181-191): F.kl_div(log_softmax(shift_logits[mask_q]),
 * softmax(prefix_logits[mask_p]), reduction="batchmean") and its autograd, without the two masked copies.
 *   z [R, V] query-pass logits (row stride z_row_stride); y [*, V] context-pass logits (y_row_stride);
 *   y_row int32 [R]: row of y paired with row i of z, or < 0 when row i is not selected.
 *   dz [R, V] with the strides of z (may alias z); NULL = forward only:
 *            dz = grad_scale / n_rows * (softmax(z) - softmax(y)), 0 on unselected rows.
 *   loss fp32 [1] = sum of row KLs / n_rows (0 when none); n_rows int32 [1].
 *   workspace: mimic_ce_workspace(R) bytes, ZERO on first use.
 * ------------------------------------------------------------------------------------------ */
MIMIC_API int mimic_kl_fwd_bwd(const void* z, int64_t z_row_stride, const void* y, int64_t y_row_stride, int32_t dtype,
                     const int32_t* y_row, void* dz, float grad_scale, int32_t R, int32_t V,
                     float* loss, int32_t* n_rows, void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * Rotary position embedding of q and k in ONE launch: the RoPE step of the replaced attention forwards
 * (src/shift_encoder.py:307-312, 417-422, 495-501; HF apply_rotary_pos_emb: x*cos + rotate_half(x)*sin).
 *   q [B,T,H,d], k [B,T,Hkv,d] read through element strides (*_sb batch, *_st token, *_sh head; head_dim stride 1):
 *   the q/k projection outputs are consumed where they lie.  cos, sin [B,T,d] (c_sb, c_st; c_sb = 0 broadcasts).
 *   q_out [B,T,H,d], k_out [B,T,Hkv,d] contiguous (may alias the inputs when those are contiguous).
 *   inverse != 0 applies the transpose (the backward: d/dx of the rotation).   d % 16 == 0.
 * ------------------------------------------------------------------------------------------ */
MIMIC_API int mimic_rope_qk(const void* q, int64_t q_sb, int64_t q_st, int64_t q_sh,
                  const void* k, int64_t k_sb, int64_t k_st, int64_t k_sh,
                  const void* cos, const void* sin, int64_t c_sb, int64_t c_st,
                  void* q_out, void* k_out, int32_t dtype,
                  int32_t B, int32_t T, int32_t H, int32_t Hkv, int32_t d, int32_t inverse, void* stream);

/* ------------------------------------------------------------------------------------------
 * On-device optimizer step for the flat fp32 shift-parameter buffer: global-norm clip
 * (train.yaml grad_clip_val) + AdamW with per-element weight decay (the reference's two groups,
 * src/shift_model.py:306-372: names containing "bias" get 0).  Replaces DeepSpeedCPUAdam's host
 * round trip.  What changes per step arrives in device memory so a captured CUDA graph can replay it:
 *   hyper: fp32 [3] = { lr, 1 - beta1^step, sqrt(1 - beta2^step) }.
 *   sumsq: workspace of mimic_adamw_workspace() bytes, zero on first use (the kernels leave it reusable); its first
 *          float holds the squared gradient norm after the call.  The norm is summed in a FIXED order (per-CTA
 *          partials added by the last CTA): ranks holding bit-identical gradients clip by the same factor.
 *          grad is scaled in place by grad_mul (1/world, 1/accum) first.
 * ------------------------------------------------------------------------------------------ */
MIMIC_API size_t mimic_adamw_workspace(void);
MIMIC_API int mimic_adamw_step(float* param, float* grad, float* exp_avg, float* exp_avg_sq,
                     const float* weight_decay, int64_t n, const float* hyper, float beta1, float beta2,
                     float eps, float grad_mul, float max_norm, float* sumsq, void* stream);

/* Same step with a per-element learning-rate multiplier lr_mul fp32 [n] (or NULL = all ones): the reference's optional
 * second learning rate, cfg.peft.scale_lr, for parameters whose name contains "log_Z1" or "scale"
 * (src/shift_model.py:341-355; "typically used for LIVE").  lr_mul[i] = scale_lr / lr for those, 1 elsewhere — the
 * cosine-with-warm-up factor multiplies both groups alike (LambdaLR). */
MIMIC_API int mimic_adamw_step_groups(float* param, float* grad, float* exp_avg, float* exp_avg_sq,
                     const float* weight_decay, const float* lr_mul, int64_t n, const float* hyper,
                     float beta1, float beta2, float eps, float grad_mul, float max_norm, float* sumsq,
                     void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MIMIC_B200_H_ */
