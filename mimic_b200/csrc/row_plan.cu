// Kernel P — the integer side of a MimIC pass in ONE launch: from the token ids of a batch
//   attention mask (ids != pad)                                         src/shift_model.py:212, 222
//   label mask = tokens after the n-th non-leading PAD separator, minus pads  (generate_label_mask, 75-110)
//   the compact row plan of the selected rows — ascending flat indices (what masked_select / boolean indexing
//   enumerate, 131-141, 185-186), row -> compact index, per-row loss weights 1 / (scale * N_sample)  (253-257, 145-179)
//   the shifted LM labels with -100 at positions whose NEXT token is a pad     (call site 238-250; oracle/losses.py)
//   optionally the pairing of the k-th selected row with the k-th entry of another plan (185-186)
// replacing ~40 eager launches per step (comparisons, cumsums, a stable radix sort, wheres, gathers).
//
// Integer work on a few thousand tokens: latency-bound, so ONE CTA of 1024 threads walks the batch rows in order (the
// compact index is a prefix count over the FLAT index, i.e. over rows in order).  Per row: a block min-reduction for the
// first non-pad position, then chunks of 1024 tokens with ballot + popc warp scans for the two running counts
// (separators before the position, selected rows before the position).
#include "common.cuh"

namespace mimic {

namespace {

constexpr int RP_THREADS = 1024, RP_WARPS = RP_THREADS / 32;

struct RowPlanParams {
  const int64_t* ids;
  int B, S;
  int64_t pad_id, bos_id;
  int left_padding, num_separator, select_mode, keep_bos;
  float w_scale;
  uint8_t *attn_mask, *label_mask;
  int32_t* rows;
  int cap;
  int32_t* row_to_compact;
  float* row_w;
  int64_t* lm_labels;
  const int32_t* pair_src;
  int pair_cap;
  int32_t* paired;
  int32_t *count, *row_counts, *sep_counts;
};

// exclusive prefix of `flag` over the block's 1024 threads (thread order) + block total; `tot` is shared scratch [RP_WARPS + 1]
__device__ __forceinline__ int block_excl_scan(bool flag, int* tot, int& block_total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned bal = __ballot_sync(0xffffffffu, flag);
  const int in_warp = __popc(bal & ((1u << lane) - 1u));
  if (lane == 0) tot[warp] = __popc(bal);
  __syncthreads();
  if (warp == 0) {
    const int v = tot[lane];
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int n = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += n;
    }
    tot[lane] = inc - v;
    if (lane == 31) tot[RP_WARPS] = inc;
  }
  __syncthreads();
  const int r = tot[warp] + in_warp;
  block_total = tot[RP_WARPS];
  __syncthreads();                       // scratch is reused by the next call
  return r;
}

__global__ void __launch_bounds__(RP_THREADS, 1) row_plan_kernel(RowPlanParams p) {
  __shared__ int scratch_a[RP_WARPS + 1], scratch_b[RP_WARPS + 1];
  __shared__ int s_first;
  const int tid = threadIdx.x;
  const int S = p.S;
  int base = 0;                          // selected rows in the batch rows before this one
  for (int b = 0; b < p.B; ++b) {
    const int64_t* row = p.ids + (int64_t)b * S;
    // first non-pad position (argmax of the non-pad mask: 0 when the whole row is padding, like torch.argmax)
    if (tid == 0) s_first = S;
    __syncthreads();
    int mine = S;
    for (int s = tid; s < S; s += RP_THREADS)
      if (row[s] != p.pad_id) {
        mine = s;
        break;
      }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mine = min(mine, __shfl_xor_sync(0xffffffffu, mine, o));
    if ((tid & 31) == 0 && mine < S) atomicMin(&s_first, mine);
    __syncthreads();
    const int bos_pos = s_first < S ? s_first : 0;
    int sep_before = 0, sel_before = 0;  // carries across the chunks of this row
    for (int s0 = 0; s0 < S; s0 += RP_THREADS) {
      const int s = s0 + tid;
      const bool in = s < S;
      const int64_t id = in ? row[s] : p.pad_id;
      const bool pad = in && id == p.pad_id, non_pad = in && id != p.pad_id;
      const bool sep = pad && (!p.left_padding || s > bos_pos);
      int sep_tot, sel_tot;
      const int n_sep = sep_before + block_excl_scan(sep, scratch_a, sep_tot);     // separators strictly before s
      bool label = non_pad && n_sep >= p.num_separator;
      if (p.keep_bos && p.left_padding && in && s == bos_pos) label = true;
      const bool sel = p.select_mode == 0 ? label : (non_pad && id != p.bos_id);
      const int rank = base + sel_before + block_excl_scan(sel, scratch_b, sel_tot);
      if (in) {
        const int64_t flat = (int64_t)b * S + s;
        if (p.attn_mask) p.attn_mask[flat] = non_pad ? 1 : 0;
        if (p.label_mask) p.label_mask[flat] = label ? 1 : 0;
        if (p.row_to_compact) p.row_to_compact[flat] = sel ? rank : -1;
        if (sel && p.rows && rank < p.cap) p.rows[rank] = (int32_t)flat;
        if (p.paired) p.paired[flat] = (sel && rank < p.pair_cap) ? p.pair_src[rank] : -1;
        if (p.lm_labels) {
          const bool next_ok = s + 1 < S && row[s + 1] != p.pad_id;
          p.lm_labels[flat] = next_ok ? row[s + 1] : (int64_t)-100;
        }
      }
      sep_before += sep_tot;
      sel_before += sel_tot;
    }
    // per-row weights of the compact rows of this sample: w_scale / max(N_sample, 1)
    if (p.row_w) {
      const float w = p.w_scale / (float)max(sel_before, 1);
      for (int n = base + tid; n < base + sel_before && n < p.cap; n += RP_THREADS) p.row_w[n] = w;
    }
    if (tid == 0) {
      if (p.row_counts) p.row_counts[b] = sel_before;
      if (p.sep_counts) p.sep_counts[b] = sep_before;
    }
    base += sel_before;
  }
  // tail of the plan: -1 rows, zero weights
  for (int n = base + tid; n < p.cap; n += RP_THREADS) {
    if (p.rows) p.rows[n] = -1;
    if (p.row_w) p.row_w[n] = 0.f;
  }
  if (tid == 0 && p.count) *p.count = base;
}

}  // namespace
}  // namespace mimic

using namespace mimic;

extern "C" int mimic_row_plan(const int64_t* ids, int32_t B, int32_t S, int64_t pad_id, int64_t bos_id, int32_t left_padding,
                              int32_t num_separator, int32_t select_mode, int32_t keep_bos, float w_scale,
                              uint8_t* attn_mask, uint8_t* label_mask, int32_t* rows, int32_t cap,
                              int32_t* row_to_compact, float* row_w, int64_t* lm_labels, const int32_t* pair_src,
                              int32_t pair_cap, int32_t* paired, int32_t* count, int32_t* row_counts,
                              int32_t* sep_counts, void* stream) {
  MIMIC_CHECK_ARG(ids != nullptr && B > 0 && S > 0, "ids is NULL or empty batch (B=%d S=%d)", B, S);
  MIMIC_CHECK_ARG((int64_t)B * S < (1ll << 31), "B * S = %lld does not fit the int32 row indices", (long long)B * S);
  MIMIC_CHECK_ARG(select_mode == 0 || select_mode == 1, "select_mode must be 0 (label mask) or 1 (non-pad, non-bos), got %d",
                  select_mode);
  MIMIC_CHECK_ARG(cap >= 0 && (cap == 0 || rows || row_w), "cap without rows / row_w");
  MIMIC_CHECK_ARG(!paired || pair_src, "paired needs pair_src");
  RowPlanParams p{ids, B, S, pad_id, bos_id, left_padding, num_separator, select_mode, keep_bos, w_scale,
                  attn_mask, label_mask, rows, cap, row_to_compact, row_w, lm_labels, pair_src, pair_cap, paired,
                  count, row_counts, sep_counts};
  row_plan_kernel<<<1, RP_THREADS, 0, reinterpret_cast<cudaStream_t>(stream)>>>(p);
  MIMIC_LAUNCH_OK("row_plan_kernel");
  return MIMIC_OK;
}
