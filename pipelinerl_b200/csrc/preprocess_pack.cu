// GPU-resident preprocess: rollout samples -> RL columns -> ONE packed micro-batch row, on the learner's
// GPU, from a compact binary record (pipelinerl_b200/records.py) -- replaces the pandas pipeline of
// populate_rl_data (pipelinerl/finetune/rl/__init__.py:453-570), the Python-list collate_packed
// (finetune/data.py:215-283) and the JSONL round trip of twelve [1, T] tensors (streams.py:269-277,
// finetune_loop.py:109).
//
// Bit-exactness: the reference computes the per-(group, step) statistics in float64 with pandas' groupby kernels
// (Kahan-compensated `sum`, Welford `std` with ddof = 1, rows in dataset order) and rounds to float32 when collate_packed
// builds the tensors.  stats_kernel runs exactly those recurrences in double, in the same row order; every float column
// is then a double -> float round-to-nearest of the same double the reference holds.  Integer columns are copies.
//
//   stats_kernel   one CTA: rollout token totals, per-group mean rollout length, per-(group, step) Kahan sum / count /
//                  Welford std; then per packed sample: leave-one-out advantage, group_tokens, num_labels, reward
//   overflow_kernel one CTA per packed sample: finish_reason / finished / "eos in input_ids" rule (:541-554)
//   fill_kernel    one thread per output token: binary search of its sample, all twelve columns + seq_boundaries
// HBM-bound by the 72 B/token of output columns; the record itself is 12-16 B/token.
#include "prl_common.cuh"

namespace prl {
namespace {

struct SampleCols { float reward, advantage, group_tokens, overflow, num_labels; };

__global__ void __launch_bounds__(256) preprocess_stats_kernel(prl_mb_record r, int divide_by_std, double* rollout_tokens,
                                                              double* group_mean, double* stat_sum, double* stat_std,
                                                              int* stat_cnt, SampleCols* cols) {
  const int tid = threadIdx.x;
  // rollout_tokens[(group, rollout_index)] = total tokens of that rollout's samples (:466-472)
  for (int s = tid; s < r.n_rollout_slots; s += blockDim.x) {
    double t = 0.0;
    int g = -1;
    for (int i = 0; i < r.n_chunk; ++i)
      if (r.rollout_slot[i] == s) { t += (double)r.n_tok_all[i]; g = r.group_slot[i]; }
    rollout_tokens[s] = t;
    rollout_tokens[r.n_rollout_slots + s] = (double)g;      // owning group of the slot
  }
  __syncthreads();
  // group_tokens = mean over the group's rollouts (:473-479); integer-valued sums: exact in any order
  for (int g = tid; g < r.n_groups; g += blockDim.x) {
    double t = 0.0;
    int c = 0;
    for (int s = 0; s < r.n_rollout_slots; ++s)
      if ((int)rollout_tokens[r.n_rollout_slots + s] == g) { t += rollout_tokens[s]; ++c; }
    group_mean[g] = c > 0 ? t / (double)c : 0.0;
  }
  // per (group, step_index): pandas groupby sum (Kahan), count, std (Welford, ddof = 1) in row order (:480-488)
  for (int s = tid; s < r.n_stat_slots; s += blockDim.x) {
    double sum = 0.0, comp = 0.0, mean = 0.0, m2 = 0.0;
    int n = 0;
    for (int i = 0; i < r.n_chunk; ++i) {
      if (r.stat_slot[i] != s) continue;
      const double v = r.reward[i];
      const double y = v - comp;
      const double t = sum + y;
      comp = t - sum - y;
      if (comp != comp) comp = 0.0;
      sum = t;
      ++n;
      const double old = mean;
      mean += (v - old) / (double)n;
      m2 = __dadd_rn(m2, __dmul_rn(v - mean, v - old));   // no FMA contraction: pandas multiplies, then adds
    }
    stat_sum[s] = sum;
    stat_cnt[s] = n;
    stat_std[s] = n > 1 ? sqrt(m2 / (double)(n - 1)) : 0.0;   // pandas: NaN for one member -> np.nan_to_num -> 0 (:513-519)
  }
  __syncthreads();
  for (int p = tid; p < r.n_pack; p += blockDim.x) {
    const int i = r.pack_idx[p];
    const int s = r.stat_slot[i];
    const double r0 = r.reward[i];
    const int cnt = stat_cnt[s];
    const double baseline = cnt > 1 ? (stat_sum[s] - r0) / (double)(cnt - 1) : r0;
    double std = stat_std[s];
    if (std != std) std = 0.0;
    const double adv = divide_by_std ? (r0 - baseline) / (std + 1e-4) : (r0 - baseline);
    SampleCols c;
    c.reward = __double2float_rn(r0);
    c.advantage = __double2float_rn(adv);
    c.group_tokens = __double2float_rn(group_mean[r.group_slot[i]]);
    c.num_labels = (float)(r.lp_off[p + 1] - r.lp_off[p]);
    c.overflow = cols[p].overflow;                  // written by overflow_kernel (launched first)
    cols[p] = c;
  }
}

// flags: bit 0 finished; bits 1-2 finish_reason: 0 none / other, 1 "length", 2 "stop" | "content_filter"
__global__ void __launch_bounds__(256) preprocess_overflow_kernel(prl_mb_record r, int eos_id, SampleCols* cols) {
  const int p = blockIdx.x;
  const int flags = r.pack_flags[p];
  const int reason = (flags >> 1) & 3;
  float ov;
  if (reason == 1) ov = 1.f;
  else if (reason == 2) ov = 0.f;
  else if (flags & 1) ov = 0.f;
  else {
    int found = 0;
    for (int j = r.tok_off[p] + threadIdx.x; j < r.tok_off[p + 1]; j += blockDim.x) found |= (r.input_ids[j] == eos_id);
    found = __syncthreads_or(found);
    ov = found ? 0.f : 1.f;
  }
  if (threadIdx.x == 0) cols[p].overflow = ov;
}

__global__ void __launch_bounds__(256) preprocess_fill_kernel(prl_mb_record r, prl_mb_columns o, int eos_id,
                                                             const SampleCols* __restrict__ cols) {
  const int T = r.total_tok + r.padding;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t == 0) {
    for (int p = 0; p <= r.n_pack; ++p) o.seq_boundaries[p] = r.tok_off[p];
    if (r.padding > 0) o.seq_boundaries[r.n_pack + 1] = T;
  }
  if (t >= T) return;
  o.attention_mask[t] = 1;
  if (t >= r.total_tok) {            // pad-to-seq_parallel sentinel sample (finetune/utils.py:46-78)
    const int j = t - r.total_tok;
    o.input_ids[t] = eos_id; o.labels[t] = -100; o.position_ids[t] = j; o.segment_ids[t] = r.n_pack;
    o.rewards[t] = 0.f; o.advantages[t] = 0.f; o.ref_logprobs[t] = 0.f; o.old_logprobs[t] = 0.f;
    o.group_tokens[t] = 1.f; o.num_labels[t] = 1.f; o.overflow[t] = 0.f;
    return;
  }
  int lo = 0, hi = r.n_pack;         // invariant: tok_off[lo] <= t < tok_off[hi]
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (r.tok_off[mid] <= t) lo = mid; else hi = mid;
  }
  const int p = lo;
  const int j = t - r.tok_off[p];
  const int n_tok = r.tok_off[p + 1] - r.tok_off[p];
  const int n_lp = r.lp_off[p + 1] - r.lp_off[p];
  o.input_ids[t] = r.input_ids[t];
  // the first token of every sample but the first has no in-sample predecessor: never a target (data.py:264-265)
  o.labels[t] = (j == 0 && p > 0) ? -100 : (int64_t)r.labels[t];
  o.position_ids[t] = j;
  o.segment_ids[t] = p;
  const SampleCols c = cols[p];
  o.rewards[t] = c.reward;
  o.advantages[t] = c.advantage;
  o.group_tokens[t] = c.group_tokens;
  o.num_labels[t] = c.num_labels;
  o.overflow[t] = c.overflow;
  // logprobs are right-aligned to the end of the sample, zeros over the prompt (rl/__init__.py:586-589)
  const int k = j - (n_tok - n_lp);
  const float old_lp = k >= 0 ? r.logprobs[r.lp_off[p] + k] : 0.f;
  o.old_logprobs[t] = old_lp;
  o.ref_logprobs[t] = r.ref_logprobs ? (k >= 0 ? r.ref_logprobs[r.lp_off[p] + k] : 0.f) : old_lp;
}

inline size_t align16(size_t x) { return (x + 15) & ~(size_t)15; }

}  // namespace
}  // namespace prl

using namespace prl;

extern "C" size_t prl_preprocess_workspace_bytes(int32_t n_pack, int32_t n_stat_slots, int32_t n_rollout_slots,
                                                 int32_t n_groups) {
  return align16((size_t)2 * n_rollout_slots * 8) + align16((size_t)n_groups * 8) + 2 * align16((size_t)n_stat_slots * 8) +
         align16((size_t)n_stat_slots * 4) + align16((size_t)n_pack * sizeof(SampleCols)) + 64;
}

extern "C" int prl_preprocess_pack(const prl_mb_record* rec, int32_t divide_advantage_by_std, int32_t eos_token_id,
                                   const prl_mb_columns* out, void* workspace, size_t workspace_bytes,
                                   prl_stream_t stream_) {
  PRL_CHECK_ARG(rec && out && workspace, "prl_preprocess_pack: NULL argument");
  const prl_mb_record& r = *rec;
  PRL_CHECK_ARG(r.n_chunk >= 1 && r.n_pack >= 1 && r.total_tok >= 1 && r.padding >= 0 && r.n_stat_slots >= 1 &&
                    r.n_rollout_slots >= 1 && r.n_groups >= 1,
                "prl_preprocess_pack: empty record");
  PRL_CHECK_ARG(r.reward && r.stat_slot && r.rollout_slot && r.group_slot && r.n_tok_all && r.pack_idx && r.pack_flags &&
                    r.tok_off && r.lp_off && r.input_ids && r.labels && r.logprobs,
                "prl_preprocess_pack: NULL record section");
  PRL_CHECK_ARG(out->input_ids && out->labels && out->attention_mask && out->position_ids && out->segment_ids &&
                    out->rewards && out->advantages && out->ref_logprobs && out->old_logprobs && out->group_tokens &&
                    out->num_labels && out->overflow && out->seq_boundaries,
                "prl_preprocess_pack: NULL output column");
  PRL_CHECK_ARG(workspace_bytes >= prl_preprocess_workspace_bytes(r.n_pack, r.n_stat_slots, r.n_rollout_slots, r.n_groups),
                "prl_preprocess_pack: workspace too small");
  cudaStream_t stream = (cudaStream_t)stream_;
  uint8_t* w = (uint8_t*)workspace;
  double* rollout_tokens = (double*)w; w += align16((size_t)2 * r.n_rollout_slots * 8);
  double* group_mean = (double*)w; w += align16((size_t)r.n_groups * 8);
  double* stat_sum = (double*)w; w += align16((size_t)r.n_stat_slots * 8);
  double* stat_std = (double*)w; w += align16((size_t)r.n_stat_slots * 8);
  int* stat_cnt = (int*)w; w += align16((size_t)r.n_stat_slots * 4);
  SampleCols* cols = (SampleCols*)w;
  preprocess_overflow_kernel<<<(unsigned)r.n_pack, 256, 0, stream>>>(r, (int)eos_token_id, cols);
  PRL_LAUNCH_CHECK();
  preprocess_stats_kernel<<<1, 256, 0, stream>>>(r, (int)divide_advantage_by_std, rollout_tokens, group_mean, stat_sum,
                                                 stat_std, stat_cnt, cols);
  PRL_LAUNCH_CHECK();
  const int T = r.total_tok + r.padding;
  preprocess_fill_kernel<<<(unsigned)((T + 255) / 256), 256, 0, stream>>>(r, *out, (int)eos_token_id, cols);
  PRL_LAUNCH_CHECK();
  return PRL_OK;
}
