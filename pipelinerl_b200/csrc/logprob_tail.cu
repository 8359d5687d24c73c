// Hot path (2b, generic-model variant): per-token log-probability and exact entropy
// from MATERIALISED fp32 logits, single pass, plus the backward.
//
// Replaces pipelinerl/finetune/rl/__init__.py:207-233: the reference makes a
// `/temperature` copy, a gather, a logsumexp and a 38-chunk entropy loop (>= 3
// full passes over 608 KB/token of logits, plus softmax again in backward).
// Here: forward reads each logit once (online max / sum / sum(e*z)), backward
// reads once and writes once.  The fused-head kernel (lm_head.cu) removes the
// logits from HBM altogether; this kernel serves models that hand us logits
// (any torch module, as rl_step's contract allows).
//
//   z = logits / temperature
//   lse = logsumexp(z);  new_lp = z[next] - lse
//   H = -sum p log p = lse - sum(p * z)
//   dz_v = g_lp (1[v==next] - p_v) - g_H p_v (log p_v + H)
#include "prl_common.cuh"
#include <math.h>

namespace prl {
namespace {

constexpr int kThreads = 512;

struct Online {
  float m, s, u;  // running max, sum exp(z-m), sum exp(z-m)*z
};
__device__ __forceinline__ void online_add(Online& a, float z) {
  if (z > a.m) {
    const float r = __expf(a.m - z);  // exp(-inf)=0 on first element
    a.s = a.s * r + 1.f;
    a.u = a.u * r + z;
    a.m = z;
  } else {
    const float e = __expf(z - a.m);
    a.s += e;
    a.u = fmaf(e, z, a.u);
  }
}
__device__ __forceinline__ Online online_merge(const Online& a, const Online& b) {
  Online o;
  o.m = fmaxf(a.m, b.m);
  const float ra = (a.m == -INFINITY) ? 0.f : __expf(a.m - o.m);
  const float rb = (b.m == -INFINITY) ? 0.f : __expf(b.m - o.m);
  o.s = a.s * ra + b.s * rb;
  o.u = a.u * ra + b.u * rb;
  return o;
}

__global__ void __launch_bounds__(kThreads) tail_fwd_kernel(const float* __restrict__ logits, int64_t V,
                                                           int64_t row_stride, const int64_t* __restrict__ ids,
                                                           float inv_temp, float* __restrict__ new_lp,
                                                           float* __restrict__ entropy, float* __restrict__ lse_out) {
  const int64_t row = blockIdx.x;
  const float* z = logits + row * row_stride;
  Online acc{-INFINITY, 0.f, 0.f};
  const bool vec = ((reinterpret_cast<uintptr_t>(z) & 15) == 0);
  const int64_t v4 = vec ? (V / 4) : 0;
  for (int64_t i = threadIdx.x; i < v4; i += kThreads) {
    const float4 x = ld_stream_f4(reinterpret_cast<const float4*>(z) + i);
    online_add(acc, x.x * inv_temp); online_add(acc, x.y * inv_temp);
    online_add(acc, x.z * inv_temp); online_add(acc, x.w * inv_temp);
  }
  for (int64_t i = v4 * 4 + threadIdx.x; i < V; i += kThreads) online_add(acc, z[i] * inv_temp);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    Online other;
    other.m = __shfl_xor_sync(0xffffffffu, acc.m, o);
    other.s = __shfl_xor_sync(0xffffffffu, acc.s, o);
    other.u = __shfl_xor_sync(0xffffffffu, acc.u, o);
    acc = online_merge(acc, other);
  }
  __shared__ Online s_w[kThreads / kWarp];
  if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    Online t = s_w[0];
    for (int w = 1; w < kThreads / kWarp; ++w) t = online_merge(t, s_w[w]);
    const float lse = t.m + logf(t.s);
    const int64_t nxt = ids[row + 1];
    float lp = NAN;  // out-of-vocabulary ids surface through the non-finite check, as an index error would
    if (nxt >= 0 && nxt < V) lp = z[nxt] * inv_temp - lse;
    new_lp[row] = lp;
    if (entropy) entropy[row] = lse - t.u / t.s;
    if (lse_out) lse_out[row] = lse;
  }
}

__global__ void __launch_bounds__(kThreads) tail_bwd_kernel(const float* __restrict__ logits, int64_t V,
                                                           int64_t row_stride, int64_t n_rows,
                                                           const int64_t* __restrict__ ids, float inv_temp,
                                                           const float* __restrict__ lse_in,
                                                           const float* __restrict__ entropy,
                                                           const float* __restrict__ g_lp,
                                                           const float* __restrict__ g_ent,
                                                           float* __restrict__ dlogits, int64_t d_stride) {
  const int64_t row = blockIdx.x;
  float* d = dlogits + row * d_stride;
  if (row >= n_rows) {  // the last position has no target (logits[:, :-1])
    for (int64_t i = threadIdx.x; i < V; i += kThreads) d[i] = 0.f;
    return;
  }
  const float* z = logits + row * row_stride;
  const float lse = lse_in[row];
  const float gl = g_lp ? g_lp[row] : 0.f;
  const float ge = g_ent ? g_ent[row] : 0.f;
  const float H = (g_ent && entropy) ? entropy[row] : 0.f;
  const int64_t nxt = ids[row + 1];
  const bool vec = ((reinterpret_cast<uintptr_t>(z) & 15) == 0) && ((reinterpret_cast<uintptr_t>(d) & 15) == 0);
  const int64_t v4 = vec ? (V / 4) : 0;
  auto one = [&](float zi, int64_t idx) -> float {
    const float lp = zi * inv_temp - lse;
    const float p = __expf(lp);
    float g = -gl * p - ge * p * (lp + H);
    if (idx == nxt) g += gl;
    return g * inv_temp;
  };
  for (int64_t i = threadIdx.x; i < v4; i += kThreads) {
    const float4 x = ld_stream_f4(reinterpret_cast<const float4*>(z) + i);
    float4 o;
    o.x = one(x.x, i * 4); o.y = one(x.y, i * 4 + 1); o.z = one(x.z, i * 4 + 2); o.w = one(x.w, i * 4 + 3);
    st_stream_f4(reinterpret_cast<float4*>(d) + i, o);
  }
  for (int64_t i = v4 * 4 + threadIdx.x; i < V; i += kThreads) d[i] = one(z[i], i);
}

}  // namespace
}  // namespace prl

using namespace prl;

extern "C" int prl_logprob_tail_fwd(const float* logits, int64_t T, int64_t V, int64_t row_stride,
                                    const int64_t* input_ids, float temperature, float* new_logprobs,
                                    float* entropy, float* lse, prl_stream_t stream_) {
  PRL_CHECK_ARG(T >= 1 && V >= 1 && row_stride >= V, "prl_logprob_tail_fwd: bad shape T=%lld V=%lld stride=%lld",
                (long long)T, (long long)V, (long long)row_stride);
  PRL_CHECK_ARG(temperature > 0.f, "prl_logprob_tail_fwd: temperature must be > 0");
  if (T == 1) return PRL_OK;
  PRL_CHECK_ARG(logits && input_ids && new_logprobs, "prl_logprob_tail_fwd: NULL argument");
  tail_fwd_kernel<<<(unsigned)(T - 1), kThreads, 0, (cudaStream_t)stream_>>>(
      logits, V, row_stride, input_ids, 1.f / temperature, new_logprobs, entropy, lse);
  PRL_LAUNCH_CHECK();
  return PRL_OK;
}

extern "C" int prl_logprob_tail_bwd(const float* logits, int64_t T, int64_t V, int64_t row_stride,
                                    const int64_t* input_ids, float temperature, const float* lse,
                                    const float* entropy, const float* g_logprobs, const float* g_entropy,
                                    float* dlogits, int64_t dlogits_stride, prl_stream_t stream_) {
  PRL_CHECK_ARG(T >= 1 && V >= 1 && row_stride >= V && dlogits_stride >= V, "prl_logprob_tail_bwd: bad shape");
  PRL_CHECK_ARG(temperature > 0.f, "prl_logprob_tail_bwd: temperature must be > 0");
  PRL_CHECK_ARG(logits && input_ids && lse && dlogits, "prl_logprob_tail_bwd: NULL argument");
  PRL_CHECK_ARG(!g_entropy || entropy, "prl_logprob_tail_bwd: g_entropy needs the forward entropy");
  tail_bwd_kernel<<<(unsigned)T, kThreads, 0, (cudaStream_t)stream_>>>(
      logits, V, row_stride, T - 1, input_ids, 1.f / temperature, lse, entropy, g_logprobs, g_entropy, dlogits,
      dlogits_stride);
  PRL_LAUNCH_CHECK();
  return PRL_OK;
}

// Backward over an explicit (row, target) list: rows of a recomputed logits CHUNK (the fused head's backward
// recomputes logits chunk by chunk instead of keeping 608 KB/token alive between forward and backward).
extern "C" int prl_logprob_rows_bwd(const float* logits, int64_t n_rows, int64_t V, int64_t row_stride,
                                    const int64_t* targets, float temperature, const float* lse, const float* entropy,
                                    const float* g_logprobs, const float* g_entropy, float* dlogits,
                                    int64_t dlogits_stride, prl_stream_t stream_) {
  PRL_CHECK_ARG(n_rows >= 1 && V >= 1 && row_stride >= V && dlogits_stride >= V, "prl_logprob_rows_bwd: bad shape");
  PRL_CHECK_ARG(temperature > 0.f, "prl_logprob_rows_bwd: temperature must be > 0");
  PRL_CHECK_ARG(logits && targets && lse && dlogits, "prl_logprob_rows_bwd: NULL argument");
  PRL_CHECK_ARG(!g_entropy || entropy, "prl_logprob_rows_bwd: g_entropy needs the forward entropy");
  // the kernel reads ids[row + 1]: bias the pointer by one element so that it sees targets[row]
  tail_bwd_kernel<<<(unsigned)n_rows, kThreads, 0, (cudaStream_t)stream_>>>(
      logits, V, row_stride, n_rows, targets - 1, 1.f / temperature, lse, entropy, g_logprobs, g_entropy, dlogits,
      dlogits_stride);
  PRL_LAUNCH_CHECK();
  return PRL_OK;
}
