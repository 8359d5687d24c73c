// Hot path (2c): fused AdamW over a flat parameter arena.
//
// Replaces, in one read-modify-write pass over the optimizer state:
//   torch.optim.AdamW built by pipelinerl/finetune/optim.py:8-29 (decay groups:
//   names containing "bias" / "LayerNorm.weight" get weight_decay 0),
//   clip_grad_norm_ at finetune_loop.py:739 and the fp32->bf16 re-cast that the
//   DeepSpeed bf16 optimizer performs at finetune_loop.py:727-736.
//
// Two launches, no host sync: (1) sum-of-squares partials of the (scaled)
// gradient, (2) the update; every block of (2) re-reduces the few hundred
// partials in a fixed order, so the clip coefficient is deterministic.
//
// HBM-bound: 14 B/param read (g bf16 2, master 4, m 4, v 4) + 14 B/param
// written (master 4, m 4, v 4, bf16 2) = 28 B/param (+2 with the bf16 "lo"
// residual that makes the output head fp32-equivalent).
#include "prl_common.cuh"
#include <math.h>

namespace prl {
namespace {

constexpr int kThreads = 256;
constexpr int kVec = 4;                         // elements per thread per iteration
constexpr int kChunk = kThreads * kVec * 4;     // 4096 elements per block iteration
constexpr int kMaxNormBlocks = 148 * 8;

struct AdamWorkspace {
  double partial[kMaxNormBlocks];
  int n_partials;
};

template <bool kBf16>
__device__ __forceinline__ void load_grad4(const void* g, int64_t i, float scale, float out[4]) {
  if (kBf16) {
    const uint2 raw = ld_stream_u2(reinterpret_cast<const uint2*>(static_cast<const __nv_bfloat16*>(g) + i));
    out[0] = bf16_bits_to_float(raw.x & 0xffffu) * scale;
    out[1] = bf16_bits_to_float(raw.x >> 16) * scale;
    out[2] = bf16_bits_to_float(raw.y & 0xffffu) * scale;
    out[3] = bf16_bits_to_float(raw.y >> 16) * scale;
  } else {
    const float4 raw = ld_stream_f4(reinterpret_cast<const float4*>(static_cast<const float*>(g) + i));
    out[0] = raw.x * scale; out[1] = raw.y * scale; out[2] = raw.z * scale; out[3] = raw.w * scale;
  }
}
template <bool kBf16>
__device__ __forceinline__ float load_grad1(const void* g, int64_t i, float scale) {
  if (kBf16) return __bfloat162float(static_cast<const __nv_bfloat16*>(g)[i]) * scale;
  return static_cast<const float*>(g)[i] * scale;
}

template <bool kBf16>
__global__ void __launch_bounds__(kThreads) grad_sumsq_kernel(const void* __restrict__ g, int64_t n, float scale,
                                                             AdamWorkspace* ws) {
  double acc = 0.0;
  const int64_t n4 = n / 4;
  for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < n4; v += (int64_t)gridDim.x * blockDim.x) {
    float x[4];
    load_grad4<kBf16>(g, v * 4, scale, x);
    float s = x[0] * x[0];
    s = fmaf(x[1], x[1], s); s = fmaf(x[2], x[2], s); s = fmaf(x[3], x[3], s);
    acc += (double)s;
  }
  if (blockIdx.x == 0) {
    for (int64_t i = n4 * 4 + threadIdx.x; i < n; i += blockDim.x) {
      const float x = load_grad1<kBf16>(g, i, scale);
      acc += (double)x * (double)x;
    }
  }
  __shared__ double s_w[kThreads / kWarp];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0;
    for (int i = 0; i < kThreads / kWarp; ++i) t += s_w[i];
    ws->partial[blockIdx.x] = t;
    if (blockIdx.x == 0) ws->n_partials = gridDim.x;
  }
}

struct UpdateConsts {
  float lr, beta1, beta2, eps;
  float weight_decay;     // decay FACTOR 1 - lr*wd (1 for no-decay tensors)
  float step_size;        // lr / (1 - beta1^t)
  float bc2_sqrt;         // sqrt(1 - beta2^t)
  // torch evaluates these scalars in double precision before the fp32 tensor op; 1.f - 0.999f would be
  // off by 4.7e-5 relative
  float one_minus_beta1, one_minus_beta2;
  float max_grad_norm, grad_scale;
};

__device__ __forceinline__ void adam_elem(float g, float& p, float& m, float& v, const UpdateConsts& k, float wd /* decay factor */) {
  // same operation order as torch.optim.adamw (_single_tensor_adam)
  p = p * wd;                                  // wd = 1 - lr * weight_decay, evaluated in double on the host
  m = m + (g - m) * k.one_minus_beta1;         // lerp_(grad, 1 - beta1)
  v = v * k.beta2 + k.one_minus_beta2 * g * g; // mul_(beta2).addcmul_(g, g, 1 - beta2)
  const float denom = sqrtf(v) / k.bc2_sqrt + k.eps;
  p = p - k.step_size * (m / denom);
}

template <bool kBf16>
__global__ void __launch_bounds__(kThreads) adamw_kernel(prl_adamw_args a, UpdateConsts k, const AdamWorkspace* ws,
                                                         float* __restrict__ grad_norm_out, int n_norm_blocks) {
  // global grad norm: fixed-order re-reduction of the partials (identical in every block)
  __shared__ float s_clip;
  __shared__ double s_red[kThreads / kWarp];
  {
    double t = 0;
    for (int i = threadIdx.x; i < n_norm_blocks; i += blockDim.x) t += ws->partial[i];
    t = warp_sum(t);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = t;
    __syncthreads();
    if (threadIdx.x == 0) {
      double tot = 0;
      for (int i = 0; i < kThreads / kWarp; ++i) tot += s_red[i];
      const float norm = (float)sqrt(tot);
      float clip = 1.f;
      if (k.max_grad_norm > 0.f) {
        // torch.nn.utils.clip_grad_norm_: clip_coef = max_norm / (total_norm + 1e-6), clamped to 1
        clip = fminf(k.max_grad_norm / (norm + 1e-6f), 1.f);
      }
      s_clip = clip;
      if (blockIdx.x == 0 && grad_norm_out) *grad_norm_out = norm;
    }
    __syncthreads();
  }
  const float gscale = k.grad_scale * s_clip;

  __shared__ int s_tensor;
  const int64_t n_chunks = (a.n + kChunk - 1) / kChunk;
  for (int64_t c = blockIdx.x; c < n_chunks; c += gridDim.x) {
    const int64_t base = c * kChunk;
    const int64_t end = (base + kChunk < a.n) ? base + kChunk : a.n;
    // decay group lookup: binary search for the tensor containing `base`
    if (threadIdx.x == 0) {
      int lo = 0, hi = a.n_tensors;  // invariant: offsets[lo] <= base < offsets[hi]
      while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (a.tensor_offsets[mid] <= base) lo = mid; else hi = mid;
      }
      s_tensor = lo;
    }
    __syncthreads();
    int tix = s_tensor;
    int64_t t_end = a.tensor_offsets[tix + 1];
    __syncthreads();

#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int64_t i = base + ((int64_t)it * kThreads + threadIdx.x) * kVec;
      if (i >= end) break;
      while (i >= t_end) { ++tix; t_end = a.tensor_offsets[tix + 1]; }
      const bool fast = (i + kVec <= end) && (i + kVec <= t_end) && ((i & 3) == 0);
      if (fast) {
        float g[4];
        load_grad4<kBf16>(a.grad, i, gscale, g);
        float4 p = *reinterpret_cast<const float4*>(a.master + i);
        float4 m = *reinterpret_cast<const float4*>(a.exp_avg + i);
        float4 v = *reinterpret_cast<const float4*>(a.exp_avg_sq + i);
        const float wd = a.tensor_no_decay[tix] ? 1.f : k.weight_decay;
        adam_elem(g[0], p.x, m.x, v.x, k, wd);
        adam_elem(g[1], p.y, m.y, v.y, k, wd);
        adam_elem(g[2], p.z, m.z, v.z, k, wd);
        adam_elem(g[3], p.w, m.w, v.w, k, wd);
        *reinterpret_cast<float4*>(a.master + i) = p;
        *reinterpret_cast<float4*>(a.exp_avg + i) = m;
        *reinterpret_cast<float4*>(a.exp_avg_sq + i) = v;
        if (a.param_bf16) {
          const uint32_t b0 = float_to_bf16_bits(p.x), b1 = float_to_bf16_bits(p.y);
          const uint32_t b2 = float_to_bf16_bits(p.z), b3 = float_to_bf16_bits(p.w);
          st_stream_u2(reinterpret_cast<uint2*>(static_cast<__nv_bfloat16*>(a.param_bf16) + i),
                       make_uint2(b0 | (b1 << 16), b2 | (b3 << 16)));
          if (a.param_bf16_lo) {
            const uint32_t l0 = float_to_bf16_bits(p.x - bf16_bits_to_float(b0));
            const uint32_t l1 = float_to_bf16_bits(p.y - bf16_bits_to_float(b1));
            const uint32_t l2 = float_to_bf16_bits(p.z - bf16_bits_to_float(b2));
            const uint32_t l3 = float_to_bf16_bits(p.w - bf16_bits_to_float(b3));
            st_stream_u2(reinterpret_cast<uint2*>(static_cast<__nv_bfloat16*>(a.param_bf16_lo) + i),
                         make_uint2(l0 | (l1 << 16), l2 | (l3 << 16)));
          }
        }
      } else {
        int tj = tix;
        int64_t tj_end = t_end;
        for (int64_t j = i; j < i + kVec && j < end; ++j) {
          while (j >= tj_end) { ++tj; tj_end = a.tensor_offsets[tj + 1]; }
          const float g = load_grad1<kBf16>(a.grad, j, gscale);
          float p = a.master[j], m = a.exp_avg[j], v = a.exp_avg_sq[j];
          adam_elem(g, p, m, v, k, a.tensor_no_decay[tj] ? 1.f : k.weight_decay);
          a.master[j] = p; a.exp_avg[j] = m; a.exp_avg_sq[j] = v;
          if (a.param_bf16) {
            const __nv_bfloat16 hi = __float2bfloat16_rn(p);
            static_cast<__nv_bfloat16*>(a.param_bf16)[j] = hi;
            if (a.param_bf16_lo)
              static_cast<__nv_bfloat16*>(a.param_bf16_lo)[j] = __float2bfloat16_rn(p - __bfloat162float(hi));
          }
        }
      }
    }
  }
}

}  // namespace
}  // namespace prl

using namespace prl;

extern "C" size_t prl_adamw_workspace_bytes(void) { return sizeof(AdamWorkspace); }

extern "C" int prl_adamw_step(const prl_adamw_args* a, float* grad_norm_out, void* workspace,
                              size_t workspace_bytes, prl_stream_t stream_) {
  PRL_CHECK_ARG(a && workspace, "prl_adamw_step: NULL argument");
  PRL_CHECK_ARG(workspace_bytes >= sizeof(AdamWorkspace), "prl_adamw_step: workspace too small");
  PRL_CHECK_ARG(a->n >= 0, "prl_adamw_step: n < 0");
  if (a->n == 0) return PRL_OK;
  PRL_CHECK_ARG(a->master && a->exp_avg && a->exp_avg_sq && a->grad, "prl_adamw_step: NULL state pointer");
  PRL_CHECK_ARG(a->n_tensors >= 1 && a->tensor_offsets && a->tensor_no_decay, "prl_adamw_step: missing tensor table");
  PRL_CHECK_ARG(a->step >= 1, "prl_adamw_step: step is 1-based");
  PRL_CHECK_ARG(((uintptr_t)a->master % 16 == 0) && ((uintptr_t)a->exp_avg % 16 == 0) &&
                    ((uintptr_t)a->exp_avg_sq % 16 == 0) && ((uintptr_t)a->grad % 16 == 0) &&
                    (!a->param_bf16 || (uintptr_t)a->param_bf16 % 8 == 0) &&
                    (!a->param_bf16_lo || (uintptr_t)a->param_bf16_lo % 8 == 0),
                "prl_adamw_step: arena pointers must be 16-byte aligned");
  cudaStream_t stream = (cudaStream_t)stream_;
  AdamWorkspace* ws = (AdamWorkspace*)workspace;

  const float gs = a->grad_scale == 0.f ? 1.f : a->grad_scale;
  int norm_blocks = (int)((a->n / 4 + kThreads - 1) / kThreads);
  if (norm_blocks < 1) norm_blocks = 1;
  const int cap = num_sms() * 8 < kMaxNormBlocks ? num_sms() * 8 : kMaxNormBlocks;
  if (norm_blocks > cap) norm_blocks = cap;
  if (a->grad_is_bf16) grad_sumsq_kernel<true><<<norm_blocks, kThreads, 0, stream>>>(a->grad, a->n, gs, ws);
  else grad_sumsq_kernel<false><<<norm_blocks, kThreads, 0, stream>>>(a->grad, a->n, gs, ws);
  PRL_LAUNCH_CHECK();

  UpdateConsts k;
  k.lr = (float)a->lr; k.beta1 = (float)a->beta1; k.beta2 = (float)a->beta2; k.eps = (float)a->eps;
  k.weight_decay = (float)(1.0 - a->lr * a->weight_decay);  // decay FACTOR
  k.one_minus_beta1 = (float)(1.0 - a->beta1);
  k.one_minus_beta2 = (float)(1.0 - a->beta2);
  const double bc1 = 1.0 - pow(a->beta1, (double)a->step);
  const double bc2 = 1.0 - pow(a->beta2, (double)a->step);
  k.step_size = (float)(a->lr / bc1);
  k.bc2_sqrt = (float)sqrt(bc2);
  k.max_grad_norm = a->max_grad_norm;
  k.grad_scale = gs;

  const int64_t n_chunks = (a->n + kChunk - 1) / kChunk;
  int blocks = (int)(n_chunks < (int64_t)num_sms() * 8 ? n_chunks : (int64_t)num_sms() * 8);
  if (a->grad_is_bf16) adamw_kernel<true><<<blocks, kThreads, 0, stream>>>(*a, k, ws, grad_norm_out, norm_blocks);
  else adamw_kernel<false><<<blocks, kThreads, 0, stream>>>(*a, k, ws, grad_norm_out, norm_blocks);
  PRL_LAUNCH_CHECK();
  return PRL_OK;
}
