// Hot path (1): the small fused kernels around the tcgen05 GEMMs of one token step.
//
// They replace, for the reference's sampler (vLLM behind pipelinerl/async_llm.py:134),
// vLLM's fused_add_rms_norm / rotary_embedding / reshape_and_cache_flash / silu_and_mul
// custom ops and the sampler + processed_logprobs path (conf/base.yaml:65).  Each one
// consumes the fp32 split-K partials of the preceding GEMM (summed in split order, so
// results are deterministic) and emits the bf16 operand of the next GEMM, i.e. the
// split-K reduction, bias, RoPE, KV-page write, residual add, RMSNorm and SiLU*mul are
// all "epilogue" work on L2-resident activations, never a separate pass over weights.
//
// Numerics contract (restated by oracle/decode_oracle.py): residual stream fp32,
// GEMM operands bf16, accumulation fp32, RoPE angles fp32, KV cache bf16.
#include "prl_common.cuh"
#include <math.h>

namespace prl {
namespace {

// Software prefetch ACROSS kernels through the 126 MB L2: a small epilogue kernel asks L2 to fetch the weights
// of an upcoming latency-bound GEMM (qkv / o_proj: 26-33 MB) while the long HBM-bound kernel in between
// (attention, gate_up, down) runs, so that GEMM then streams from L2 instead of paying DRAM latency cold.
// Lines are requested evict_last so the evict_first streams of the kernels in between do not displace them.
__device__ __forceinline__ void l2_prefetch_range(const void* ptr, size_t bytes, size_t tid, size_t nthreads) {
  if (!ptr) return;
  const char* base = static_cast<const char*>(ptr);
  const size_t n_lines = bytes >> 7;
  for (size_t i = tid; i < n_lines; i += nthreads)
    asm volatile("prefetch.global.L2::evict_last [%0];" ::"l"(base + (i << 7)));
}

__device__ __forceinline__ float block_sum(float v, float* s_red) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) s_red[warp] = v;
  __syncthreads();
  float t = 0.f;
  const int nw = (blockDim.x + 31) >> 5;
  for (int i = 0; i < nw; ++i) t += s_red[i];
  __syncthreads();
  return t;
}

// ---- embedding gather + RMSNorm ------------------------------------------------------
// h[b,:] = embed[token[b],:] ; x[b,:] = bf16(h * rsqrt(mean(h^2) + eps) * g)
__global__ void __launch_bounds__(256) embed_rmsnorm_kernel(const int32_t* __restrict__ tokens,
                                                           const __nv_bfloat16* __restrict__ embed,
                                                           const __nv_bfloat16* __restrict__ gamma, float eps, int H,
                                                           int vocab, float* __restrict__ h,
                                                           __nv_bfloat16* __restrict__ x) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float s_red[32];
  const int b = blockIdx.x;
  int tok = tokens[b];
  if (tok < 0 || tok >= vocab) tok = 0;
  const __nv_bfloat16* row = embed + (int64_t)tok * H;
  float ss = 0.f;
  for (int i = threadIdx.x; i < H; i += blockDim.x) {
    const float v = __bfloat162float(row[i]);
    h[(int64_t)b * H + i] = v;
    ss += v * v;
  }
  const float tot = block_sum(ss, s_red);
  const float r = rsqrtf(tot / (float)H + eps);
  for (int i = threadIdx.x; i < H; i += blockDim.x) {
    const float v = __bfloat162float(row[i]);
    x[(int64_t)b * H + i] = __float2bfloat16_rn(v * r * __bfloat162float(gamma[i]));
  }
}

// ---- split-K reduce + residual add + RMSNorm --------------------------------------------
// h[b,:] += sum_s part[s,b,:] ; x[b,:] = bf16(rmsnorm(h) * g)
// One block per token; each thread owns float4 column groups, issues the residual load and all
// n_split partial loads back to back (independent, L2-resident), so the kernel costs about one
// L2 round trip + one block reduction instead of a serial chain per element.
constexpr int kMaxSplitUnroll = 8;
__global__ void __launch_bounds__(1024) residual_rmsnorm_kernel(const float* __restrict__ part, int n_split, int B,
                                                               int H, const __nv_bfloat16* __restrict__ gamma,
                                                               float eps, float* __restrict__ h,
                                                               __nv_bfloat16* __restrict__ x, const void* pf_ptr,
                                                               size_t pf_bytes) {
  pdl_launch_dependents();
  l2_prefetch_range(pf_ptr, pf_bytes, (size_t)blockIdx.x * blockDim.x + threadIdx.x, (size_t)gridDim.x * blockDim.x);
  pdl_wait();
  extern __shared__ float s_row[];  // H floats (only used when a thread owns more than one group)
  __shared__ float s_red[32];
  const int b = blockIdx.x;
  const int H4 = H >> 2;
  float4* hrow = reinterpret_cast<float4*>(h + (int64_t)b * H);
  float ss = 0.f;
  float4 keep = make_float4(0.f, 0.f, 0.f, 0.f);
  const bool single = H4 <= (int)blockDim.x;
  for (int i = threadIdx.x; i < H4; i += blockDim.x) {
    float4 v = hrow[i];
    for (int s = 0; s < n_split; s += kMaxSplitUnroll) {
      float4 p[kMaxSplitUnroll];
#pragma unroll
      for (int u = 0; u < kMaxSplitUnroll; ++u)
        if (s + u < n_split) p[u] = reinterpret_cast<const float4*>(part + ((int64_t)(s + u) * B + b) * H)[i];
#pragma unroll
      for (int u = 0; u < kMaxSplitUnroll; ++u)
        if (s + u < n_split) { v.x += p[u].x; v.y += p[u].y; v.z += p[u].z; v.w += p[u].w; }
    }
    hrow[i] = v;
    if (single) keep = v; else reinterpret_cast<float4*>(s_row)[i] = v;
    ss += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  }
  const float tot = block_sum(ss, s_red);
  const float r = rsqrtf(tot / (float)H + eps);
  for (int i = threadIdx.x; i < H4; i += blockDim.x) {
    const float4 v = single ? keep : reinterpret_cast<const float4*>(s_row)[i];
    const uint2 gw = reinterpret_cast<const uint2*>(gamma)[i];
    const float g0 = bf16_bits_to_float(gw.x & 0xffffu), g1 = bf16_bits_to_float(gw.x >> 16);
    const float g2 = bf16_bits_to_float(gw.y & 0xffffu), g3 = bf16_bits_to_float(gw.y >> 16);
    const uint32_t o0 = float_to_bf16_bits(v.x * r * g0), o1 = float_to_bf16_bits(v.y * r * g1);
    const uint32_t o2 = float_to_bf16_bits(v.z * r * g2), o3 = float_to_bf16_bits(v.w * r * g3);
    reinterpret_cast<uint2*>(x + (int64_t)b * H)[i] = make_uint2(o0 | (o1 << 16), o2 | (o3 << 16));
  }
}

// ---- split-K reduce + bias + RoPE + KV-page write ---------------------------------------
// part [n_split, B, (n_q + 2 n_kv) * 128]; one block per (token, head), 64 threads = 64 rotation pairs.
// q -> q_out [B, n_q, 128] bf16 ; k, v -> cache rows ((layer*2 + kv) * n_pages + page) * n_kv * PAGE + kvh*PAGE + slot
__global__ void __launch_bounds__(64) qkv_rope_cache_kernel(const float* __restrict__ part, int n_split, int B,
                                                           const __nv_bfloat16* __restrict__ bias, int n_q, int n_kv,
                                                           const int32_t* __restrict__ positions,
                                                           const int32_t* __restrict__ block_table, int max_blocks,
                                                           const int32_t* __restrict__ row_slot,
                                                           const float* __restrict__ inv_freq_tab,
                                                           __nv_bfloat16* __restrict__ q_out,
                                                           __nv_bfloat16* __restrict__ kv_cache, int64_t n_pages,
                                                           int layer, int page_size, const void* pf_ptr,
                                                           size_t pf_bytes) {
  pdl_launch_dependents();
  l2_prefetch_range(pf_ptr, pf_bytes, ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * blockDim.x + threadIdx.x,
                    (size_t)gridDim.x * gridDim.y * blockDim.x);
  pdl_wait();
  constexpr int D = 128;
  const int b = blockIdx.x, head = blockIdx.y, i = threadIdx.x;  // i in [0, 64)
  const int n_heads = n_q + 2 * n_kv;
  const int64_t ncol = (int64_t)n_heads * D;
  const int col = head * D + i;
  float x1 = 0.f, x2 = 0.f;
  for (int s = 0; s < n_split; ++s) {
    const float* p = part + ((int64_t)s * B + b) * ncol;
    x1 += p[col];
    x2 += p[col + 64];
  }
  if (bias) {
    x1 += __bfloat162float(bias[col]);
    x2 += __bfloat162float(bias[col + 64]);
  }
  const int pos = positions[b];
  const bool is_v = head >= n_q + n_kv;
  float o1 = x1, o2 = x2;
  if (!is_v) {
    // NeoX-style rotation of the pair (i, i + 64); inv_freq[i] = 1 / theta^(2i/128) is tabulated by the host
    // with the exact fp32 expression HF's rotary embedding uses, the angle is an fp32 product as there
    const float inv_freq = inv_freq_tab[i];
    float sn, cs;
    sincosf((float)pos * inv_freq, &sn, &cs);
    o1 = x1 * cs - x2 * sn;
    o2 = x2 * cs + x1 * sn;
  }
  if (head < n_q) {
    __nv_bfloat16* q = q_out + ((int64_t)b * n_q + head) * D;
    q[i] = __float2bfloat16_rn(o1);
    q[i + 64] = __float2bfloat16_rn(o2);
  } else {
    const int kv = is_v ? 1 : 0;
    const int kvh = is_v ? head - n_q - n_kv : head - n_q;
    const int slot_row = row_slot ? row_slot[b] : b;
    const int page = block_table[(int64_t)slot_row * max_blocks + pos / page_size];
    const int slot = pos % page_size;
    const int64_t row = (((int64_t)(layer * 2 + kv) * n_pages + page) * n_kv + kvh) * page_size + slot;
    __nv_bfloat16* dst = kv_cache + row * D;
    dst[i] = __float2bfloat16_rn(o1);
    dst[i + 64] = __float2bfloat16_rn(o2);
  }
}

// Same arithmetic, laid out for MANY rows (prefill chunks): one block walks whole token rows, the row's 64
// (cos, sin) pairs are computed once (the per-head kernel above recomputes them for each of the n_q + 2 n_kv heads)
// and every thread handles (head, pair) items with coalesced 4-byte loads.  Bitwise identical results.
__global__ void __launch_bounds__(256) qkv_rope_cache_rows_kernel(const float* __restrict__ part, int n_split, int B,
                                                                const __nv_bfloat16* __restrict__ bias, int n_q, int n_kv,
                                                                const int32_t* __restrict__ positions,
                                                                const int32_t* __restrict__ block_table, int max_blocks,
                                                                const int32_t* __restrict__ row_slot,
                                                                const float* __restrict__ inv_freq_tab,
                                                                __nv_bfloat16* __restrict__ q_out,
                                                                __nv_bfloat16* __restrict__ kv_cache, int64_t n_pages,
                                                                int layer, int page_size) {
  pdl_launch_dependents();
  pdl_wait();
  constexpr int D = 128;
  __shared__ float s_cs[64], s_sn[64];
  const int n_heads = n_q + 2 * n_kv;
  const int64_t ncol = (int64_t)n_heads * D;
  for (int b = blockIdx.x; b < B; b += gridDim.x) {
    const int pos = positions[b];
    __syncthreads();
    if (threadIdx.x < 64) {
      float sn, cs;
      sincosf((float)pos * inv_freq_tab[threadIdx.x], &sn, &cs);
      s_cs[threadIdx.x] = cs;
      s_sn[threadIdx.x] = sn;
    }
    __syncthreads();
    const int slot_row = row_slot ? row_slot[b] : b;
    const int page = block_table[(int64_t)slot_row * max_blocks + pos / page_size];
    const int slot = pos % page_size;
    for (int w = threadIdx.x; w < n_heads * 64; w += blockDim.x) {
      const int head = w >> 6, i = w & 63;
      const int col = head * D + i;
      float x1 = 0.f, x2 = 0.f;
      for (int s = 0; s < n_split; ++s) {
        const float* p = part + ((int64_t)s * B + b) * ncol;
        x1 += p[col];
        x2 += p[col + 64];
      }
      if (bias) {
        x1 += __bfloat162float(bias[col]);
        x2 += __bfloat162float(bias[col + 64]);
      }
      const bool is_v = head >= n_q + n_kv;
      float o1 = x1, o2 = x2;
      if (!is_v) {
        const float cs = s_cs[i], sn = s_sn[i];
        o1 = x1 * cs - x2 * sn;
        o2 = x2 * cs + x1 * sn;
      }
      if (head < n_q) {
        __nv_bfloat16* q = q_out + ((int64_t)b * n_q + head) * D;
        q[i] = __float2bfloat16_rn(o1);
        q[i + 64] = __float2bfloat16_rn(o2);
      } else {
        const int kv = is_v ? 1 : 0;
        const int kvh = is_v ? head - n_q - n_kv : head - n_q;
        const int64_t row = (((int64_t)(layer * 2 + kv) * n_pages + page) * n_kv + kvh) * page_size + slot;
        __nv_bfloat16* dst = kv_cache + row * D;
        dst[i] = __float2bfloat16_rn(o1);
        dst[i + 64] = __float2bfloat16_rn(o2);
      }
    }
  }
}

// ---- split-K reduce + SiLU(gate) * up ------------------------------------------------------
// part [n_split, B, 2I] (gate columns first, as in the fused gate_up weight) -> act [B, I] bf16; 4 columns/thread
__global__ void __launch_bounds__(256) silu_mul_kernel(const float* __restrict__ part, int n_split, int B, int I,
                                                      __nv_bfloat16* __restrict__ act, const void* pf_ptr,
                                                      size_t pf_bytes) {
  pdl_launch_dependents();
  l2_prefetch_range(pf_ptr, pf_bytes, ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * blockDim.x + threadIdx.x,
                    (size_t)gridDim.x * gridDim.y * blockDim.x);
  pdl_wait();
  const int b = blockIdx.y;
  const int i4 = blockIdx.x * blockDim.x + threadIdx.x;
  if (i4 * 4 >= I) return;
  float4 g = make_float4(0.f, 0.f, 0.f, 0.f), u = g;
  for (int s = 0; s < n_split; ++s) {
    const float* p = part + ((int64_t)s * B + b) * 2 * I;
    const float4 a = reinterpret_cast<const float4*>(p)[i4];
    const float4 c = reinterpret_cast<const float4*>(p + I)[i4];
    g.x += a.x; g.y += a.y; g.z += a.z; g.w += a.w;
    u.x += c.x; u.y += c.y; u.z += c.z; u.w += c.w;
  }
  auto f = [](float gg, float uu) { return (gg / (1.f + __expf(-gg))) * uu; };
  const uint32_t o0 = float_to_bf16_bits(f(g.x, u.x)), o1 = float_to_bf16_bits(f(g.y, u.y));
  const uint32_t o2 = float_to_bf16_bits(f(g.z, u.z)), o3 = float_to_bf16_bits(f(g.w, u.w));
  reinterpret_cast<uint2*>(act + (int64_t)b * I)[i4] = make_uint2(o0 | (o1 << 16), o2 | (o3 << 16));
}

// ---- sampler + in-kernel logprob capture ----------------------------------------------------
// logits [B, V] fp32 (split-K already 1 for the head).  Per token: z = logits / T;
// id = argmax(z + Gumbel noise) (== a sample from softmax(z)), or argmax(z) when greedy;
// logprob = z[id] - logsumexp(z)  (vLLM's processed_logprobs at top_p = 1, top_k = -1).
struct ArgMax { float v; int i; };
__device__ __forceinline__ ArgMax better(ArgMax a, ArgMax b) {
  // ties -> lowest index, like torch.argmax
  if (b.v > a.v || (b.v == a.v && b.i < a.i)) return b;
  return a;
}

constexpr int kSampleParts = 16;   // CTAs per row: 64 rows x 16 = 1024 CTAs keep every SM busy
constexpr int kSampleThreads = 256;
struct SamplePartial { float m, s, v, z; int i; int pad[3]; };  // z = logit/T of the best key (logprob without re-reading logits)

// phase 1: each CTA scans a contiguous 1/16 of the vocabulary of one row
__global__ void __launch_bounds__(kSampleThreads) sample_partial_kernel(const float* __restrict__ logits, int V,
                                                                       float inv_temp, int greedy, uint64_t seed,
                                                                       uint32_t step, int vocab_offset,
                                                                       SamplePartial* __restrict__ part,
                                                                       const float* __restrict__ inv_temp_rows,
                                                                       const uint8_t* __restrict__ greedy_rows) {
  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.x, pi = blockIdx.y;
  // per-sequence sampling parameters (requests of different LLM handles share an engine: train T=1, eval greedy ...)
  if (inv_temp_rows != nullptr) inv_temp = inv_temp_rows[b];
  if (greedy_rows != nullptr) greedy = greedy_rows[b];
  const float* z = logits + (int64_t)b * V;
  const int per = (V + kSampleParts - 1) / kSampleParts;
  const int lo = pi * per, hi = (lo + per < V) ? lo + per : V;
  float m = -INFINITY, s = 0.f;
  ArgMax best{-INFINITY, 0x7fffffff};
  float best_z = 0.f;
  for (int i = lo + threadIdx.x; i < hi; i += kSampleThreads) {
    const float zi = z[i] * inv_temp;
    if (zi > m) { s = s * __expf(m - zi) + 1.f; m = zi; } else { s += __expf(zi - m); }
    const int gid = vocab_offset + i;   // global vocabulary id (vocab-parallel head: this rank owns a slice)
    const float key = greedy ? zi : zi + gumbel(seed, step, (uint32_t)b, (uint32_t)gid);
    const ArgMax cand{key, gid};
    const ArgMax nb = better(best, cand);
    if (nb.i != best.i) best_z = zi;
    best = nb;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float m2 = __shfl_xor_sync(0xffffffffu, m, o), s2 = __shfl_xor_sync(0xffffffffu, s, o);
    const float mm = fmaxf(m, m2);
    s = (m == -INFINITY ? 0.f : s * __expf(m - mm)) + (m2 == -INFINITY ? 0.f : s2 * __expf(m2 - mm));
    m = mm;
    ArgMax o2{__shfl_xor_sync(0xffffffffu, best.v, o), __shfl_xor_sync(0xffffffffu, best.i, o)};
    const float z2 = __shfl_xor_sync(0xffffffffu, best_z, o);
    const ArgMax nb = better(best, o2);
    if (nb.i != best.i) best_z = z2;
    best = nb;
  }
  __shared__ float s_m[8], s_s[8], s_v[8], s_z[8];
  __shared__ int s_i[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) { s_m[warp] = m; s_s[warp] = s; s_v[warp] = best.v; s_i[warp] = best.i; s_z[warp] = best_z; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float M = -INFINITY, S = 0.f, bz = 0.f;
    ArgMax bb{-INFINITY, 0x7fffffff};
    for (int w = 0; w < kSampleThreads / 32; ++w) {
      const float mm = fmaxf(M, s_m[w]);
      S = (M == -INFINITY ? 0.f : S * __expf(M - mm)) + (s_m[w] == -INFINITY ? 0.f : s_s[w] * __expf(s_m[w] - mm));
      M = mm;
      const ArgMax nb = better(bb, ArgMax{s_v[w], s_i[w]});
      if (nb.i != bb.i) bz = s_z[w];
      bb = nb;
    }
    SamplePartial out;
    out.m = M; out.s = S; out.v = bb.v; out.z = bz; out.i = bb.i; out.pad[0] = out.pad[1] = out.pad[2] = 0;
    part[b * kSampleParts + pi] = out;
  }
}

// phase 2: merge the partials of a row — 16 per vocabulary slice, `n_groups` slices ([group][B][16]; one group
// unless the head is vocab-parallel over tensor-parallel ranks) — in a fixed order; emit id and log-probability
__global__ void sample_finalize_kernel(int B, int n_groups, const SamplePartial* __restrict__ part,
                                       int32_t* __restrict__ out_ids, float* __restrict__ out_logprobs) {
  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  float M = -INFINITY, S = 0.f, bz = 0.f;
  ArgMax bb{-INFINITY, 0x7fffffff};
  for (int g = 0; g < n_groups; ++g)
    for (int k = 0; k < kSampleParts; ++k) {
      const SamplePartial q = part[((int64_t)g * B + b) * kSampleParts + k];
      const float mm = fmaxf(M, q.m);
      S = (M == -INFINITY ? 0.f : S * __expf(M - mm)) + (q.m == -INFINITY ? 0.f : q.s * __expf(q.m - mm));
      M = mm;
      const ArgMax nb = better(bb, ArgMax{q.v, q.i});
      if (nb.i != bb.i) bz = q.z;
      bb = nb;
    }
  out_ids[b] = bb.i;
  out_logprobs[b] = bz - (M + logf(S));
}

// ---- tensor-parallel synchronisation through peer memory (no NCCL on the token path) -----------------------------
// signal: after this rank's P2P stores (previous kernels of the stream) bump a counter in the PEER's memory.
__global__ void tp_signal_kernel(unsigned long long* peer_flag) {
  pdl_launch_dependents();
  pdl_wait();
  if (threadIdx.x == 0) {
    __threadfence_system();
    atomicAdd_system(peer_flag, 1ull);
  }
}
// wait: spin until the local counter (bumped by the peer) reaches epoch * per_step + k, i.e. the peer has issued
// its k-th signal of this token step.  `epoch` counts completed steps on this rank (tp_epoch_kernel).
__global__ void tp_wait_kernel(const unsigned long long* flag, const unsigned long long* epoch, int per_step, int k) {
  pdl_launch_dependents();
  pdl_wait();
  if (threadIdx.x == 0) {
    const unsigned long long want = (*epoch) * (unsigned long long)per_step + (unsigned long long)k;
    const long long t0 = clock64();
    while (true) {
      unsigned long long v;
      asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(flag) : "memory");
      if (v >= want) break;
      if (clock64() - t0 > 20000000000LL) asm volatile("trap;");  // ~10 s: the peer died
    }
  }
}
__global__ void tp_epoch_kernel(unsigned long long* epoch) {
  pdl_launch_dependents();
  pdl_wait();
  if (threadIdx.x == 0) *epoch += 1ull;
}

// ---- advance the per-sequence state after a step (device-side, no host round trip) -------------
// Slot b just processed the token at position pos.  While the next position is still inside the
// prompt the sample is discarded and the next prompt token is fed (prefill-by-decode); afterwards
// the sampled id / logprob are appended to the slot's output ring and become the next input.
__global__ void advance_kernel(prl_engine_state st) {
  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= st.B) return;
  if (!st.active[b]) return;
  const int pos = st.positions[b];
  const int next = pos + 1;
  if (next < st.prompt_len[b]) {
    st.tokens[b] = st.prompt_buf[(int64_t)b * st.prompt_stride + next];
  } else {
    const int n = st.gen_count[b];
    const int id = st.sampled[b];
    st.out_ids[(int64_t)b * st.out_stride + n] = id;
    st.out_logprobs[(int64_t)b * st.out_stride + n] = st.sampled_logprobs[b];
    st.gen_count[b] = n + 1;
    st.tokens[b] = id;
    const bool ignore = st.ignore_eos || (st.ignore_eos_rows != nullptr && st.ignore_eos_rows[b]);
    const bool eos = (id == st.eos_id) && !ignore;
    if (eos || n + 1 >= st.max_new[b]) {
      st.finished[b] = eos ? 1 : 2;  // 1 = stop, 2 = length
      st.active[b] = 0;
      st.seq_lens[b] = 0;            // the slot stops reading its KV
      st.positions[b] = 0;
      return;
    }
  }
  st.positions[b] = next;
  st.seq_lens[b] = next + 1;
}

}  // namespace
}  // namespace prl

using namespace prl;

extern "C" int prl_embed_rmsnorm(const int32_t* tokens, const void* embed, const void* gamma, float eps, int32_t B,
                                 int32_t H, int32_t vocab, float* h, void* x_bf16, prl_stream_t st) {
  PRL_CHECK_ARG(tokens && embed && gamma && h && x_bf16 && B >= 1 && H >= 1, "prl_embed_rmsnorm: bad argument");
  PRL_CUDA(launch_pdl(embed_rmsnorm_kernel, dim3(B), dim3(256), 0, (cudaStream_t)st, tokens, (const __nv_bfloat16*)embed,
                      (const __nv_bfloat16*)gamma, eps, (int)H, (int)vocab, h, (__nv_bfloat16*)x_bf16));
  PRL_LAUNCH_CHECK();
  return PRL_OK;
}

extern "C" int prl_residual_rmsnorm(const float* partials, int32_t n_split, int32_t B, int32_t H, const void* gamma,
                                    float eps, float* h, void* x_bf16, const void* l2_prefetch, size_t l2_prefetch_bytes,
                                    prl_stream_t st) {
  PRL_CHECK_ARG(partials && gamma && h && x_bf16 && B >= 1 && H >= 4 && n_split >= 0, "prl_residual_rmsnorm: bad argument");
  PRL_CHECK_ARG(H % 4 == 0, "prl_residual_rmsnorm: hidden size must be a multiple of 4 (got %d)", H);
  PRL_CHECK_ARG(H * 4 <= 96 * 1024, "prl_residual_rmsnorm: hidden size too large for the row buffer");
  static SmemAttr smem_attr = {};
  PRL_CUDA(ensure_smem(residual_rmsnorm_kernel, 96 * 1024, smem_attr));
  int threads = ((H / 4 + 31) / 32) * 32;
  if (threads > 1024) threads = 1024;
  const size_t smem = (H / 4 > threads) ? (size_t)H * 4 : 0;
  PRL_CUDA(launch_pdl(residual_rmsnorm_kernel, dim3(B), dim3(threads), smem, (cudaStream_t)st, partials, (int)n_split,
                      (int)B, (int)H, (const __nv_bfloat16*)gamma, eps, h, (__nv_bfloat16*)x_bf16, l2_prefetch,
                      l2_prefetch_bytes));
  PRL_LAUNCH_CHECK();
  return PRL_OK;
}

extern "C" int prl_qkv_rope_cache(const float* partials, int32_t n_split, int32_t B, const void* bias, int32_t n_q,
                                  int32_t n_kv, int32_t head_dim, const int32_t* positions,
                                  const int32_t* block_table, int32_t max_blocks, const int32_t* row_slot,
                                  const float* inv_freq, void* q_out,
                                  void* kv_cache, int64_t n_pages, int32_t layer, int32_t page_size,
                                  const void* l2_prefetch, size_t l2_prefetch_bytes, prl_stream_t st) {
  PRL_CHECK_ARG(partials && positions && block_table && q_out && kv_cache && inv_freq, "prl_qkv_rope_cache: NULL argument");
  PRL_CHECK_ARG(head_dim == 128, "prl_qkv_rope_cache: head_dim must be 128 (got %d)", head_dim);
  PRL_CHECK_ARG(B >= 1 && n_q >= 1 && n_kv >= 1 && page_size >= 1 && max_blocks >= 1, "prl_qkv_rope_cache: bad shape");
  if (B > 128) {  // prefill chunk: row-walking variant (same results, ~4x less time at 1024 rows)
    const unsigned blocks = (unsigned)(B < 148 * 8 ? B : 148 * 8);
    PRL_CUDA(launch_pdl(qkv_rope_cache_rows_kernel, dim3(blocks), dim3(256), 0, (cudaStream_t)st, partials, (int)n_split,
                        (int)B, (const __nv_bfloat16*)bias, (int)n_q, (int)n_kv, positions, block_table, (int)max_blocks,
                        row_slot, inv_freq, (__nv_bfloat16*)q_out, (__nv_bfloat16*)kv_cache, n_pages, (int)layer,
                        (int)page_size));
    PRL_LAUNCH_CHECK();
    return PRL_OK;
  }
  dim3 grid((unsigned)B, (unsigned)(n_q + 2 * n_kv));
  PRL_CUDA(launch_pdl(qkv_rope_cache_kernel, grid, dim3(64), 0, (cudaStream_t)st, partials, (int)n_split, (int)B,
                      (const __nv_bfloat16*)bias, (int)n_q, (int)n_kv, positions, block_table, (int)max_blocks, row_slot,
                      inv_freq, (__nv_bfloat16*)q_out, (__nv_bfloat16*)kv_cache, n_pages, (int)layer, (int)page_size,
                      l2_prefetch, l2_prefetch_bytes));
  PRL_LAUNCH_CHECK();
  return PRL_OK;
}

extern "C" int prl_silu_mul(const float* partials, int32_t n_split, int32_t B, int32_t I, void* act_bf16,
                            const void* l2_prefetch, size_t l2_prefetch_bytes, prl_stream_t st) {
  PRL_CHECK_ARG(partials && act_bf16 && B >= 1 && I >= 4 && n_split >= 1, "prl_silu_mul: bad argument");
  PRL_CHECK_ARG(I % 4 == 0, "prl_silu_mul: intermediate size must be a multiple of 4 (got %d)", I);
  dim3 grid((unsigned)((I / 4 + 255) / 256), (unsigned)B);
  PRL_CUDA(launch_pdl(silu_mul_kernel, grid, dim3(256), 0, (cudaStream_t)st, partials, (int)n_split, (int)B, (int)I,
                      (__nv_bfloat16*)act_bf16, l2_prefetch, l2_prefetch_bytes));
  PRL_LAUNCH_CHECK();
  return PRL_OK;
}

extern "C" size_t prl_sample_workspace_bytes(int32_t B) {
  return (size_t)(B < 1 ? 1 : B) * kSampleParts * sizeof(SamplePartial);
}

// phase 1 only: per-row partials of logits[:, vocab slice] into `partials` ([B][16]); ids are global (offset added)
extern "C" int prl_sample_partials(const float* logits, int32_t B, int32_t V, float temperature, int32_t greedy,
                                   uint64_t seed, uint32_t step, int32_t vocab_offset, void* partials,
                                   prl_stream_t st) {
  PRL_CHECK_ARG(logits && partials && B >= 1 && V >= 1, "prl_sample_partials: bad argument");
  PRL_CHECK_ARG(temperature > 0.f, "prl_sample_partials: temperature must be > 0 (use greedy=1 for argmax)");
  dim3 grid((unsigned)B, kSampleParts);
  PRL_CUDA(launch_pdl(sample_partial_kernel, grid, dim3(kSampleThreads), 0, (cudaStream_t)st, logits, (int)V,
                      1.f / temperature, (int)greedy, seed, step, (int)vocab_offset, (SamplePartial*)partials,
                      (const float*)nullptr, (const uint8_t*)nullptr));
  PRL_LAUNCH_CHECK();
  return PRL_OK;
}

extern "C" int prl_sample_logprob_rows(const float* logits, int32_t B, int32_t V, const float* inv_temperature_rows,
                                       const uint8_t* greedy_rows, uint64_t seed, uint32_t step, int32_t* out_ids,
                                       float* out_logprobs, void* workspace, size_t workspace_bytes, prl_stream_t st) {
  PRL_CHECK_ARG(logits && inv_temperature_rows && greedy_rows && out_ids && out_logprobs && B >= 1 && V >= 1,
                "prl_sample_logprob_rows: bad argument");
  PRL_CHECK_ARG(workspace && workspace_bytes >= prl_sample_workspace_bytes(B), "prl_sample_logprob_rows: workspace too small");
  dim3 grid((unsigned)B, kSampleParts);
  PRL_CUDA(launch_pdl(sample_partial_kernel, grid, dim3(kSampleThreads), 0, (cudaStream_t)st, logits, (int)V, 1.f, 0, seed,
                      step, 0, (SamplePartial*)workspace, inv_temperature_rows, greedy_rows));
  PRL_LAUNCH_CHECK();
  return prl_sample_finalize(workspace, B, 1, out_ids, out_logprobs, st);
}

// phase 2 only: merge n_groups x 16 partials per row ([group][B][16]) -> ids, logprobs
extern "C" int prl_sample_finalize(const void* partials, int32_t B, int32_t n_groups, int32_t* out_ids,
                                   float* out_logprobs, prl_stream_t st) {
  PRL_CHECK_ARG(partials && out_ids && out_logprobs && B >= 1 && n_groups >= 1, "prl_sample_finalize: bad argument");
  PRL_CUDA(launch_pdl(sample_finalize_kernel, dim3((B + 63) / 64), dim3(64), 0, (cudaStream_t)st, (int)B, (int)n_groups,
                      (const SamplePartial*)partials, out_ids, out_logprobs));
  PRL_LAUNCH_CHECK();
  return PRL_OK;
}

extern "C" int prl_sample_logprob(const float* logits, int32_t B, int32_t V, float temperature, int32_t greedy,
                                  uint64_t seed, uint32_t step, int32_t* out_ids, float* out_logprobs,
                                  void* workspace, size_t workspace_bytes, prl_stream_t st) {
  PRL_CHECK_ARG(workspace && workspace_bytes >= prl_sample_workspace_bytes(B), "prl_sample_logprob: workspace too small");
  int rc = prl_sample_partials(logits, B, V, temperature, greedy, seed, step, 0, workspace, st);
  if (rc) return rc;
  return prl_sample_finalize(workspace, B, 1, out_ids, out_logprobs, st);
}

extern "C" int prl_tp_signal(void* peer_flag, prl_stream_t st) {
  PRL_CHECK_ARG(peer_flag, "prl_tp_signal: NULL flag");
  PRL_CUDA(launch_pdl(tp_signal_kernel, dim3(1), dim3(32), 0, (cudaStream_t)st, (unsigned long long*)peer_flag));
  PRL_LAUNCH_CHECK();
  return PRL_OK;
}
extern "C" int prl_tp_wait(const void* flag, const void* epoch, int32_t signals_per_step, int32_t k, prl_stream_t st) {
  PRL_CHECK_ARG(flag && epoch && signals_per_step >= 1 && k >= 1 && k <= signals_per_step, "prl_tp_wait: bad argument");
  PRL_CUDA(launch_pdl(tp_wait_kernel, dim3(1), dim3(32), 0, (cudaStream_t)st, (const unsigned long long*)flag,
                      (const unsigned long long*)epoch, (int)signals_per_step, (int)k));
  PRL_LAUNCH_CHECK();
  return PRL_OK;
}
extern "C" int prl_tp_epoch(void* epoch, prl_stream_t st) {
  PRL_CHECK_ARG(epoch, "prl_tp_epoch: NULL counter");
  PRL_CUDA(launch_pdl(tp_epoch_kernel, dim3(1), dim3(32), 0, (cudaStream_t)st, (unsigned long long*)epoch));
  PRL_LAUNCH_CHECK();
  return PRL_OK;
}

extern "C" int prl_advance_state(const prl_engine_state* state, prl_stream_t st) {
  PRL_CHECK_ARG(state && state->B >= 1, "prl_advance_state: bad argument");
  PRL_CHECK_ARG(state->sampled && state->sampled_logprobs && state->tokens && state->positions && state->seq_lens &&
                    state->active && state->prompt_buf && state->prompt_len && state->out_ids &&
                    state->out_logprobs && state->gen_count && state->max_new && state->finished,
                "prl_advance_state: NULL field");
  PRL_CUDA(launch_pdl(advance_kernel, dim3((state->B + 127) / 128), dim3(128), 0, (cudaStream_t)st, *state));
  PRL_LAUNCH_CHECK();
  return PRL_OK;
}
