// tcgen05 chunked-prefill attention (hot path 1: the <= 1024-token prefill chunks vLLM interleaves with decode,
// conf/base.yaml:64,72; also prefix-shared prefill and reference-logprob scoring).
//
// One CTA per (query tile, kv head, sequence).  A query tile packs nq = 128 / R consecutive query tokens x the R
// query heads of one GQA group into the 128 rows of a UMMA tile (row = token * R + head), so every K/V page staged
// by TMA serves all R heads.  Per 128-key step:
//     S[128 x 128]  = Q K^T      tcgen05.mma, both operands K-major (k = head dim), accumulator in TMEM
//     P             = exp2(S * scale - rowmax)   four softmax warps, thread = row, straight out of TMEM;
//                                                bf16 P goes to shared memory in the 128-byte-swizzled K-major layout
//     Ot[128 x 128] = P V        tcgen05.mma, A = P (smem), B = V read AS STORED (MN-major operand: keys are rows)
// and the softmax warps fold Ot into their fp32 register accumulator with the online-softmax rescale.  S and Ot are
// double-buffered in the 512 TMEM columns: the tensor core runs Q K^T of step i+1 and P V of step i while the
// softmax warps work on step i / fold step i-1.  Warp roles: 0 = TMA producer, 1 = MMA issuer + TMEM owner,
// 2..5 = softmax / epilogue.
//
// Tensor-bound: 4 * 128 * S^2 / 2 FLOP per head (causal); K/V bytes are re-read from L2 by the other query tiles.
#include "prl_common.cuh"
#include "tc_ptx.cuh"

namespace prl {
namespace {

constexpr int kPageT = 64;
constexpr int kDT = 128;
constexpr int kKeys = 128;                 // keys per step = 2 pages
constexpr int kTile16K = 16384;            // one [128 rows x 128 B] operand tile
constexpr int kStageBytesT = 4 * kTile16K; // K lo/hi + V lo/hi
constexpr int kThreadsT = 192;

struct TcPrefillParams {
  __nv_bfloat16* out;            // [rows, n_q*128]
  const int32_t* block_table;    // [slots, max_blocks]
  const int32_t* seq_q_start;
  const int32_t* seq_q_len;
  const int32_t* seq_pos0;
  const int32_t* seq_slot;
  int max_blocks, n_q, n_kv, R, nq;
  int64_t n_pages;
  int layer;
  float scale_log2;
};

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

__global__ void __launch_bounds__(kThreadsT, 1)
attn_prefill_tc_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_kv,
                       TcPrefillParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t q_smem = base;                          // Q lo | Q hi
  const uint32_t kv_smem = base + 2 * kTile16K;          // 2 stages x (K lo | K hi | V lo | V hi)
  const uint32_t p_smem = kv_smem + 2 * kStageBytesT;    // P keys 0..63 | keys 64..127
  const uint32_t bar_base = p_smem + 2 * kTile16K;
  auto bar = [&](int i) { return bar_base + 8u * (uint32_t)i; };
  // 0 q_full | 1,2 kv_full | 3,4 kv_empty | 5,6 s_full | 7,8 s_empty | 9 p_full | 10,11 o_full | 12,13 o_empty
  const uint32_t tmem_slot = bar(14);

  const int qtile = blockIdx.x, kvh = blockIdx.y, z = blockIdx.z;
  const int q_len = p.seq_q_len[z];
  const int t0 = qtile * p.nq;
  if (t0 >= q_len) return;                               // uniform across the CTA, before any barrier / TMEM use
  const int row0 = p.seq_q_start[z] + t0;
  const int pos_first = p.seq_pos0[z] + t0;
  const int n_valid = (q_len - t0) < p.nq ? (q_len - t0) : p.nq;
  const int kv_end = pos_first + n_valid;                // keys [0, kv_end) are visible to the last query of the tile
  const int n_it = (kv_end + kKeys - 1) / kKeys;
  const int last_page = (kv_end - 1) / kPageT;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < 14; ++i) ptx::mbar_init(bar(i), 1);
    ptx::fence_barrier_init();
    ptx::fence_proxy_async();
    ptx::prefetch_tensormap(&tm_q);
    ptx::prefetch_tensormap(&tm_kv);
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_slot, 512);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      ptx::mbar_arrive_expect_tx(bar(0), (uint32_t)(2 * 128 * p.R * p.nq));
      ptx::tma_load_3d(q_smem, &tm_q, 0, kvh * p.R, row0, bar(0), ptx::kEvictFirst);
      ptx::tma_load_3d(q_smem + kTile16K, &tm_q, 64, kvh * p.R, row0, bar(0), ptx::kEvictFirst);
      const int32_t* bt = p.block_table + (int64_t)p.seq_slot[z] * p.max_blocks;
      for (int it = 0; it < n_it; ++it) {
        const int s = it & 1;
        const uint32_t ph = (uint32_t)((it >> 1) & 1);
        ptx::mbar_wait(bar(3 + s), ph ^ 1u);
        ptx::mbar_arrive_expect_tx(bar(1 + s), (uint32_t)kStageBytesT);
        const uint32_t dst = kv_smem + (uint32_t)(s * kStageBytesT);
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          int pg = 2 * it + half;
          if (pg > last_page) pg = last_page;            // the tail step re-reads the last page; its keys are masked
          const int page = bt[pg];
          const int row_k = (int)(((((int64_t)p.layer * 2 + 0) * p.n_pages + page) * p.n_kv + kvh) * kPageT);
          const int row_v = (int)(((((int64_t)p.layer * 2 + 1) * p.n_pages + page) * p.n_kv + kvh) * kPageT);
          const uint32_t off = (uint32_t)(half * 8192);
          ptx::tma_load_2d(dst + off, &tm_kv, 0, row_k, bar(1 + s), ptx::kEvictLast);
          ptx::tma_load_2d(dst + kTile16K + off, &tm_kv, 64, row_k, bar(1 + s), ptx::kEvictLast);
          ptx::tma_load_2d(dst + 2 * kTile16K + off, &tm_kv, 0, row_v, bar(1 + s), ptx::kEvictLast);
          ptx::tma_load_2d(dst + 3 * kTile16K + off, &tm_kv, 64, row_v, bar(1 + s), ptx::kEvictLast);
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      constexpr uint32_t idesc_qk = ptx::make_idesc_bf16_f32(128, kKeys);
      constexpr uint32_t idesc_pv = ptx::make_idesc_bf16_f32(128, kDT) | (1u << 16);  // B (= V) is MN-major
      auto issue_qk = [&](int j) {
        const int s = j & 1;
        const uint32_t ph = (uint32_t)((j >> 1) & 1);
        ptx::mbar_wait(bar(1 + s), ph);          // K/V of step j landed
        ptx::mbar_wait(bar(7 + s), ph ^ 1u);     // S[s] drained by the softmax warps (step j - 2)
        ptx::tc_fence_after_sync();
        const uint32_t k_addr = kv_smem + (uint32_t)(s * kStageBytesT);
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
          const uint64_t a = ptx::make_kmajor_sw128_desc(q_smem + (uint32_t)((ks >> 2) * kTile16K)) + (uint64_t)(2 * (ks & 3));
          const uint64_t b = ptx::make_kmajor_sw128_desc(k_addr + (uint32_t)((ks >> 2) * kTile16K)) + (uint64_t)(2 * (ks & 3));
          ptx::mma_bf16_ss(tmem_base + (uint32_t)(s * 128), a, b, idesc_qk, ks > 0 ? 1u : 0u);
        }
        ptx::tc_commit(bar(5 + s));
      };
      ptx::mbar_wait(bar(0), 0);
      issue_qk(0);
      for (int i = 0; i < n_it; ++i) {
        if (i + 1 < n_it) issue_qk(i + 1);
        const int s = i & 1;
        const uint32_t ph = (uint32_t)((i >> 1) & 1);
        ptx::mbar_wait(bar(9), (uint32_t)(i & 1));   // P of step i is in shared memory
        ptx::mbar_wait(bar(12 + s), ph ^ 1u);        // Ot[s] folded by the softmax warps (step i - 2)
        ptx::tc_fence_after_sync();
        const uint32_t v_addr = kv_smem + (uint32_t)(s * kStageBytesT) + 2 * kTile16K;
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
          const uint64_t a = ptx::make_kmajor_sw128_desc(p_smem + (uint32_t)((ks >> 2) * kTile16K)) + (uint64_t)(2 * (ks & 3));
          const uint64_t b = ptx::make_mnmajor_sw128_desc(v_addr, kTile16K) + (uint64_t)(128 * ks);
          ptx::mma_bf16_ss(tmem_base + (uint32_t)(256 + s * 128), a, b, idesc_pv, ks > 0 ? 1u : 0u);
        }
        ptx::tc_commit(bar(10 + s));   // Ot[s] complete (and P free again)
        ptx::tc_commit(bar(3 + s));    // K/V stage free
      }
    }
    __syncwarp();
  } else {
    // ===== softmax + epilogue: thread = one (token, head) row =====
    const int q = warp & 3;
    const int m = q * 32 + lane;
    const int qi = m / p.R, r = m - qi * p.R;
    const int qpos = pos_first + qi;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    const uint32_t p_row = p_smem + (uint32_t)(m * 128);
    float m_run = -INFINITY, l_run = 0.f, alpha_prev = 0.f;
    float o[kDT];
#pragma unroll
    for (int d = 0; d < kDT; ++d) o[d] = 0.f;

    auto fold = [&](int j) {   // o = o * alpha_j + Ot_j
      const int s = j & 1;
      ptx::mbar_wait(bar(10 + s), (uint32_t)((j >> 1) & 1));
      ptx::tc_fence_after_sync();
#pragma unroll
      for (int c0 = 0; c0 < kDT; c0 += 32) {
        uint32_t v[32];
        ptx::tmem_ld_32x32b_x32(lane_addr + (uint32_t)(256 + s * 128 + c0), v);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 32; ++e) o[c0 + e] = fmaf(o[c0 + e], alpha_prev, __uint_as_float(v[e]));
      }
      ptx::tc_fence_before_sync();
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (threadIdx.x == 64) ptx::mbar_arrive(bar(12 + s));
    };

    for (int i = 0; i < n_it; ++i) {
      const int s = i & 1;
      ptx::mbar_wait(bar(5 + s), (uint32_t)((i >> 1) & 1));
      ptx::tc_fence_after_sync();
      const int key0 = i * kKeys;
      const bool diag = key0 + kKeys - 1 > pos_first;    // some (row, key) of this step is masked
      // pass 1: row maximum
      float mx = -INFINITY;
#pragma unroll
      for (int c0 = 0; c0 < kKeys; c0 += 32) {
        uint32_t v[32];
        ptx::tmem_ld_32x32b_x32(lane_addr + (uint32_t)(s * 128 + c0), v);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 32; ++e) {
          const float sc = __uint_as_float(v[e]);
          if (!diag || key0 + c0 + e <= qpos) mx = fmaxf(mx, sc);
        }
      }
      mx *= p.scale_log2;                                // scale > 0: max commutes with the scaling
      const float m_new = fmaxf(m_run, mx);
      const float m_use = (m_new == -INFINITY) ? 0.f : m_new;
      const float alpha = ex2(m_run - m_use);            // 0 on the first step
      if (i > 0) fold(i - 1);                            // P V of the previous step (also frees the P buffer)
      // pass 2: P = exp2(S * scale - max) -> bf16 -> swizzled shared memory; row sum in fp32
      float sum = 0.f;
#pragma unroll
      for (int c0 = 0; c0 < kKeys; c0 += 32) {
        uint32_t v[32];
        ptx::tmem_ld_32x32b_x32(lane_addr + (uint32_t)(s * 128 + c0), v);
        ptx::tmem_ld_wait();
        float pv[32];
#pragma unroll
        for (int e = 0; e < 32; ++e) {
          const float pe = ex2(fmaf(__uint_as_float(v[e]), p.scale_log2, -m_use));
          pv[e] = (!diag || key0 + c0 + e <= qpos) ? pe : 0.f;
          sum += pv[e];
        }
        const uint32_t tile = p_row + (uint32_t)((c0 >> 6) * kTile16K);
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
          const uint32_t ch = (uint32_t)(((c0 & 63) + j) >> 3);
          st_shared_v4(tile + ((ch ^ (uint32_t)(m & 7)) << 4), pack2(pv[j], pv[j + 1]), pack2(pv[j + 2], pv[j + 3]),
                       pack2(pv[j + 4], pv[j + 5]), pack2(pv[j + 6], pv[j + 7]));
        }
      }
      l_run = l_run * alpha + sum;
      m_run = m_new;
      alpha_prev = alpha;
      ptx::tc_fence_before_sync();
      ptx::fence_proxy_async();                          // generic-proxy stores of P -> visible to the UMMA reads
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (threadIdx.x == 64) {
        ptx::mbar_arrive(bar(7 + s));                    // S[s] drained
        ptx::mbar_arrive(bar(9));                        // P ready
      }
    }
    fold(n_it - 1);
    if (qi < n_valid && qi < p.nq) {
      const float inv = l_run > 0.f ? 1.f / l_run : 0.f;
      __nv_bfloat16* dst = p.out + ((int64_t)(row0 + qi) * p.n_q + (kvh * p.R + r)) * kDT;
#pragma unroll
      for (int d = 0; d < kDT; d += 8) {
        uint4 u;
        u.x = pack2(o[d] * inv, o[d + 1] * inv);
        u.y = pack2(o[d + 2] * inv, o[d + 3] * inv);
        u.z = pack2(o[d + 4] * inv, o[d + 5] * inv);
        u.w = pack2(o[d + 6] * inv, o[d + 7] * inv);
        *reinterpret_cast<uint4*>(dst + d) = u;
      }
    }
  }

  ptx::tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after_sync();
    ptx::tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace
}  // namespace prl

using namespace prl;

extern "C" int prl_paged_attn_prefill_tc(const void* q, int32_t q_rows, const void* kv_cache, int64_t n_pages,
                                         int32_t n_layers, int32_t layer, const int32_t* block_table,
                                         int32_t max_blocks, const int32_t* seq_q_start, const int32_t* seq_q_len,
                                         const int32_t* seq_pos0, const int32_t* seq_slot, int32_t n_seqs,
                                         int32_t max_q_len, int32_t n_q, int32_t n_kv, int32_t head_dim,
                                         int32_t page_size, float sm_scale, void* out_bf16, prl_stream_t stream_) {
  PRL_CHECK_ARG(q && kv_cache && block_table && seq_q_start && seq_q_len && seq_pos0 && seq_slot && out_bf16,
                "prl_paged_attn_prefill_tc: NULL argument");
  PRL_CHECK_ARG(head_dim == kDT && page_size == kPageT, "prl_paged_attn_prefill_tc: head_dim must be 128 and page_size 64");
  PRL_CHECK_ARG(q_rows >= 1 && n_seqs >= 1 && max_q_len >= 1 && n_kv >= 1 && n_q % n_kv == 0 && n_q / n_kv <= 128,
                "prl_paged_attn_prefill_tc: bad shape");
  PRL_CHECK_ARG(layer >= 0 && layer < n_layers, "prl_paged_attn_prefill_tc: bad layer");
  const int64_t total_rows = (int64_t)n_layers * 2 * n_pages * n_kv * kPageT;
  PRL_CHECK_ARG(total_rows < (1ll << 31), "prl_paged_attn_prefill_tc: KV cache too large for 32-bit TMA row coordinates");
  TcPrefillParams p;
  p.out = (__nv_bfloat16*)out_bf16;
  p.block_table = block_table; p.seq_q_start = seq_q_start; p.seq_q_len = seq_q_len; p.seq_pos0 = seq_pos0;
  p.seq_slot = seq_slot; p.max_blocks = max_blocks; p.n_q = n_q; p.n_kv = n_kv; p.R = n_q / n_kv;
  p.nq = 128 / p.R;
  p.n_pages = n_pages; p.layer = layer; p.scale_log2 = sm_scale * 1.4426950408889634f;
  CUtensorMap tq, tkv;
  int rc = make_tmap_2d_bf16(&tkv, kv_cache, kDT, (uint64_t)total_rows, kDT * 2, 64, kPageT);
  if (rc) return rc;
  rc = make_tmap_3d_bf16(&tq, q, kDT, (uint64_t)n_q, (uint64_t)q_rows, kDT * 2, (uint64_t)n_q * kDT * 2, 64, (uint32_t)p.R,
                         (uint32_t)p.nq);
  if (rc) return rc;
  const int smem = 2 * kTile16K + 2 * kStageBytesT + 2 * kTile16K + 1024 + 8 * 16 + 16;
  static bool configured = false;
  if (!configured) {
    PRL_CUDA(cudaFuncSetAttribute(attn_prefill_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  dim3 grid((unsigned)((max_q_len + p.nq - 1) / p.nq), (unsigned)n_kv, (unsigned)n_seqs);
  attn_prefill_tc_kernel<<<grid, kThreadsT, (size_t)smem, (cudaStream_t)stream_>>>(tq, tkv, p);
  PRL_LAUNCH_CHECK();
  return PRL_OK;
}
