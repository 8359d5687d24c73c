// tcgen05 backward of the learner's block-diagonal causal attention over one packed row (hot path 2: what the
// reference gets from flash-attn varlen through HF, pipelinerl/finetune/rl/__init__.py:204 + backward
// finetune_loop.py:716-725; conf/finetune/base.yaml:12-13,64).  Forward: attn_tc.cu (prl_attn_varlen_fwd).
//
// Rows of every UMMA tile pack the R query heads of one GQA group: row = token * R + head.  With that packing the
// contraction over query rows in dK = dS^T Q and dV = P^T dO sums over the R heads of the group inside the tensor core,
// in a fixed order -- no atomics, no cross-head reduction pass, bitwise reproducible.
//
//     P  = exp2(S * scale_log2 - lse)         S = Q K^T        (lse saved by the forward, log2 domain)
//     dP = dO V^T                              delta = rowsum(dO o O)
//     dS = P o (dP - delta)
//     dV = P^T dO        dK = scale * dS^T Q        dQ = scale * dS K
//
// Two kernels, each deterministic and each keeping every accumulator in TMEM:
//   * attn_bwd_dkdv_kernel  (K/V stationary): one CTA per (128-key tile, kv head, sequence); streams 64-row query
//     sub-tiles (Q, dO) through a 3-slot TMA ring.  It works on the TRANSPOSED scores, S^T = K Q^T and dP^T = V dO^T
//     (UMMA 128 x 64, all operands K-major), so a softmax thread owns a KEY and writes P^T / dS^T rows straight into
//     the K-major shared-memory layout that  dV += P^T dO  and  dK += dS^T Q  consume as operand A, while dO / Q are
//     read AS STORED as MN-major operand B.  TMEM: S^T, dP^T double-buffered (4 x 64 columns) + dV + dK (2 x 128).
//   * attn_bwd_dq_kernel    (Q stationary): the forward's structure (cluster pair of adjacent query tiles, K/V pages
//     TMA-multicast to both CTAs, 128 keys per step) with  dP = dO V^T  as a third MMA and  dQ += dS K  in place of
//     P V (K read as stored, MN-major).  TMEM: S double-buffered + dP + dQ.
// Both: warp 0 = TMA producer, warp 1 = MMA issuer + TMEM owner (the WHOLE warp runs converged and elects one lane per batch of
// UMMAs: issued from `if (lane == 0)` every UMMA costs ~180 cycles of R2UR + ELECT waterfall, tc_ptx.cuh: elect_one), the other
// warps = softmax: 8 in lock-step in generations 1-3, 16 that never meet in generation 4 (the default; further down).
// All four generations produce the same bits.
//
// Tensor-bound (profiles/r2_attention.md: the MMA warps sit in UMMA issue back-pressure).  Per (128 query rows x 128 keys)
// pair: 4 + 3 UMMA products of 4.2 MFLOP against 5 for the atomics-based single-kernel formulation; 2 x 16 K exp2.
#include <stdlib.h>
#include "prl_common.cuh"
#include "tc_ptx.cuh"

namespace prl {
namespace {

constexpr int kD = 128;
constexpr int kT16 = 16384;   // [128 rows x 128 B] operand tile
constexpr int kT8 = 8192;     // [64 rows x 128 B]
constexpr int kThreadsB = 320;

struct BwdParams {
  const float* lse;              // [T, n_q]
  const float* delta;            // [T, n_q]
  __nv_bfloat16* dqkv;           // [T, dqkv_stride]: dQ | dK | dV in the layout of qkv
  int64_t dqkv_stride;
  const int32_t* seg_start;
  const int32_t* seg_len;
  int n_q, n_kv, R;
  int nq;                        // query tokens per tile: 128 / R (dq kernel) or 64 / R (dkdv kernel)
  int col_k, col_v;              // element column of K / V head 0 inside a qkv row
  float scale_log2, sm_scale;
  const __nv_bfloat16* q_base;   // qkv (query heads first) and d_out, for the kernels that stage rows themselves
  const __nv_bfloat16* do_base;
  int64_t q_stride, do_stride;
  long long* timing;             // measurement only (prl_attn_debug_bwd_timing): per-phase cycle sums of one CTA, else NULL
  // sequence-parallel form (generation 4 only; NULL / dqkv values otherwise): the queries are a slice of their sequences
  const int32_t* seg_pos0;       // position of a segment's first LOCAL query inside its sequence
  const int32_t* seg_kv_start;   // row of the sequence's first key in the K / V matrix
  __nv_bfloat16* dkv;            // where dK / dV rows go: [kv rows, dkv_stride], columns dkv_col_k / dkv_col_v (+ head * 128)
  int64_t dkv_stride;
  int dkv_col_k, dkv_col_v;
};

__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t pk2(float a, float b) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}
__device__ __forceinline__ void sts_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void softmax_bar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

// =====================================================================================================
// delta[t, h] = sum_d dO[t, h, d] * O[t, h, d]      one warp per (token, head)
// =====================================================================================================
__global__ void attn_delta_kernel(const __nv_bfloat16* __restrict__ o, const __nv_bfloat16* __restrict__ d_o,
                                  int64_t n_rows /* T * n_q */, float* __restrict__ delta) {
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= n_rows) return;
  const int lane = threadIdx.x & 31;
  const uint2 a = ld_stream_u2(reinterpret_cast<const uint2*>(o + row * kD) + lane);
  const uint2 b = ld_stream_u2(reinterpret_cast<const uint2*>(d_o + row * kD) + lane);
  float s = bf16_bits_to_float(a.x & 0xFFFFu) * bf16_bits_to_float(b.x & 0xFFFFu);
  s = fmaf(bf16_bits_to_float(a.x >> 16), bf16_bits_to_float(b.x >> 16), s);
  s = fmaf(bf16_bits_to_float(a.y & 0xFFFFu), bf16_bits_to_float(b.y & 0xFFFFu), s);
  s = fmaf(bf16_bits_to_float(a.y >> 16), bf16_bits_to_float(b.y >> 16), s);
  s = warp_sum(s);
  if (lane == 0) delta[row] = s;
}

// =====================================================================================================
// dK, dV: K/V-stationary
// =====================================================================================================
constexpr int kQStages = 3;
constexpr int kQSlot = 4 * kT8;    // Q lo | Q hi | dO lo | dO hi   (64 query rows each)
constexpr int kPdsSlot = 2 * kT16; // P^T | dS^T   ([128 keys x 64 query rows] each)
constexpr int kKvBytes = 4 * kT16; // K lo | K hi | V lo | V hi
constexpr int kMetaBytes = 2 * 3 * 64 * 4;   // [2 buffers][lse | delta | qpos][64 columns]
constexpr int kSmemDkdv = 1024 + kKvBytes + kQStages * kQSlot + 2 * kPdsSlot + kMetaBytes + 8 * 18 + 16;
static_assert(kSmemDkdv <= 232448, "dkdv kernel exceeds the 227 KB shared-memory limit");

template <bool kTS>   // kTS: P^T / dS^T reach the tensor core THROUGH TMEM (A operand in tensor memory, written in place of S^T / dP^T)
__global__ void __launch_bounds__(kThreadsB, 1)
attn_bwd_dkdv_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_do,
                     const __grid_constant__ CUtensorMap tm_kv, BwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t kv_smem = base;                                   // K lo | K hi | V lo | V hi
  const uint32_t q_ring = base + kKvBytes;
  const uint32_t pds_smem = q_ring + kQStages * kQSlot;
  const uint32_t meta_smem = pds_smem + 2 * kPdsSlot;
  const uint32_t bar_base = meta_smem + kMetaBytes;
  auto bar = [&](int i) { return bar_base + 8u * (uint32_t)i; };
  // 0 kv_full | 1..3 q_full | 4..6 q_empty | 7,8 sdp_full | 9,10 sdp_empty | 11,12 pds_full | 13,14 pds_empty | 15 acc_done
  const uint32_t tmem_slot = bar(16);
  float* meta = reinterpret_cast<float*>(smem_raw + (meta_smem - ptx::smem_u32(smem_raw)));

  const int jt = blockIdx.x, kvh = blockIdx.y, z = blockIdx.z;
  const int q_len = ptx::warp_uniform(p.seg_len[z]);
  const int key0 = jt * 128;
  if (key0 >= q_len) return;                       // before any barrier / TMEM use
  const int seg0 = ptx::warp_uniform(p.seg_start[z]);
  const int nqa = p.nq;
  const int u_first = key0 / nqa;                  // first 64-row query sub-tile holding a token >= key0
  const int u_end = (q_len + nqa - 1) / nqa;
  const int n_it = u_end - u_first;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // query rows R * nqa .. 63 of a sub-tile are never written by TMA: they must read as zeros (they are contracted over)
  {
    const uint32_t n16 = (uint32_t)(kQStages * kQSlot) / 16;
    for (uint32_t i = threadIdx.x; i < n16; i += kThreadsB) sts_v4(q_ring + i * 16, 0u, 0u, 0u, 0u);
    ptx::fence_proxy_async();
  }
  if (threadIdx.x == 0) {
    for (int i = 0; i < 16; ++i) ptx::mbar_init(bar(i), 1);
    ptx::fence_barrier_init();
    ptx::fence_proxy_async();
    ptx::prefetch_tensormap(&tm_q);
    ptx::prefetch_tensormap(&tm_do);
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
  tmem_base = ptx::warp_uniform(tmem_base);

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      ptx::mbar_arrive_expect_tx(bar(0), (uint32_t)kKvBytes);
      const int krow = seg0 + key0;
#pragma unroll
      for (int kv = 0; kv < 2; ++kv) {
        const int c0 = (kv ? p.col_v : p.col_k) + kvh * kD;
#pragma unroll
        for (int half = 0; half < 2; ++half) {       // d 0..63 | d 64..127
          const uint32_t dst = kv_smem + (uint32_t)((kv * 2 + half) * kT16);
          ptx::tma_load_2d(dst, &tm_kv, c0 + 64 * half, krow, bar(0), ptx::kEvictFirst);
          ptx::tma_load_2d(dst + kT8, &tm_kv, c0 + 64 * half, krow + 64, bar(0), ptx::kEvictFirst);
        }
      }
      const uint32_t q_bytes = (uint32_t)(4 * 128 * p.R * nqa);
      for (int it = 0; it < n_it; ++it) {
        const int st = it % kQStages;
        const uint32_t ph = (uint32_t)((it / kQStages) & 1);
        ptx::mbar_wait(bar(4 + st), ph ^ 1u);
        ptx::mbar_arrive_expect_tx(bar(1 + st), q_bytes);
        const int row = seg0 + (u_first + it) * nqa;
        const uint32_t dst = q_ring + (uint32_t)(st * kQSlot);
        ptx::tma_load_3d(dst, &tm_q, 0, kvh * p.R, row, bar(1 + st), ptx::kEvictLast);
        ptx::tma_load_3d(dst + kT8, &tm_q, 64, kvh * p.R, row, bar(1 + st), ptx::kEvictLast);
        ptx::tma_load_3d(dst + 2 * kT8, &tm_do, 0, kvh * p.R, row, bar(1 + st), ptx::kEvictLast);
        ptx::tma_load_3d(dst + 3 * kT8, &tm_do, 64, kvh * p.R, row, bar(1 + st), ptx::kEvictLast);
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    {   // the whole warp, converged: one elected lane issues each tcgen05 instruction (ptx::elect_one)
      constexpr uint32_t idesc_t = ptx::make_idesc_bf16_f32(128, 64);                 // S^T, dP^T: both operands K-major
      constexpr uint32_t idesc_acc = ptx::make_idesc_bf16_f32(128, kD) | (1u << 16);  // dV, dK: B (dO / Q) MN-major
      auto issue_sdp = [&](int it) {
        const int st = it % kQStages, s = it & 1;
        ptx::mbar_wait(bar(1 + st), (uint32_t)((it / kQStages) & 1));   // Q, dO of this sub-tile landed
        // kTS: S^T[s] / dP^T[s] hold P^T / dS^T of step it - 2 until its dV / dK UMMAs, issued earlier by this thread, have
        // read them -- UMMAs of one thread execute in issue order, no barrier needed
        if (!kTS) ptx::mbar_wait(bar(9 + s), (uint32_t)(((it >> 1) & 1) ^ 1));    // S^T[s], dP^T[s] read by the softmax warps
        ptx::tc_fence_after_sync();
        const uint32_t q_addr = q_ring + (uint32_t)(st * kQSlot);
        if (ptx::elect_one()) {
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) {
            const uint64_t a = ptx::make_kmajor_sw128_desc(kv_smem + (uint32_t)((ks >> 2) * kT16)) + (uint64_t)(2 * (ks & 3));
            const uint64_t b = ptx::make_kmajor_sw128_desc(q_addr + (uint32_t)((ks >> 2) * kT8)) + (uint64_t)(2 * (ks & 3));
            ptx::mma_bf16_ss(tmem_base + (uint32_t)(s * 64), a, b, idesc_t, ks > 0 ? 1u : 0u);
          }
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) {
            const uint64_t a = ptx::make_kmajor_sw128_desc(kv_smem + (uint32_t)(2 * kT16 + (ks >> 2) * kT16)) + (uint64_t)(2 * (ks & 3));
            const uint64_t b = ptx::make_kmajor_sw128_desc(q_addr + (uint32_t)(2 * kT8 + (ks >> 2) * kT8)) + (uint64_t)(2 * (ks & 3));
            ptx::mma_bf16_ss(tmem_base + (uint32_t)(128 + s * 64), a, b, idesc_t, ks > 0 ? 1u : 0u);
          }
          ptx::tc_commit(bar(7 + s));
        }
      };
      ptx::mbar_wait(bar(0), 0);
      issue_sdp(0);
      for (int it = 0; it < n_it; ++it) {
        if (it + 1 < n_it) issue_sdp(it + 1);
        const int st = it % kQStages, s = it & 1;
        ptx::mbar_wait(bar(11 + s), (uint32_t)((it >> 1) & 1));         // P^T[s], dS^T[s] are in shared memory
        ptx::tc_fence_after_sync();
        const uint32_t q_addr = q_ring + (uint32_t)(st * kQSlot);
        const uint32_t pt_addr = pds_smem + (uint32_t)(s * kPdsSlot);
        if (ptx::elect_one()) {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {     // dV += P^T dO      (contraction over the 64 query rows)
            const uint64_t b = ptx::make_mnmajor_sw128_desc(q_addr + 2 * kT8, kT8) + (uint64_t)(128 * ks);
            if (kTS) ptx::mma_bf16_ts(tmem_base + 256u, tmem_base + (uint32_t)(s * 64 + ks * 8), b, idesc_acc, (it > 0 || ks > 0) ? 1u : 0u);
            else ptx::mma_bf16_ss(tmem_base + 256u, ptx::make_kmajor_sw128_desc(pt_addr) + (uint64_t)(2 * ks), b, idesc_acc,
                                  (it > 0 || ks > 0) ? 1u : 0u);
          }
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {     // dK += dS^T Q
            const uint64_t b = ptx::make_mnmajor_sw128_desc(q_addr, kT8) + (uint64_t)(128 * ks);
            if (kTS) ptx::mma_bf16_ts(tmem_base + 384u, tmem_base + (uint32_t)(128 + s * 64 + ks * 8), b, idesc_acc, (it > 0 || ks > 0) ? 1u : 0u);
            else ptx::mma_bf16_ss(tmem_base + 384u, ptx::make_kmajor_sw128_desc(pt_addr + kT16) + (uint64_t)(2 * ks), b, idesc_acc,
                                  (it > 0 || ks > 0) ? 1u : 0u);
          }
          ptx::tc_commit(bar(13 + s));     // P^T[s] / dS^T[s] may be rewritten
          ptx::tc_commit(bar(4 + st));     // Q / dO slot may be refilled
        }
      }
      if (ptx::elect_one()) {
        ptx::tc_commit(bar(15));
      }
    }
    __syncwarp();
  } else {
    // ===== softmax warps: a thread owns one KEY (TMEM lane) and half of the 64 query-row columns =====
    const int q = warp & 3;
    const int h = (warp - 2) >> 2;
    const int m = q * 32 + lane;                 // key row inside the tile
    const int kpos = key0 + m;
    const int ts = threadIdx.x - 64;             // 0..255
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    const int c0 = h * 32;
    // per-column metadata of a sub-tile, staged one step ahead: lse (+inf on padding columns -> P = 0), delta, token
    const int mc = ts & 63, mwhich = ts >> 6;    // 0 lse | 1 delta | 2 query position | 3 idle
    const int mqi = mc / p.R, mr = mc - mqi * p.R;
    auto fetch = [&](int u) -> float {
      const int tok = u * nqa + mqi;
      const bool valid = (mqi < nqa) && (tok < q_len);
      if (mwhich == 2) return __int_as_float(valid ? tok : -1);
      if (mwhich == 3) return 0.f;
      if (!valid) return mwhich == 0 ? INFINITY : 0.f;
      const float* src = mwhich == 0 ? p.lse : p.delta;
      return __ldg(src + (int64_t)(seg0 + tok) * p.n_q + (kvh * p.R + mr));
    };
    if (mwhich < 3) meta[mwhich * 64 + mc] = fetch(u_first);

    for (int it = 0; it < n_it; ++it) {
      const int s = it & 1;
      const int u = u_first + it;
      float nxt = 0.f;
      if (it + 1 < n_it) nxt = fetch(u + 1);     // global load in flight across the whole step
      ptx::mbar_wait(bar(7 + s), (uint32_t)((it >> 1) & 1));
      ptx::tc_fence_after_sync();
      uint32_t sv[32], dv[32];
      ptx::tmem_ld_32x32b_x32(lane_addr + (uint32_t)(s * 64 + c0), sv);
      ptx::tmem_ld_wait();
      // the dP^T read is issued now and completes under the exponentials: the TMEM read port and the MUFU pipe are the
      // two longest phases of a step and must not take turns
      ptx::tmem_ld_32x32b_x32(lane_addr + (uint32_t)(128 + s * 64 + c0), dv);
      const float* mb = meta + (it & 1) * 192;
      const bool diag = u * nqa < key0 + 127;                  // some (key, query) of this step is causally masked
      float pe[32];
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        const float4 l0 = *reinterpret_cast<const float4*>(mb + c0 + j);
        pe[j + 0] = ex2f(fmaf(__uint_as_float(sv[j + 0]), p.scale_log2, -l0.x));
        pe[j + 1] = ex2f(fmaf(__uint_as_float(sv[j + 1]), p.scale_log2, -l0.y));
        pe[j + 2] = ex2f(fmaf(__uint_as_float(sv[j + 2]), p.scale_log2, -l0.z));
        pe[j + 3] = ex2f(fmaf(__uint_as_float(sv[j + 3]), p.scale_log2, -l0.w));
      }
      if (diag) {
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const int4 q0 = *reinterpret_cast<const int4*>(mb + 128 + c0 + j);
          if (kpos > q0.x) pe[j + 0] = 0.f;
          if (kpos > q0.y) pe[j + 1] = 0.f;
          if (kpos > q0.z) pe[j + 2] = 0.f;
          if (kpos > q0.w) pe[j + 3] = 0.f;
        }
      }
      ptx::tmem_ld_wait();                                     // dP^T values are in registers
      ptx::tc_fence_before_sync();
      softmax_bar();                                           // every thread holds its S^T / dP^T values
      if (kTS) {
        // P^T / dS^T rows go back into TMEM in place of S^T[s] / dP^T[s] (lane = key, one 32-bit column = two adjacent
        // query rows; this thread owns packed columns [16 h, 16 h + 16)): no shared-memory store, no A-operand read
        uint32_t pp[16], dd[16];
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const float4 d0 = *reinterpret_cast<const float4*>(mb + 64 + c0 + j);
          pp[(j >> 1)] = pk2(pe[j], pe[j + 1]);
          pp[(j >> 1) + 1] = pk2(pe[j + 2], pe[j + 3]);
          dd[(j >> 1)] = pk2(pe[j] * (__uint_as_float(dv[j]) - d0.x), pe[j + 1] * (__uint_as_float(dv[j + 1]) - d0.y));
          dd[(j >> 1) + 1] = pk2(pe[j + 2] * (__uint_as_float(dv[j + 2]) - d0.z), pe[j + 3] * (__uint_as_float(dv[j + 3]) - d0.w));
        }
        ptx::tmem_st_32x32b_x16(lane_addr + (uint32_t)(s * 64 + h * 16), pp);
        ptx::tmem_st_32x32b_x16(lane_addr + (uint32_t)(128 + s * 64 + h * 16), dd);
        ptx::tmem_st_wait();
      } else {
      if (threadIdx.x == 64) ptx::mbar_arrive(bar(9 + s));    // -> S^T[s], dP^T[s] of step it + 2 may be issued
      ptx::mbar_wait(bar(13 + s), (uint32_t)(((it >> 1) & 1) ^ 1));   // dV / dK of step it - 2 consumed P^T[s], dS^T[s]
      const uint32_t prow = pds_smem + (uint32_t)(s * kPdsSlot + m * 128);
#pragma unroll
      for (int j = 0; j < 32; j += 8) {
        const float4 d0 = *reinterpret_cast<const float4*>(mb + 64 + c0 + j);
        const float4 d1 = *reinterpret_cast<const float4*>(mb + 64 + c0 + j + 4);
        const float dl[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
        float de[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) de[e] = pe[j + e] * (__uint_as_float(dv[j + e]) - dl[e]);
        const uint32_t off = (((uint32_t)((c0 + j) >> 3) ^ (uint32_t)(m & 7)) << 4);
        sts_v4(prow + off, pk2(pe[j], pe[j + 1]), pk2(pe[j + 2], pe[j + 3]), pk2(pe[j + 4], pe[j + 5]), pk2(pe[j + 6], pe[j + 7]));
        sts_v4(prow + kT16 + off, pk2(de[0], de[1]), pk2(de[2], de[3]), pk2(de[4], de[5]), pk2(de[6], de[7]));
      }
      }
      if (it + 1 < n_it && mwhich < 3) meta[((it + 1) & 1) * 192 + mwhich * 64 + mc] = nxt;
      if (kTS) ptx::tc_fence_before_sync();
      else ptx::fence_proxy_async();                           // generic-proxy stores -> visible to the UMMA reads
      softmax_bar();
      if (threadIdx.x == 64) ptx::mbar_arrive(bar(11 + s));   // P^T[s], dS^T[s] ready
    }

    // ---- epilogue: dV, dK rows of this key ----
    ptx::mbar_wait(bar(15), 0);
    ptx::tc_fence_after_sync();
    const bool valid = kpos < q_len;
    __nv_bfloat16* drow = p.dqkv + (int64_t)(seg0 + kpos) * p.dqkv_stride + kvh * kD + h * 64;
#pragma unroll
    for (int which = 0; which < 2; ++which) {     // 0: dV (TMEM 256..383), 1: dK (384..511)
      uint32_t v0[32], v1[32];
      ptx::tmem_ld_32x32b_x32(lane_addr + (uint32_t)(256 + which * 128 + h * 64), v0);
      ptx::tmem_ld_32x32b_x32(lane_addr + (uint32_t)(256 + which * 128 + h * 64 + 32), v1);
      ptx::tmem_ld_wait();
      const float sc = which ? p.sm_scale : 1.f;
      if (valid) {
        __nv_bfloat16* dst = drow + (which ? p.col_k : p.col_v);
#pragma unroll
        for (int d = 0; d < 32; d += 8) {
          uint4 a, b;
          a.x = pk2(__uint_as_float(v0[d]) * sc, __uint_as_float(v0[d + 1]) * sc);
          a.y = pk2(__uint_as_float(v0[d + 2]) * sc, __uint_as_float(v0[d + 3]) * sc);
          a.z = pk2(__uint_as_float(v0[d + 4]) * sc, __uint_as_float(v0[d + 5]) * sc);
          a.w = pk2(__uint_as_float(v0[d + 6]) * sc, __uint_as_float(v0[d + 7]) * sc);
          b.x = pk2(__uint_as_float(v1[d]) * sc, __uint_as_float(v1[d + 1]) * sc);
          b.y = pk2(__uint_as_float(v1[d + 2]) * sc, __uint_as_float(v1[d + 3]) * sc);
          b.z = pk2(__uint_as_float(v1[d + 4]) * sc, __uint_as_float(v1[d + 5]) * sc);
          b.w = pk2(__uint_as_float(v1[d + 6]) * sc, __uint_as_float(v1[d + 7]) * sc);
          *reinterpret_cast<uint4*>(dst + d) = a;
          *reinterpret_cast<uint4*>(dst + 32 + d) = b;
        }
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

// =====================================================================================================
// dQ: Q-stationary (cluster pair of adjacent query tiles, K/V pages multicast)
// =====================================================================================================
constexpr int kStageKV = 4 * kT16;   // K lo | K hi | V lo | V hi   (128 keys)
constexpr int kSmemDq = 1024 + 4 * kT16 + 2 * kStageKV + 2 * kT16 + 8 * 20 + 16;
static_assert(kSmemDq <= 232448, "dq kernel exceeds the 227 KB shared-memory limit");

// kGen 1: every operand through shared memory.  2: dS reaches the tensor core through TMEM, written in place of dP.
// 3: additionally the STATIONARY operands Q and dO live in TMEM for the whole kernel (64 packed columns each, staged once
//    by the softmax threads straight from global memory; S single-buffered to make room): the A operands of all three
//    UMMAs of a step then cost no shared-memory bandwidth -- 288 -> 160 KB of shared-memory traffic per step.
template <int kGen>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreadsB, 1)
attn_bwd_dq_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_do,
                   const __grid_constant__ CUtensorMap tm_kv, BwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t q_smem = base;                         // Q lo | Q hi
  const uint32_t do_smem = base + 2 * kT16;             // dO lo | dO hi
  const uint32_t kv_smem = base + 4 * kT16;             // 2 stages x (K lo | K hi | V lo | V hi)
  const uint32_t ds_smem = kv_smem + 2 * kStageKV;      // dS keys 0..63 | keys 64..127
  const uint32_t bar_base = ds_smem + 2 * kT16;
  auto bar = [&](int i) { return bar_base + 8u * (uint32_t)i; };
  // 0 q_full | 1,2 k_full | 3,4 k_empty | 5,6 s_full | 7,8 s_empty | 9 dp_full | 10 dp_empty | 11 ds_full | 12 ds_empty |
  // 13,14 v_full | 15,16 v_empty | 17 dq_done
  const uint32_t tmem_slot = bar(18);
  constexpr bool kTS = kGen >= 2;
  constexpr bool kQT = kGen >= 3;
  // TMEM columns: gen 1/2: S[2] 0,128 | dP 256 | dQ 384.   gen 3: S 0 | dP 128 | dQ 256 | Q 384 | dO 448
  constexpr uint32_t cDP = kQT ? 128u : 256u, cDQ = kQT ? 256u : 384u, cQ = 384u, cDO = 448u;

  const int qtile = (int)(gridDim.x - 1 - blockIdx.x);  // heaviest (latest) query tiles first; pairs stay adjacent
  const int kvh = blockIdx.y, z = blockIdx.z;
  const uint32_t rank = ptx::cluster_ctarank();
  const int q_len = ptx::warp_uniform(p.seg_len[z]);
  if ((qtile & ~1) * p.nq >= q_len) return;             // uniform across the cluster
  const int seg0 = ptx::warp_uniform(p.seg_start[z]);
  const int t0 = qtile * p.nq;
  const int row0 = seg0 + t0;
  const int n_valid = t0 >= q_len ? 0 : ((q_len - t0) < p.nq ? (q_len - t0) : p.nq);
  const int pair_rows = ((qtile | 1) + 1) * p.nq;
  const int kv_end = pair_rows < q_len ? pair_rows : q_len;
  const int n_it = (kv_end + 127) / 128;
  const int last_page = (kv_end - 1) / 64;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < 18; ++i) ptx::mbar_init(bar(i), (i == 3 || i == 4 || i == 15 || i == 16) ? 2 : 1);
    ptx::fence_barrier_init();
    ptx::fence_proxy_async();
    ptx::prefetch_tensormap(&tm_q);
    ptx::prefetch_tensormap(&tm_do);
    ptx::prefetch_tensormap(&tm_kv);
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_slot, 512);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before_sync();
  ptx::cluster_sync();
  ptx::tc_fence_after_sync();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  tmem_base = ptx::warp_uniform(tmem_base);

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      if (!kQT) {
        ptx::mbar_arrive_expect_tx(bar(0), (uint32_t)(4 * 128 * p.R * p.nq));
        ptx::tma_load_3d(q_smem, &tm_q, 0, kvh * p.R, row0, bar(0), ptx::kEvictFirst);
        ptx::tma_load_3d(q_smem + kT16, &tm_q, 64, kvh * p.R, row0, bar(0), ptx::kEvictFirst);
        ptx::tma_load_3d(do_smem, &tm_do, 0, kvh * p.R, row0, bar(0), ptx::kEvictFirst);
        ptx::tma_load_3d(do_smem + kT16, &tm_do, 64, kvh * p.R, row0, bar(0), ptx::kEvictFirst);
      }
      // this CTA fetches ONE of the two 64-key pages of a step and multicasts it to both CTAs of the pair
      auto load_page = [&](int it, int kv) {
        const int s = it & 1;
        const uint32_t ph = (uint32_t)((it >> 1) & 1);
        const uint32_t full = bar((kv ? 13 : 1) + s), empty = bar((kv ? 15 : 3) + s);
        ptx::mbar_wait(empty, ph ^ 1u);
        ptx::mbar_arrive_expect_tx(full, (uint32_t)(2 * kT16));
        int pg = 2 * it + (int)rank;
        if (pg > last_page) pg = last_page;            // tail: re-read the last page, its keys are causally masked
        const int row = seg0 + pg * 64;
        const int c0 = (kv ? p.col_v : p.col_k) + kvh * kD;
        const uint32_t dst = kv_smem + (uint32_t)(s * kStageKV + kv * 2 * kT16) + (uint32_t)(rank * kT8);
        ptx::tma_load_2d_multicast(dst, &tm_kv, c0, row, full, 3, ptx::kEvictLast);
        ptx::tma_load_2d_multicast(dst + kT16, &tm_kv, c0 + 64, row, full, 3, ptx::kEvictLast);
      };
      for (int it = 0; it < n_it; ++it) {
        load_page(it, 0);
        load_page(it, 1);
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    {   // the whole warp, converged: one elected lane issues each tcgen05 instruction (ptx::elect_one)
      constexpr uint32_t idesc_kk = ptx::make_idesc_bf16_f32(128, 128);                 // S, dP: K-major x K-major
      constexpr uint32_t idesc_dq = ptx::make_idesc_bf16_f32(128, kD) | (1u << 16);     // dQ: B (= K) MN-major
      auto issue_s = [&](int j) {
        const int s = j & 1;
        const uint32_t ph = (uint32_t)((j >> 1) & 1);
        ptx::mbar_wait(bar(1 + s), ph);            // K of step j landed
        if (kQT) ptx::mbar_wait(bar(7), (uint32_t)((j & 1) ^ 1));   // the single S buffer was read (step j - 1)
        else ptx::mbar_wait(bar(7 + s), ph ^ 1u);                   // S[s] drained (step j - 2)
        ptx::tc_fence_after_sync();
        const uint32_t k_addr = kv_smem + (uint32_t)(s * kStageKV);
        if (ptx::elect_one()) {
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) {
            const uint64_t b = ptx::make_kmajor_sw128_desc(k_addr + (uint32_t)((ks >> 2) * kT16)) + (uint64_t)(2 * (ks & 3));
            if (kQT) ptx::mma_bf16_ts(tmem_base, tmem_base + cQ + (uint32_t)(ks * 8), b, idesc_kk, ks > 0 ? 1u : 0u);
            else ptx::mma_bf16_ss(tmem_base + (uint32_t)(s * 128),
                                  ptx::make_kmajor_sw128_desc(q_smem + (uint32_t)((ks >> 2) * kT16)) + (uint64_t)(2 * (ks & 3)), b,
                                  idesc_kk, ks > 0 ? 1u : 0u);
          }
          ptx::tc_commit(bar(kQT ? 5 : 5 + s));
        }
      };
      auto issue_dp = [&](int j) {
        const int s = j & 1;
        ptx::mbar_wait(bar(13 + s), (uint32_t)((j >> 1) & 1));   // V of step j landed
        // kTS: dP's columns hold dS of step j - 1 until its dQ UMMAs -- issued BEFORE this call -- have read it (in order)
        if (!kTS) ptx::mbar_wait(bar(10), (uint32_t)((j & 1) ^ 1));        // dP drained (step j - 1)
        ptx::tc_fence_after_sync();
        const uint32_t v_addr = kv_smem + (uint32_t)(s * kStageKV + 2 * kT16);
        if (ptx::elect_one()) {
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) {
            const uint64_t b = ptx::make_kmajor_sw128_desc(v_addr + (uint32_t)((ks >> 2) * kT16)) + (uint64_t)(2 * (ks & 3));
            if (kQT) ptx::mma_bf16_ts(tmem_base + cDP, tmem_base + cDO + (uint32_t)(ks * 8), b, idesc_kk, ks > 0 ? 1u : 0u);
            else ptx::mma_bf16_ss(tmem_base + cDP,
                                  ptx::make_kmajor_sw128_desc(do_smem + (uint32_t)((ks >> 2) * kT16)) + (uint64_t)(2 * (ks & 3)), b,
                                  idesc_kk, ks > 0 ? 1u : 0u);
          }
          ptx::tc_commit(bar(9));
          ptx::tc_commit_multicast(bar(15 + s), 3);  // V slot consumed: tell BOTH producers
        }
      };
      ptx::mbar_wait(bar(0), 0);
      issue_s(0);
      issue_dp(0);
      for (int i = 0; i < n_it; ++i) {
        if (i + 1 < n_it) {
          issue_s(i + 1);
          if (!kTS) issue_dp(i + 1);
        }
        const int s = i & 1;
        ptx::mbar_wait(bar(11), (uint32_t)(i & 1));              // dS of step i is ready
        ptx::tc_fence_after_sync();
        const uint32_t k_addr = kv_smem + (uint32_t)(s * kStageKV);
        if (ptx::elect_one()) {
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) {           // dQ += dS K     (K read as stored: keys are the contraction rows)
            const uint64_t b = ptx::make_mnmajor_sw128_desc(k_addr, kT16) + (uint64_t)(128 * ks);
            if (kTS) ptx::mma_bf16_ts(tmem_base + cDQ, tmem_base + cDP + (uint32_t)(ks * 8), b, idesc_dq, (i > 0 || ks > 0) ? 1u : 0u);
            else ptx::mma_bf16_ss(tmem_base + cDQ, ptx::make_kmajor_sw128_desc(ds_smem + (uint32_t)((ks >> 2) * kT16)) + (uint64_t)(2 * (ks & 3)),
                                  b, idesc_dq, (i > 0 || ks > 0) ? 1u : 0u);
          }
          ptx::tc_commit(bar(12));                   // dS may be rewritten
          ptx::tc_commit_multicast(bar(3 + s), 3);   // K slot consumed: tell BOTH producers
        }
        if (kTS && i + 1 < n_it) issue_dp(i + 1);  // dP of the next step goes where dS of this one was: after its dQ UMMAs
      }
      if (ptx::elect_one()) {
        ptx::tc_commit(bar(17));
      }
    }
    __syncwarp();
  } else {
    // ===== softmax warps: a PAIR of threads owns one (token, head) row; half h works on keys [64 h, 64 h + 64) =====
    const int q = warp & 3;
    const int h = (warp - 2) >> 2;
    const int m = q * 32 + lane;
    const int qi = m / p.R, r = m - qi * p.R;
    const int qpos = t0 + qi;
    const bool valid = qi < n_valid && qi < p.nq;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    const int64_t stat = (int64_t)(row0 + qi) * p.n_q + (kvh * p.R + r);
    const float lse_row = valid ? __ldg(p.lse + stat) : INFINITY;   // padding rows: P = 0
    const float delta_row = valid ? __ldg(p.delta + stat) : 0.f;
    const uint32_t ds_row = ds_smem + (uint32_t)(h * kT16 + m * 128);
    if (kQT) {
      // stage this row's half (64 head-dim elements = 32 packed columns) of Q and dO in TMEM, straight from global memory
#pragma unroll
      for (int which = 0; which < 2; ++which) {
        const __nv_bfloat16* src = which == 0
            ? p.q_base + (int64_t)(row0 + qi) * p.q_stride + (kvh * p.R + r) * kD + h * 64
            : p.do_base + (int64_t)(row0 + qi) * p.do_stride + (kvh * p.R + r) * kD + h * 64;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t w[16];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const uint4 u = valid ? ld_stream_u4(reinterpret_cast<const uint4*>(src + c * 32) + e) : make_uint4(0u, 0u, 0u, 0u);
            w[4 * e] = u.x; w[4 * e + 1] = u.y; w[4 * e + 2] = u.z; w[4 * e + 3] = u.w;
          }
          ptx::tmem_st_32x32b_x16(lane_addr + (which == 0 ? cQ : cDO) + (uint32_t)(h * 32 + c * 16), w);
        }
      }
      ptx::tmem_st_wait();
      ptx::tc_fence_before_sync();
      softmax_bar();
      if (threadIdx.x == 64) ptx::mbar_arrive(bar(0));       // Q, dO are resident -> the UMMAs may start
    }

    for (int i = 0; i < n_it; ++i) {
      const int s = i & 1;
      if (kQT) ptx::mbar_wait(bar(5), (uint32_t)(i & 1));
      else ptx::mbar_wait(bar(5 + s), (uint32_t)((i >> 1) & 1));
      ptx::tc_fence_after_sync();
      const uint32_t cS = kQT ? 0u : (uint32_t)(s * 128);
      float sv[64];
      {
        uint32_t v0[32], v1[32];
        ptx::tmem_ld_32x32b_x32(lane_addr + cS + (uint32_t)(h * 64), v0);
        ptx::tmem_ld_32x32b_x32(lane_addr + cS + (uint32_t)(h * 64 + 32), v1);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 32; ++e) { sv[e] = __uint_as_float(v0[e]); sv[32 + e] = __uint_as_float(v1[e]); }
      }
      ptx::tc_fence_before_sync();
      softmax_bar();
      if (threadIdx.x == 64) ptx::mbar_arrive(bar(kQT ? 7 : 7 + s));   // S drained -> the next Q K^T into this buffer
      const int key0 = i * 128 + h * 64;
      const bool diag = i * 128 + 127 > t0;
      // start the dP read now: it completes under the exponentials (TMEM read port and MUFU pipe overlap)
      ptx::mbar_wait(bar(9), (uint32_t)(i & 1));             // dP of step i
      ptx::tc_fence_after_sync();
      uint32_t d0[32], d1[32];
      ptx::tmem_ld_32x32b_x32(lane_addr + cDP + (uint32_t)(h * 64), d0);
      ptx::tmem_ld_32x32b_x32(lane_addr + cDP + (uint32_t)(h * 64 + 32), d1);
#pragma unroll
      for (int e = 0; e < 64; ++e) sv[e] = ex2f(fmaf(sv[e], p.scale_log2, -lse_row));
      if (diag) {
#pragma unroll
        for (int e = 0; e < 64; ++e)
          if (key0 + e > qpos) sv[e] = 0.f;
      }
      ptx::tmem_ld_wait();
#pragma unroll
      for (int e = 0; e < 32; ++e) {
        sv[e] *= (__uint_as_float(d0[e]) - delta_row);
        sv[32 + e] *= (__uint_as_float(d1[e]) - delta_row);
      }
      ptx::tc_fence_before_sync();
      softmax_bar();                                         // every thread of the row pair has its dP values
      if (kTS) {
        // dS goes back into TMEM in place of dP (lane = row; one 32-bit column = two adjacent keys; this thread owns the
        // packed columns [32 h, 32 h + 32) of the 64): the A operand of dQ += dS K costs no shared-memory traffic
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t pp[16];
#pragma unroll
          for (int e = 0; e < 16; ++e) pp[e] = pk2(sv[c * 32 + 2 * e], sv[c * 32 + 2 * e + 1]);
          ptx::tmem_st_32x32b_x16(lane_addr + cDP + (uint32_t)(h * 32 + c * 16), pp);
        }
        ptx::tmem_st_wait();
        ptx::tc_fence_before_sync();
      } else {
      if (threadIdx.x == 64) ptx::mbar_arrive(bar(10));      // dP drained -> dO V^T of step i + 1
      ptx::mbar_wait(bar(12), (uint32_t)((i & 1) ^ 1));      // dQ MMA of step i - 1 has consumed the dS buffer
#pragma unroll
      for (int j = 0; j < 64; j += 8)
        sts_v4(ds_row + (((uint32_t)(j >> 3) ^ (uint32_t)(m & 7)) << 4), pk2(sv[j], sv[j + 1]), pk2(sv[j + 2], sv[j + 3]),
               pk2(sv[j + 4], sv[j + 5]), pk2(sv[j + 6], sv[j + 7]));
      ptx::fence_proxy_async();
      }
      softmax_bar();
      if (threadIdx.x == 64) ptx::mbar_arrive(bar(11));      // dS ready -> dQ += dS K
    }

    // ---- epilogue: this thread's 64 head-dim columns of its dQ row ----
    ptx::mbar_wait(bar(17), 0);
    ptx::tc_fence_after_sync();
    uint32_t v0[32], v1[32];
    ptx::tmem_ld_32x32b_x32(lane_addr + cDQ + (uint32_t)(h * 64), v0);
    ptx::tmem_ld_32x32b_x32(lane_addr + cDQ + (uint32_t)(h * 64 + 32), v1);
    ptx::tmem_ld_wait();
    if (valid) {
      __nv_bfloat16* dst = p.dqkv + (int64_t)(row0 + qi) * p.dqkv_stride + (kvh * p.R + r) * kD + h * 64;
      const float sc = p.sm_scale;
#pragma unroll
      for (int d = 0; d < 32; d += 8) {
        uint4 a, b;
        a.x = pk2(__uint_as_float(v0[d]) * sc, __uint_as_float(v0[d + 1]) * sc);
        a.y = pk2(__uint_as_float(v0[d + 2]) * sc, __uint_as_float(v0[d + 3]) * sc);
        a.z = pk2(__uint_as_float(v0[d + 4]) * sc, __uint_as_float(v0[d + 5]) * sc);
        a.w = pk2(__uint_as_float(v0[d + 6]) * sc, __uint_as_float(v0[d + 7]) * sc);
        b.x = pk2(__uint_as_float(v1[d]) * sc, __uint_as_float(v1[d + 1]) * sc);
        b.y = pk2(__uint_as_float(v1[d + 2]) * sc, __uint_as_float(v1[d + 3]) * sc);
        b.z = pk2(__uint_as_float(v1[d + 4]) * sc, __uint_as_float(v1[d + 5]) * sc);
        b.w = pk2(__uint_as_float(v1[d + 6]) * sc, __uint_as_float(v1[d + 7]) * sc);
        *reinterpret_cast<uint4*>(dst + d) = a;
        *reinterpret_cast<uint4*>(dst + 32 + d) = b;
      }
    }
  }

  ptx::tc_fence_before_sync();
  ptx::cluster_sync();   // the partner may still multicast into this CTA's shared memory / arrive on its barriers
  if (warp == 1) {
    ptx::tc_fence_after_sync();
    ptx::tmem_dealloc(tmem_base, 512);
  }
}


// =====================================================================================================
// Generation 4 of both kernels: the softmax warps stop moving in lock-step.
//
// What generations 1-3 lose (profiles/r2_attention.md): eight softmax warps walk through  wait -> tcgen05.ld -> exp2 ->
// pack -> tcgen05.st -> signal  TOGETHER, with three 256-thread barriers per step, so the TMEM read port, the MUFU pipe
// and the tensor core take turns; and the dQ kernel does not start its exponentials before dP of the step has arrived.
// Here: 16 softmax warps (4 per TMEM lane quarter, 32 score columns per thread, 112 registers), no CTA-wide barrier inside
// the loop (a warp signals an mbarrier by itself; every thread writes its bf16 results over columns IT read, so nothing
// is shared between threads), and
//   * dK/dV: two groups of 8 warps own the two S^T / dP^T buffers and alternate steps (ping-pong), so one group's exp2 runs
//     under the other group's TMEM traffic and under both groups' UMMAs;
//   * dQ: P = exp2(S - lse) of step i is computed while the tensor core still runs  dQ(i-1)  and  dP(i) = dO V^T.
// The UMMA A operands (P^T, dS^T, dS in TMEM) are addressed per 16-row k-step, so a thread's packed output may stay inside
// its own 32 columns: k-step ks reads packed columns 32 (ks / 2) + 8 (ks % 2).
// =====================================================================================================
constexpr int kThreadsB4 = 576;
constexpr int kQStages4 = 4;
constexpr int kMeta4 = 2 * 2 * 192 * 4;      // [group][buffer][lse | delta | qpos][64 columns]
constexpr int kSmemDkdv4 = 1024 + kKvBytes + kQStages4 * kQSlot + kMeta4 + 8 * 16 + 16;
static_assert(kSmemDkdv4 <= 232448, "dkdv4 kernel exceeds the 227 KB shared-memory limit");

__device__ __forceinline__ void group_bar256(int g) {
  if (g == 0) asm volatile("bar.sync 1, 256;" ::: "memory");
  else asm volatile("bar.sync 2, 256;" ::: "memory");
}

template <bool kTimed>
__global__ void __launch_bounds__(kThreadsB4, 1)
attn_bwd_dkdv4_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_do,
                      const __grid_constant__ CUtensorMap tm_kv, BwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t kv_smem = base;                                   // K lo | K hi | V lo | V hi
  const uint32_t q_ring = base + kKvBytes;
  const uint32_t meta_smem = q_ring + kQStages4 * kQSlot;
  const uint32_t bar_base = meta_smem + kMeta4;
  auto bar = [&](int i) { return bar_base + 8u * (uint32_t)i; };
  // 0 kv_full | 1..4 q_full | 5..8 q_empty | 9,10 sdp_full | 11,12 pds_full (one arrival per warp of the group) | 13 acc_done
  const uint32_t tmem_slot = bar(14);
  float* meta = reinterpret_cast<float*>(smem_raw + (meta_smem - ptx::smem_u32(smem_raw)));

  const int jt = blockIdx.x, kvh = blockIdx.y, z = blockIdx.z;
  const int q_len = ptx::warp_uniform(p.seg_len[z]);
  const int key0 = jt * 128;
  const int pos0 = p.seg_pos0 ? ptx::warp_uniform(p.seg_pos0[z]) : 0;   // position of local query 0 (sequence-parallel slice)
  if (key0 >= pos0 + q_len) return;                // before any barrier / TMEM use
  const int seg0 = ptx::warp_uniform(p.seg_start[z]);
  const int kv0 = p.seg_kv_start ? ptx::warp_uniform(p.seg_kv_start[z]) : seg0;
  const int nqa = p.nq;
  const int u_first = (key0 > pos0 ? key0 - pos0 : 0) / nqa;   // first 64-row query sub-tile holding a position >= key0
  const int u_end = (q_len + nqa - 1) / nqa;
  const int n_it = u_end - u_first;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  {
    const uint32_t n16 = (uint32_t)(kQStages4 * kQSlot) / 16;     // rows R * nqa .. 63 of a sub-tile must read as zeros
    for (uint32_t i = threadIdx.x; i < n16; i += kThreadsB4) sts_v4(q_ring + i * 16, 0u, 0u, 0u, 0u);
    ptx::fence_proxy_async();
  }
  if (threadIdx.x == 0) {
    for (int i = 0; i < 14; ++i) ptx::mbar_init(bar(i), (i == 11 || i == 12) ? 8 : 1);
    ptx::fence_barrier_init();
    ptx::fence_proxy_async();
    ptx::prefetch_tensormap(&tm_q);
    ptx::prefetch_tensormap(&tm_do);
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
  tmem_base = ptx::warp_uniform(tmem_base);

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      ptx::mbar_arrive_expect_tx(bar(0), (uint32_t)kKvBytes);
      const int krow = kv0 + key0;
#pragma unroll
      for (int kv = 0; kv < 2; ++kv) {
        const int c0 = (kv ? p.col_v : p.col_k) + kvh * kD;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const uint32_t dst = kv_smem + (uint32_t)((kv * 2 + half) * kT16);
          ptx::tma_load_2d(dst, &tm_kv, c0 + 64 * half, krow, bar(0), ptx::kEvictFirst);
          ptx::tma_load_2d(dst + kT8, &tm_kv, c0 + 64 * half, krow + 64, bar(0), ptx::kEvictFirst);
        }
      }
      const uint32_t q_bytes = (uint32_t)(4 * 128 * p.R * nqa);
      for (int it = 0; it < n_it; ++it) {
        const int st = it % kQStages4;
        const uint32_t ph = (uint32_t)((it / kQStages4) & 1);
        ptx::mbar_wait(bar(5 + st), ph ^ 1u);
        ptx::mbar_arrive_expect_tx(bar(1 + st), q_bytes);
        const int row = seg0 + (u_first + it) * nqa;
        const uint32_t dst = q_ring + (uint32_t)(st * kQSlot);
        ptx::tma_load_3d(dst, &tm_q, 0, kvh * p.R, row, bar(1 + st), ptx::kEvictLast);
        ptx::tma_load_3d(dst + kT8, &tm_q, 64, kvh * p.R, row, bar(1 + st), ptx::kEvictLast);
        ptx::tma_load_3d(dst + 2 * kT8, &tm_do, 0, kvh * p.R, row, bar(1 + st), ptx::kEvictLast);
        ptx::tma_load_3d(dst + 3 * kT8, &tm_do, 64, kvh * p.R, row, bar(1 + st), ptx::kEvictLast);
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    {   // the whole warp, converged: one elected lane issues each tcgen05 instruction (ptx::elect_one)
      constexpr uint32_t idesc_t = ptx::make_idesc_bf16_f32(128, 64);
      constexpr uint32_t idesc_acc = ptx::make_idesc_bf16_f32(128, kD) | (1u << 16);
      long long w_q = 0;
      auto issue_sdp = [&](int it) {
        const int st = it % kQStages4, s = it & 1;
        const long long cq = clock64();
        ptx::mbar_wait(bar(1 + st), (uint32_t)((it / kQStages4) & 1));
        if (kTimed) w_q += clock64() - cq;
        // S^T[s] / dP^T[s] hold P^T / dS^T of step it - 2 until its dV / dK UMMAs, issued earlier by this thread, have read
        // them: UMMAs of one thread execute in issue order
        ptx::tc_fence_after_sync();
        const uint32_t q_addr = q_ring + (uint32_t)(st * kQSlot);
        if (ptx::elect_one()) {
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) {
            const uint64_t a = ptx::make_kmajor_sw128_desc(kv_smem + (uint32_t)((ks >> 2) * kT16)) + (uint64_t)(2 * (ks & 3));
            const uint64_t b = ptx::make_kmajor_sw128_desc(q_addr + (uint32_t)((ks >> 2) * kT8)) + (uint64_t)(2 * (ks & 3));
            ptx::mma_bf16_ss(tmem_base + (uint32_t)(s * 64), a, b, idesc_t, ks > 0 ? 1u : 0u);
          }
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) {
            const uint64_t a = ptx::make_kmajor_sw128_desc(kv_smem + (uint32_t)(2 * kT16 + (ks >> 2) * kT16)) + (uint64_t)(2 * (ks & 3));
            const uint64_t b = ptx::make_kmajor_sw128_desc(q_addr + (uint32_t)(2 * kT8 + (ks >> 2) * kT8)) + (uint64_t)(2 * (ks & 3));
            ptx::mma_bf16_ss(tmem_base + (uint32_t)(128 + s * 64), a, b, idesc_t, ks > 0 ? 1u : 0u);
          }
          ptx::tc_commit(bar(9 + s));
        }
      };
      const bool timed = kTimed && p.timing != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0;
      long long w_pds = 0;
      const long long w0 = clock64();
      ptx::mbar_wait(bar(0), 0);
      issue_sdp(0);
      for (int it = 0; it < n_it; ++it) {
        if (it + 1 < n_it) issue_sdp(it + 1);
        const int st = it % kQStages4, s = it & 1;
        const long long c0 = clock64();
        ptx::mbar_wait(bar(11 + s), (uint32_t)((it >> 1) & 1));         // P^T[s], dS^T[s] are in TMEM
        if (kTimed) w_pds += clock64() - c0;
        ptx::tc_fence_after_sync();
        const uint32_t q_addr = q_ring + (uint32_t)(st * kQSlot);
        if (ptx::elect_one()) {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {     // dV += P^T dO      (contraction over the 64 query rows, 16 per k-step)
            const uint64_t b = ptx::make_mnmajor_sw128_desc(q_addr + 2 * kT8, kT8) + (uint64_t)(128 * ks);
            ptx::mma_bf16_ts(tmem_base + 256u, tmem_base + (uint32_t)(s * 64 + (ks >> 1) * 32 + (ks & 1) * 8), b, idesc_acc,
                             (it > 0 || ks > 0) ? 1u : 0u);
          }
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {     // dK += dS^T Q
            const uint64_t b = ptx::make_mnmajor_sw128_desc(q_addr, kT8) + (uint64_t)(128 * ks);
            ptx::mma_bf16_ts(tmem_base + 384u, tmem_base + (uint32_t)(128 + s * 64 + (ks >> 1) * 32 + (ks & 1) * 8), b, idesc_acc,
                             (it > 0 || ks > 0) ? 1u : 0u);
          }
          ptx::tc_commit(bar(5 + st));     // Q / dO slot may be refilled
        }
      }
      if (ptx::elect_one()) {
        ptx::tc_commit(bar(13));
      }
      if (timed && lane == 0) { p.timing[8] = w_q; p.timing[9] = w_pds; p.timing[12] = clock64() - w0; p.timing[13] = n_it; }
    }
    __syncwarp();
  } else {
    // ===== softmax warps: group g owns buffer g and the steps it = g (mod 2); a thread owns one KEY (TMEM lane) and 32 of
    // the 64 query-row columns =====
    const int q = warp & 3;
    const int idx = (warp - 2) >> 2;             // 0..3
    const int g = idx & 1, h = idx >> 1;
    const int m = q * 32 + lane;                 // key row inside the tile
    const int kpos = key0 + m;
    const int tsg = h * 128 + m;                 // thread index inside the group, 0..255
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    const int c0 = h * 32;
    const int mc = tsg & 63, mwhich = tsg >> 6;  // 0 lse | 1 delta | 2 query position | 3 idle
    const int mqi = mc / p.R, mr = mc - mqi * p.R;
    float* mg = meta + g * 384;
    auto fetch = [&](int u) -> float {
      const int tok = u * nqa + mqi;
      const bool valid = (mqi < nqa) && (tok < q_len);
      if (mwhich == 2) return __int_as_float(valid ? pos0 + tok : -1);
      if (mwhich == 3) return 0.f;
      if (!valid) return mwhich == 0 ? INFINITY : 0.f;
      const float* src = mwhich == 0 ? p.lse : p.delta;
      return __ldg(src + (int64_t)(seg0 + tok) * p.n_q + (kvh * p.R + mr));
    };
    const int n_own = n_it > g ? (n_it - g + 1) >> 1 : 0;
    if (n_own > 0 && mwhich < 3) mg[mwhich * 64 + mc] = fetch(u_first + g);
    float nxt = (g + 2 < n_it) ? fetch(u_first + g + 2) : 0.f;
    group_bar256(g);                             // metadata of own step 0 is visible to the whole group

    const bool timed = kTimed && p.timing != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && warp == 2 && lane == 0;
    long long tph[5] = {0, 0, 0, 0, 0};
    for (int j = 0; j < n_own; ++j) {
      const int it = 2 * j + g;
      const int u = u_first + it;
      const long long tk0 = clock64();
      const float* mb = mg + (j & 1) * 192;
      const long long tk1 = clock64();
      ptx::mbar_wait(bar(9 + g), (uint32_t)(j & 1));
      ptx::tc_fence_after_sync();
      // No barrier inside the loop.  S^T / dP^T of own step j exist only after dV / dK of own step j - 1 were issued, i.e. after
      // EVERY warp of the group arrived at the end of step j - 1: (a) nobody reads the metadata buffer of step j - 1 any more,
      // so it may be overwritten with step j + 1's now; (b) the writes made during step j - 1 for step j happened before those
      // arrivals (release) and this wait (acquire) -- they are visible.
      if (it + 2 < n_it) {
        if (mwhich < 3) mg[((j + 1) & 1) * 192 + mwhich * 64 + mc] = nxt;
        nxt = (it + 4 < n_it) ? fetch(u + 4) : 0.f;            // global load in flight across a whole step
      }
      const long long tk2 = clock64();
      uint32_t sv[32], dv[32];
      ptx::tmem_ld_32x32b_x32(lane_addr + (uint32_t)(g * 64 + c0), sv);
      ptx::tmem_ld_wait();
      ptx::tmem_ld_32x32b_x32(lane_addr + (uint32_t)(128 + g * 64 + c0), dv);   // completes under the exponentials
      const bool diag = pos0 + u * nqa < key0 + 127;
      float pe[32];
#pragma unroll
      for (int e = 0; e < 32; e += 4) {
        const float4 l0 = *reinterpret_cast<const float4*>(mb + c0 + e);
        pe[e + 0] = ex2f(fmaf(__uint_as_float(sv[e + 0]), p.scale_log2, -l0.x));
        pe[e + 1] = ex2f(fmaf(__uint_as_float(sv[e + 1]), p.scale_log2, -l0.y));
        pe[e + 2] = ex2f(fmaf(__uint_as_float(sv[e + 2]), p.scale_log2, -l0.z));
        pe[e + 3] = ex2f(fmaf(__uint_as_float(sv[e + 3]), p.scale_log2, -l0.w));
      }
      if (diag) {
#pragma unroll
        for (int e = 0; e < 32; e += 4) {
          const int4 q0 = *reinterpret_cast<const int4*>(mb + 128 + c0 + e);
          if (kpos > q0.x) pe[e + 0] = 0.f;
          if (kpos > q0.y) pe[e + 1] = 0.f;
          if (kpos > q0.z) pe[e + 2] = 0.f;
          if (kpos > q0.w) pe[e + 3] = 0.f;
        }
      }
      const long long tk3 = clock64();
      ptx::tmem_ld_wait();
      uint32_t pp[16], dd[16];
#pragma unroll
      for (int e = 0; e < 32; e += 4) {
        const float4 d0 = *reinterpret_cast<const float4*>(mb + 64 + c0 + e);
        pp[(e >> 1)] = pk2(pe[e], pe[e + 1]);
        pp[(e >> 1) + 1] = pk2(pe[e + 2], pe[e + 3]);
        dd[(e >> 1)] = pk2(pe[e] * (__uint_as_float(dv[e]) - d0.x), pe[e + 1] * (__uint_as_float(dv[e + 1]) - d0.y));
        dd[(e >> 1) + 1] = pk2(pe[e + 2] * (__uint_as_float(dv[e + 2]) - d0.z), pe[e + 3] * (__uint_as_float(dv[e + 3]) - d0.w));
      }
      // packed P^T / dS^T over the first 16 of this thread's own 32 columns of S^T[g] / dP^T[g]
      ptx::tmem_st_32x32b_x16(lane_addr + (uint32_t)(g * 64 + c0), pp);
      ptx::tmem_st_32x32b_x16(lane_addr + (uint32_t)(128 + g * 64 + c0), dd);
      ptx::tmem_st_wait();
      ptx::tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(bar(11 + g));
      if (kTimed) {
        const long long tk4 = clock64();
        tph[0] += tk1 - tk0; tph[1] += tk2 - tk1; tph[2] += tk3 - tk2; tph[3] += tk4 - tk3;
      }
    }
    if (timed) {
#pragma unroll
      for (int e = 0; e < 4; ++e) p.timing[e] = tph[e];
    }

    // ---- epilogue: 32 head-dim columns of this key's dV and dK rows ----
    ptx::mbar_wait(bar(13), 0);
    ptx::tc_fence_after_sync();
    const bool valid = kpos < pos0 + q_len;
    __nv_bfloat16* drow = p.dkv + (int64_t)(kv0 + kpos) * p.dkv_stride + kvh * kD + idx * 32;
#pragma unroll
    for (int which = 0; which < 2; ++which) {     // 0: dV (TMEM 256..383), 1: dK (384..511)
      uint32_t v0[32];
      ptx::tmem_ld_32x32b_x32(lane_addr + (uint32_t)(256 + which * 128 + idx * 32), v0);
      ptx::tmem_ld_wait();
      const float sc = which ? p.sm_scale : 1.f;
      if (valid) {
        __nv_bfloat16* dst = drow + (which ? p.dkv_col_k : p.dkv_col_v);
#pragma unroll
        for (int d = 0; d < 32; d += 8) {
          uint4 a;
          a.x = pk2(__uint_as_float(v0[d]) * sc, __uint_as_float(v0[d + 1]) * sc);
          a.y = pk2(__uint_as_float(v0[d + 2]) * sc, __uint_as_float(v0[d + 3]) * sc);
          a.z = pk2(__uint_as_float(v0[d + 4]) * sc, __uint_as_float(v0[d + 5]) * sc);
          a.w = pk2(__uint_as_float(v0[d + 6]) * sc, __uint_as_float(v0[d + 7]) * sc);
          *reinterpret_cast<uint4*>(dst + d) = a;
        }
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

constexpr int kKSlots4 = 3;          // K of step i serves S(i) (issued a step early) and dQ(i) (issued last): two slots leave no slack
constexpr int kSmemDq4 = 1024 + 4 * kT16 + kKSlots4 * 2 * kT16 + 2 * 2 * kT16 + 8 * 24 + 16;
static_assert(kSmemDq4 <= 232448, "dq4 kernel exceeds the 227 KB shared-memory limit");

template <bool kTimed>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreadsB4, 1)
attn_bwd_dq4_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_do,
                    const __grid_constant__ CUtensorMap tm_kv, BwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t q_smem = base;                         // Q lo | Q hi
  const uint32_t do_smem = base + 2 * kT16;             // dO lo | dO hi
  const uint32_t k_smem = base + 4 * kT16;              // 3 slots x (K lo | K hi)
  const uint32_t v_smem = k_smem + kKSlots4 * 2 * kT16; // 2 slots x (V lo | V hi)
  const uint32_t bar_base = v_smem + 2 * 2 * kT16;
  auto bar = [&](int i) { return bar_base + 8u * (uint32_t)i; };
  // 0 q_full | 1..3 k_full | 4..6 k_empty | 7,8 s_full | 9,10 s_empty (16 warps) | 11 dp_full | 12 ds_full (16 warps) |
  // 13,14 v_full | 15,16 v_empty | 17 dq_done
  const uint32_t tmem_slot = bar(18);
  constexpr uint32_t cDP = 256u, cDQ = 384u;            // TMEM columns: S[2] 0,128 | dP 256 | dQ 384

  const int qtile = (int)(gridDim.x - 1 - blockIdx.x);
  const int kvh = blockIdx.y, z = blockIdx.z;
  const uint32_t rank = ptx::cluster_ctarank();
  const int q_len = ptx::warp_uniform(p.seg_len[z]);
  if ((qtile & ~1) * p.nq >= q_len) return;             // uniform across the cluster
  const int seg0 = ptx::warp_uniform(p.seg_start[z]);
  const int pos0 = p.seg_pos0 ? ptx::warp_uniform(p.seg_pos0[z]) : 0;
  const int kv0 = p.seg_kv_start ? ptx::warp_uniform(p.seg_kv_start[z]) : seg0;
  const int t0 = qtile * p.nq;
  const int row0 = seg0 + t0;
  const int n_valid = t0 >= q_len ? 0 : ((q_len - t0) < p.nq ? (q_len - t0) : p.nq);
  const int pair_rows = ((qtile | 1) + 1) * p.nq;
  const int kv_end = pos0 + (pair_rows < q_len ? pair_rows : q_len);
  const int n_it = (kv_end + 127) / 128;
  const int last_page = (kv_end - 1) / 64;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < 18; ++i) {
      int cnt = 1;
      if ((i >= 4 && i <= 6) || i == 15 || i == 16) cnt = 2;
      if (i == 9 || i == 10 || i == 12) cnt = 16;
      ptx::mbar_init(bar(i), cnt);
    }
    ptx::fence_barrier_init();
    ptx::fence_proxy_async();
    ptx::prefetch_tensormap(&tm_q);
    ptx::prefetch_tensormap(&tm_do);
    ptx::prefetch_tensormap(&tm_kv);
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_slot, 512);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before_sync();
  ptx::cluster_sync();
  ptx::tc_fence_after_sync();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  tmem_base = ptx::warp_uniform(tmem_base);

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      ptx::mbar_arrive_expect_tx(bar(0), (uint32_t)(4 * 128 * p.R * p.nq));
      ptx::tma_load_3d(q_smem, &tm_q, 0, kvh * p.R, row0, bar(0), ptx::kEvictFirst);
      ptx::tma_load_3d(q_smem + kT16, &tm_q, 64, kvh * p.R, row0, bar(0), ptx::kEvictFirst);
      ptx::tma_load_3d(do_smem, &tm_do, 0, kvh * p.R, row0, bar(0), ptx::kEvictFirst);
      ptx::tma_load_3d(do_smem + kT16, &tm_do, 64, kvh * p.R, row0, bar(0), ptx::kEvictFirst);
      auto load_page = [&](int it, int kv) {
        const int s = kv ? (it & 1) : (it % kKSlots4);
        const uint32_t ph = (uint32_t)((kv ? (it >> 1) : (it / kKSlots4)) & 1);
        const uint32_t full = bar((kv ? 13 : 1) + s), empty = bar((kv ? 15 : 4) + s);
        ptx::mbar_wait(empty, ph ^ 1u);
        ptx::mbar_arrive_expect_tx(full, (uint32_t)(2 * kT16));
        int pg = 2 * it + (int)rank;
        if (pg > last_page) pg = last_page;            // tail: re-read the last page, its keys are causally masked
        const int row = kv0 + pg * 64;
        const int c0 = (kv ? p.col_v : p.col_k) + kvh * kD;
        const uint32_t dst = (kv ? v_smem : k_smem) + (uint32_t)(s * 2 * kT16) + (uint32_t)(rank * kT8);
        ptx::tma_load_2d_multicast(dst, &tm_kv, c0, row, full, 3, ptx::kEvictLast);
        ptx::tma_load_2d_multicast(dst + kT16, &tm_kv, c0 + 64, row, full, 3, ptx::kEvictLast);
      };
      load_page(0, 0);
      for (int it = 0; it < n_it; ++it) {
        if (it + 1 < n_it) load_page(it + 1, 0);     // K runs one step ahead of V: S(i+1) is issued during step i
        load_page(it, 1);
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: S(i+1) | dQ(i) | dP(i+1), in that order =====
    {   // the whole warp, converged: one elected lane issues each tcgen05 instruction (ptx::elect_one)
      constexpr uint32_t idesc_kk = ptx::make_idesc_bf16_f32(128, 128);
      constexpr uint32_t idesc_dq = ptx::make_idesc_bf16_f32(128, kD) | (1u << 16);
      long long w_k = 0, w_sd = 0, w_ds = 0, w_v = 0;
      const bool timed = kTimed && p.timing != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0;
      const long long w0 = clock64();
      auto issue_s = [&](int j) {
        const int s = j & 1, ks_ = j % kKSlots4;
        const uint32_t ph = (uint32_t)((j >> 1) & 1);
        const long long c0 = clock64();
        ptx::mbar_wait(bar(1 + ks_), (uint32_t)((j / kKSlots4) & 1));   // K of step j landed
        const long long c1 = clock64();
        ptx::mbar_wait(bar(9 + s), ph ^ 1u);       // S[s] drained (step j - 2)
        if (kTimed) { w_k += c1 - c0; w_sd += clock64() - c1; }
        ptx::tc_fence_after_sync();
        const uint32_t k_addr = k_smem + (uint32_t)(ks_ * 2 * kT16);
        if (ptx::elect_one()) {
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) {
            const uint64_t b = ptx::make_kmajor_sw128_desc(k_addr + (uint32_t)((ks >> 2) * kT16)) + (uint64_t)(2 * (ks & 3));
            ptx::mma_bf16_ss(tmem_base + (uint32_t)(s * 128),
                             ptx::make_kmajor_sw128_desc(q_smem + (uint32_t)((ks >> 2) * kT16)) + (uint64_t)(2 * (ks & 3)), b,
                             idesc_kk, ks > 0 ? 1u : 0u);
          }
          ptx::tc_commit(bar(7 + s));
        }
      };
      auto issue_dp = [&](int j) {
        const int s = j & 1;
        const long long c0 = clock64();
        ptx::mbar_wait(bar(13 + s), (uint32_t)((j >> 1) & 1));   // V of step j landed
        if (kTimed) w_v += clock64() - c0;
        // dP's columns hold dS of step j - 1 until its dQ UMMAs -- issued BEFORE this call -- have read it (in order)
        ptx::tc_fence_after_sync();
        const uint32_t v_addr = v_smem + (uint32_t)(s * 2 * kT16);
        if (ptx::elect_one()) {
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) {
            const uint64_t b = ptx::make_kmajor_sw128_desc(v_addr + (uint32_t)((ks >> 2) * kT16)) + (uint64_t)(2 * (ks & 3));
            ptx::mma_bf16_ss(tmem_base + cDP,
                             ptx::make_kmajor_sw128_desc(do_smem + (uint32_t)((ks >> 2) * kT16)) + (uint64_t)(2 * (ks & 3)), b,
                             idesc_kk, ks > 0 ? 1u : 0u);
          }
          ptx::tc_commit(bar(11));
          ptx::tc_commit_multicast(bar(15 + s), 3);  // V slot consumed: tell BOTH producers
        }
      };
      ptx::mbar_wait(bar(0), 0);
      issue_s(0);
      issue_dp(0);
      for (int i = 0; i < n_it; ++i) {
        if (i + 1 < n_it) issue_s(i + 1);
        const int ks_ = i % kKSlots4;
        const long long c0 = clock64();
        ptx::mbar_wait(bar(12), (uint32_t)(i & 1));              // dS of step i is in TMEM
        if (kTimed) w_ds += clock64() - c0;
        ptx::tc_fence_after_sync();
        const uint32_t k_addr = k_smem + (uint32_t)(ks_ * 2 * kT16);
        if (ptx::elect_one()) {
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) {           // dQ += dS K     (K read as stored: keys are the contraction rows)
            const uint64_t b = ptx::make_mnmajor_sw128_desc(k_addr, kT16) + (uint64_t)(128 * ks);
            ptx::mma_bf16_ts(tmem_base + cDQ, tmem_base + cDP + (uint32_t)((ks >> 1) * 32 + (ks & 1) * 8), b, idesc_dq,
                             (i > 0 || ks > 0) ? 1u : 0u);
          }
          ptx::tc_commit_multicast(bar(4 + ks_), 3);   // K slot consumed: tell BOTH producers
        }
        if (i + 1 < n_it) issue_dp(i + 1);
      }
      if (ptx::elect_one()) {
        ptx::tc_commit(bar(17));
      }
      if (timed && lane == 0) {
        p.timing[8] = w_k; p.timing[9] = w_sd; p.timing[10] = w_ds; p.timing[11] = w_v; p.timing[12] = clock64() - w0; p.timing[13] = n_it;
      }
    }
    __syncwarp();
  } else {
    // ===== softmax warps: four threads share one (token, head) row; thread k works on keys [32 k, 32 k + 32) of a step =====
    const int q = warp & 3;
    const int k = (warp - 2) >> 2;               // 0..3
    const int m = q * 32 + lane;
    const int qi = m / p.R, r = m - qi * p.R;
    const int qpos = pos0 + t0 + qi;
    const bool valid = qi < n_valid && qi < p.nq;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    const int64_t stat = (int64_t)(row0 + qi) * p.n_q + (kvh * p.R + r);
    const float lse_row = valid ? __ldg(p.lse + stat) : INFINITY;   // padding rows: P = 0
    const float delta_row = valid ? __ldg(p.delta + stat) : 0.f;

    const bool timed = kTimed && p.timing != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && warp == 2 && lane == 0;
    long long tph[5] = {0, 0, 0, 0, 0};
    for (int i = 0; i < n_it; ++i) {
      const int s = i & 1;
      const long long c0 = clock64();
      ptx::mbar_wait(bar(7 + s), (uint32_t)((i >> 1) & 1));
      ptx::tc_fence_after_sync();
      const long long c1 = clock64();
      uint32_t v0[32];
      ptx::tmem_ld_32x32b_x32(lane_addr + (uint32_t)(s * 128 + k * 32), v0);
      ptx::tmem_ld_wait();
      ptx::tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(bar(9 + s));           // this warp's part of S[s] is in registers
      // P of this step: runs while the tensor core is still busy with dQ(i-1) and dP(i)
      float sv[32];
#pragma unroll
      for (int e = 0; e < 32; ++e) sv[e] = ex2f(fmaf(__uint_as_float(v0[e]), p.scale_log2, -lse_row));
      if (i * 128 + 127 > pos0 + t0) {
        const int key0 = i * 128 + k * 32;
#pragma unroll
        for (int e = 0; e < 32; ++e)
          if (key0 + e > qpos) sv[e] = 0.f;
      }
      const long long c2 = clock64();
      ptx::mbar_wait(bar(11), (uint32_t)(i & 1));            // dP of step i
      ptx::tc_fence_after_sync();
      const long long c3 = clock64();
      uint32_t d0[32];
      ptx::tmem_ld_32x32b_x32(lane_addr + cDP + (uint32_t)(k * 32), d0);
      ptx::tmem_ld_wait();
      uint32_t pp[16];
#pragma unroll
      for (int e = 0; e < 16; ++e)
        pp[e] = pk2(sv[2 * e] * (__uint_as_float(d0[2 * e]) - delta_row), sv[2 * e + 1] * (__uint_as_float(d0[2 * e + 1]) - delta_row));
      // packed dS over the first 16 of this thread's own 32 dP columns
      ptx::tmem_st_32x32b_x16(lane_addr + cDP + (uint32_t)(k * 32), pp);
      ptx::tmem_st_wait();
      ptx::tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(bar(12));              // -> dQ += dS K
      if (kTimed) {
        const long long c4 = clock64();
        tph[0] += c1 - c0; tph[1] += c2 - c1; tph[2] += c3 - c2; tph[3] += c4 - c3;
      }
    }
    if (timed) {
#pragma unroll
      for (int e = 0; e < 4; ++e) p.timing[e] = tph[e];
    }

    // ---- epilogue: 32 head-dim columns of this row's dQ ----
    ptx::mbar_wait(bar(17), 0);
    ptx::tc_fence_after_sync();
    uint32_t v0[32];
    ptx::tmem_ld_32x32b_x32(lane_addr + cDQ + (uint32_t)(k * 32), v0);
    ptx::tmem_ld_wait();
    if (valid) {
      __nv_bfloat16* dst = p.dqkv + (int64_t)(row0 + qi) * p.dqkv_stride + (kvh * p.R + r) * kD + k * 32;
      const float sc = p.sm_scale;
#pragma unroll
      for (int d = 0; d < 32; d += 8) {
        uint4 a;
        a.x = pk2(__uint_as_float(v0[d]) * sc, __uint_as_float(v0[d + 1]) * sc);
        a.y = pk2(__uint_as_float(v0[d + 2]) * sc, __uint_as_float(v0[d + 3]) * sc);
        a.z = pk2(__uint_as_float(v0[d + 4]) * sc, __uint_as_float(v0[d + 5]) * sc);
        a.w = pk2(__uint_as_float(v0[d + 6]) * sc, __uint_as_float(v0[d + 7]) * sc);
        *reinterpret_cast<uint4*>(dst + d) = a;
      }
    }
  }

  ptx::tc_fence_before_sync();
  ptx::cluster_sync();   // the partner may still multicast into this CTA's shared memory / arrive on its barriers
  if (warp == 1) {
    ptx::tc_fence_after_sync();
    ptx::tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace
}  // namespace prl

using namespace prl;

namespace prl { namespace { int g_bwd_generation = [] { const char* e = getenv("PRL_ATTN_BWD"); return (e && e[0] >= '1' && e[0] <= '4') ? e[0] - '0' : 4; }(); } }

namespace prl { namespace { long long* g_bwd_timing = nullptr; int g_bwd_timing_kernel = 1; } }
// measurement only: per-phase cycle sums of CTA (0,0,0) of the generation-4 dQ kernel.  [0..3] one softmax warp: wait S |
// tcgen05.ld + exp2 | wait dP | tcgen05.ld + dS + tcgen05.st + arrive; [8..11] MMA warp waiting for K | S drained | dS | V,
// [12] MMA warp total, [13] steps.  NULL switches it off.
extern "C" int prl_attn_debug_bwd_timing(int64_t* out16_device) {
  prl::g_bwd_timing = (long long*)out16_device;
  prl::g_bwd_timing_kernel = 1;
  return PRL_OK;
}
// same for the dK/dV kernel: [0..3] one softmax warp: group barrier | wait S^T, dP^T | tcgen05.ld + exp2 | dS + tcgen05.st +
// arrive; [8] MMA warp waiting for Q / dO, [9] for P^T / dS^T, [12] MMA warp total, [13] sub-steps
extern "C" int prl_attn_debug_bwd_timing_dkdv(int64_t* out16_device) {
  prl::g_bwd_timing = (long long*)out16_device;
  prl::g_bwd_timing_kernel = 0;
  return PRL_OK;
}

extern "C" int prl_attn_set_bwd_generation(int32_t gen) {
  PRL_CHECK_ARG(gen >= 1 && gen <= 4, "prl_attn_set_bwd_generation: 1 (P / dS operands through shared memory), 2 (through TMEM), "
                "3 (2 + the dQ kernel keeps Q and dO in TMEM) or 4 (16 decoupled softmax warps)");
  prl::g_bwd_generation = gen;
  return PRL_OK;
}

extern "C" size_t prl_attn_varlen_bwd_workspace_bytes(int32_t T, int32_t n_q) { return (size_t)T * (size_t)n_q * sizeof(float); }

// dqkv[T, dqkv_stride] <- gradients of the packed (roped) q | k | v given d_out; every row of every segment is written.
extern "C" int prl_attn_varlen_bwd(const void* qkv, int64_t qkv_stride, int32_t T, const int32_t* seg_start,
                                   const int32_t* seg_len, int32_t n_seg, int32_t max_seg_len, int32_t n_q,
                                   int32_t n_kv, int32_t head_dim, float sm_scale, const void* out_bf16,
                                   const void* d_out_bf16, const float* lse, void* dqkv, int64_t dqkv_stride,
                                   void* workspace, size_t workspace_bytes, prl_stream_t stream_) {
  PRL_CHECK_ARG(qkv && seg_start && seg_len && out_bf16 && d_out_bf16 && lse && dqkv && workspace,
                "prl_attn_varlen_bwd: NULL argument");
  PRL_CHECK_ARG(head_dim == kD, "prl_attn_varlen_bwd: head_dim must be 128");
  PRL_CHECK_ARG(T >= 1 && n_seg >= 1 && max_seg_len >= 1 && n_kv >= 1 && n_q % n_kv == 0 && n_q / n_kv <= 64,
                "prl_attn_varlen_bwd: bad shape (GQA group size must be <= 64)");
  const int64_t width = (int64_t)(n_q + 2 * n_kv) * kD;
  PRL_CHECK_ARG(qkv_stride >= width && qkv_stride % 8 == 0 && dqkv_stride >= width && dqkv_stride % 8 == 0,
                "prl_attn_varlen_bwd: bad row stride");
  PRL_CHECK_ARG(workspace_bytes >= prl_attn_varlen_bwd_workspace_bytes(T, n_q), "prl_attn_varlen_bwd: workspace too small");
  cudaStream_t stream = (cudaStream_t)stream_;
  float* delta = (float*)workspace;
  {
    const int64_t rows = (int64_t)T * n_q;
    attn_delta_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, stream>>>((const __nv_bfloat16*)out_bf16,
                                                                     (const __nv_bfloat16*)d_out_bf16, rows, delta);
    PRL_LAUNCH_CHECK();
  }
  BwdParams p;
  p.lse = lse; p.delta = delta; p.dqkv = (__nv_bfloat16*)dqkv; p.dqkv_stride = dqkv_stride;
  p.seg_start = seg_start; p.seg_len = seg_len; p.n_q = n_q; p.n_kv = n_kv; p.R = n_q / n_kv;
  p.col_k = n_q * kD; p.col_v = (n_q + n_kv) * kD;
  p.scale_log2 = sm_scale * 1.4426950408889634f; p.sm_scale = sm_scale;
  p.q_base = (const __nv_bfloat16*)qkv; p.q_stride = qkv_stride;
  p.do_base = (const __nv_bfloat16*)d_out_bf16; p.do_stride = (int64_t)n_q * kD;
  p.timing = g_bwd_timing;
  p.seg_pos0 = nullptr; p.seg_kv_start = nullptr; p.dkv = (__nv_bfloat16*)dqkv; p.dkv_stride = dqkv_stride;
  p.dkv_col_k = p.col_k; p.dkv_col_v = p.col_v;
  CUtensorMap tkv, tq, tdo;
  int rc = make_tmap_2d_bf16(&tkv, qkv, (uint64_t)width, (uint64_t)T, (uint64_t)qkv_stride * 2, 64, 64);
  if (rc) return rc;
  {
    // ---- dK, dV ----
    p.nq = 64 / p.R;
    rc = make_tmap_3d_bf16(&tq, qkv, kD, (uint64_t)n_q, (uint64_t)T, kD * 2, (uint64_t)qkv_stride * 2, 64, (uint32_t)p.R, (uint32_t)p.nq);
    if (rc) return rc;
    rc = make_tmap_3d_bf16(&tdo, d_out_bf16, kD, (uint64_t)n_q, (uint64_t)T, kD * 2, (uint64_t)n_q * kD * 2, 64, (uint32_t)p.R, (uint32_t)p.nq);
    if (rc) return rc;
    static SmemAttr attr = {};
    dim3 grid((unsigned)((max_seg_len + 127) / 128), (unsigned)n_kv, (unsigned)n_seg);
    if (g_bwd_generation == 4) {
      static SmemAttr attr4 = {};
      if (g_bwd_timing != nullptr && g_bwd_timing_kernel == 0) {
        static SmemAttr attr4t = {};
        PRL_CUDA(ensure_smem(attn_bwd_dkdv4_kernel<true>, kSmemDkdv4, attr4t));
        attn_bwd_dkdv4_kernel<true><<<grid, kThreadsB4, (size_t)kSmemDkdv4, stream>>>(tq, tdo, tkv, p);
      } else {
        PRL_CUDA(ensure_smem(attn_bwd_dkdv4_kernel<false>, kSmemDkdv4, attr4));
        attn_bwd_dkdv4_kernel<false><<<grid, kThreadsB4, (size_t)kSmemDkdv4, stream>>>(tq, tdo, tkv, p);
      }
    } else if (g_bwd_generation == 1) {   // generations 2 and 3 share the dK / dV kernel
      PRL_CUDA(ensure_smem(attn_bwd_dkdv_kernel<false>, kSmemDkdv, attr));
      attn_bwd_dkdv_kernel<false><<<grid, kThreadsB, (size_t)kSmemDkdv, stream>>>(tq, tdo, tkv, p);
    } else {
      static SmemAttr attr2 = {};
      PRL_CUDA(ensure_smem(attn_bwd_dkdv_kernel<true>, kSmemDkdv, attr2));
      attn_bwd_dkdv_kernel<true><<<grid, kThreadsB, (size_t)kSmemDkdv, stream>>>(tq, tdo, tkv, p);
    }
    PRL_LAUNCH_CHECK();
  }
  {
    // ---- dQ ----
    p.nq = 128 / p.R;
    rc = make_tmap_3d_bf16(&tq, qkv, kD, (uint64_t)n_q, (uint64_t)T, kD * 2, (uint64_t)qkv_stride * 2, 64, (uint32_t)p.R, (uint32_t)p.nq);
    if (rc) return rc;
    rc = make_tmap_3d_bf16(&tdo, d_out_bf16, kD, (uint64_t)n_q, (uint64_t)T, kD * 2, (uint64_t)n_q * kD * 2, 64, (uint32_t)p.R, (uint32_t)p.nq);
    if (rc) return rc;
    static SmemAttr attr = {};
    dim3 grid((unsigned)(((max_seg_len + p.nq - 1) / p.nq + 1) & ~1), (unsigned)n_kv, (unsigned)n_seg);
    if (g_bwd_generation == 4) {
      static SmemAttr attr4 = {};
      if (g_bwd_timing != nullptr && g_bwd_timing_kernel == 1) {
        static SmemAttr attr4t = {};
        PRL_CUDA(ensure_smem(attn_bwd_dq4_kernel<true>, kSmemDq4, attr4t));
        attn_bwd_dq4_kernel<true><<<grid, kThreadsB4, (size_t)kSmemDq4, stream>>>(tq, tdo, tkv, p);
      } else {
        PRL_CUDA(ensure_smem(attn_bwd_dq4_kernel<false>, kSmemDq4, attr4));
        attn_bwd_dq4_kernel<false><<<grid, kThreadsB4, (size_t)kSmemDq4, stream>>>(tq, tdo, tkv, p);
      }
    } else if (g_bwd_generation == 1) {
      PRL_CUDA(ensure_smem(attn_bwd_dq_kernel<1>, kSmemDq, attr));
      attn_bwd_dq_kernel<1><<<grid, kThreadsB, (size_t)kSmemDq, stream>>>(tq, tdo, tkv, p);
    } else if (g_bwd_generation == 2) {
      static SmemAttr attr2 = {};
      PRL_CUDA(ensure_smem(attn_bwd_dq_kernel<2>, kSmemDq, attr2));
      attn_bwd_dq_kernel<2><<<grid, kThreadsB, (size_t)kSmemDq, stream>>>(tq, tdo, tkv, p);
    } else {
      static SmemAttr attr3 = {};
      PRL_CUDA(ensure_smem(attn_bwd_dq_kernel<3>, kSmemDq, attr3));
      attn_bwd_dq_kernel<3><<<grid, kThreadsB, (size_t)kSmemDq, stream>>>(tq, tdo, tkv, p);
    }
    PRL_LAUNCH_CHECK();
  }
  return PRL_OK;
}

// Sequence-parallel form of prl_attn_varlen_bwd (see prl_attn_varlen_fwd_kv for the segment description).  dq[Tq, dq_stride]
// receives the query-head gradients of the LOCAL queries; dkv[Tkv, dkv_stride] = [dK heads | dV heads] receives THIS RANK'S
// contribution to every key row (zero where no local query attends) -- the caller reduce-scatters it over the group.
extern "C" int prl_attn_varlen_bwd_kv(const void* q, int64_t q_stride, int32_t Tq, const void* kv, int64_t kv_stride,
                                      int32_t Tkv, const int32_t* seg_q_start, const int32_t* seg_q_len,
                                      const int32_t* seg_pos0, const int32_t* seg_kv_start, int32_t n_seg,
                                      int32_t max_q_len, int32_t max_kv_len, int32_t n_q, int32_t n_kv, int32_t head_dim,
                                      float sm_scale, const void* out_bf16, const void* d_out_bf16, const float* lse,
                                      void* dq, int64_t dq_stride, void* dkv, int64_t dkv_stride, void* workspace,
                                      size_t workspace_bytes, prl_stream_t stream_) {
  PRL_CHECK_ARG(q && kv && seg_q_start && seg_q_len && seg_pos0 && seg_kv_start && out_bf16 && d_out_bf16 && lse && dq && dkv && workspace,
                "prl_attn_varlen_bwd_kv: NULL argument");
  PRL_CHECK_ARG(head_dim == kD, "prl_attn_varlen_bwd_kv: head_dim must be 128");
  PRL_CHECK_ARG(Tq >= 1 && Tkv >= 1 && n_seg >= 1 && max_q_len >= 1 && max_kv_len >= max_q_len && n_kv >= 1 && n_q % n_kv == 0 &&
                n_q / n_kv <= 64, "prl_attn_varlen_bwd_kv: bad shape (GQA group size must be <= 64)");
  const int64_t kvw = (int64_t)2 * n_kv * kD;
  PRL_CHECK_ARG(q_stride >= (int64_t)n_q * kD && q_stride % 8 == 0 && dq_stride >= (int64_t)n_q * kD && dq_stride % 8 == 0 &&
                kv_stride >= kvw && kv_stride % 8 == 0 && dkv_stride >= kvw && dkv_stride % 8 == 0, "prl_attn_varlen_bwd_kv: bad row stride");
  PRL_CHECK_ARG(workspace_bytes >= prl_attn_varlen_bwd_workspace_bytes(Tq, n_q), "prl_attn_varlen_bwd_kv: workspace too small");
  cudaStream_t stream = (cudaStream_t)stream_;
  float* delta = (float*)workspace;
  {
    const int64_t rows = (int64_t)Tq * n_q;
    attn_delta_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, stream>>>((const __nv_bfloat16*)out_bf16,
                                                                     (const __nv_bfloat16*)d_out_bf16, rows, delta);
    PRL_LAUNCH_CHECK();
  }
  PRL_CUDA(cudaMemset2DAsync(dkv, (size_t)dkv_stride * 2, 0, (size_t)kvw * 2, (size_t)Tkv, stream));
  BwdParams p;
  p.lse = lse; p.delta = delta; p.dqkv = (__nv_bfloat16*)dq; p.dqkv_stride = dq_stride;
  p.seg_start = seg_q_start; p.seg_len = seg_q_len; p.n_q = n_q; p.n_kv = n_kv; p.R = n_q / n_kv;
  p.col_k = 0; p.col_v = n_kv * kD;
  p.scale_log2 = sm_scale * 1.4426950408889634f; p.sm_scale = sm_scale;
  p.q_base = (const __nv_bfloat16*)q; p.q_stride = q_stride;
  p.do_base = (const __nv_bfloat16*)d_out_bf16; p.do_stride = (int64_t)n_q * kD;
  p.timing = nullptr;
  p.seg_pos0 = seg_pos0; p.seg_kv_start = seg_kv_start; p.dkv = (__nv_bfloat16*)dkv; p.dkv_stride = dkv_stride;
  p.dkv_col_k = 0; p.dkv_col_v = n_kv * kD;
  CUtensorMap tkv, tq, tdo;
  int rc = make_tmap_2d_bf16(&tkv, kv, (uint64_t)kvw, (uint64_t)Tkv, (uint64_t)kv_stride * 2, 64, 64);
  if (rc) return rc;
  {
    p.nq = 64 / p.R;
    rc = make_tmap_3d_bf16(&tq, q, kD, (uint64_t)n_q, (uint64_t)Tq, kD * 2, (uint64_t)q_stride * 2, 64, (uint32_t)p.R, (uint32_t)p.nq);
    if (rc) return rc;
    rc = make_tmap_3d_bf16(&tdo, d_out_bf16, kD, (uint64_t)n_q, (uint64_t)Tq, kD * 2, (uint64_t)n_q * kD * 2, 64, (uint32_t)p.R, (uint32_t)p.nq);
    if (rc) return rc;
    dim3 grid((unsigned)((max_kv_len + 127) / 128), (unsigned)n_kv, (unsigned)n_seg);
    static SmemAttr attr = {};
    PRL_CUDA(ensure_smem(attn_bwd_dkdv4_kernel<false>, kSmemDkdv4, attr));
    attn_bwd_dkdv4_kernel<false><<<grid, kThreadsB4, (size_t)kSmemDkdv4, stream>>>(tq, tdo, tkv, p);
    PRL_LAUNCH_CHECK();
  }
  {
    p.nq = 128 / p.R;
    rc = make_tmap_3d_bf16(&tq, q, kD, (uint64_t)n_q, (uint64_t)Tq, kD * 2, (uint64_t)q_stride * 2, 64, (uint32_t)p.R, (uint32_t)p.nq);
    if (rc) return rc;
    rc = make_tmap_3d_bf16(&tdo, d_out_bf16, kD, (uint64_t)n_q, (uint64_t)Tq, kD * 2, (uint64_t)n_q * kD * 2, 64, (uint32_t)p.R, (uint32_t)p.nq);
    if (rc) return rc;
    dim3 grid((unsigned)(((max_q_len + p.nq - 1) / p.nq + 1) & ~1), (unsigned)n_kv, (unsigned)n_seg);
    static SmemAttr attr = {};
    PRL_CUDA(ensure_smem(attn_bwd_dq4_kernel<false>, kSmemDq4, attr));
    attn_bwd_dq4_kernel<false><<<grid, kThreadsB4, (size_t)kSmemDq4, stream>>>(tq, tdo, tkv, p);
    PRL_LAUNCH_CHECK();
  }
  return PRL_OK;
}

// ---- measurement helper: TMEM read bandwidth of one SM (tcgen05.ld 32x32b.x32 from `warps` warps) -----------------
// The attention kernels read every S / dP / O^T tile out of TMEM once per step; whether that costs 256 or 1024 cycles per
// 64 KB tile decides their structure (profiles/r2_attention.md).  out[0] = cycles, out[1] = bytes read.
namespace prl { namespace {
__global__ void __launch_bounds__(256, 1) tmem_read_bench_kernel(int iters, int warps, long long* out) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) { ptx::tmem_alloc(ptx::smem_u32(&slot), 512); ptx::tmem_relinquish(); }
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  if (warp < warps) {
    for (int i = 0; i < iters; i += 4) {          // four 4 KB reads in flight per warp, one wait
      uint32_t v0[32], v1[32], v2[32], v3[32];
      const uint32_t c = (uint32_t)(((i >> 2) * 128) & 511);
      ptx::tmem_ld_32x32b_x32(base + ((c + 0) & 511), v0);
      ptx::tmem_ld_32x32b_x32(base + ((c + 32) & 511), v1);
      ptx::tmem_ld_32x32b_x32(base + ((c + 64) & 511), v2);
      ptx::tmem_ld_32x32b_x32(base + ((c + 96) & 511), v3);
      ptx::tmem_ld_wait();
#pragma unroll
      for (int e = 0; e < 32; ++e) acc ^= v0[e] ^ v1[e] ^ v2[e] ^ v3[e];
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) { out[0] = t1 - t0; out[1] = (long long)iters * warps * 32 * 32 * 4; }
  if (acc == 0x12345u) out[2] = acc;     // keep the loads alive
  ptx::tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) { ptx::tc_fence_after_sync(); ptx::tmem_dealloc(slot, 512); }
}
} }

extern "C" int prl_debug_tmem_read_bench(int32_t iters, int32_t warps, int64_t* out3_device, prl_stream_t stream_) {
  PRL_CHECK_ARG(out3_device && iters >= 1 && warps >= 1 && warps <= 8, "prl_debug_tmem_read_bench: bad argument");
  prl::tmem_read_bench_kernel<<<(unsigned)prl::num_sms(), 256, 0, (cudaStream_t)stream_>>>((int)iters, (int)warps,
                                                                                            (long long*)out3_device);
  PRL_LAUNCH_CHECK();
  return PRL_OK;
}

// ---- measurement helper: tcgen05.mma throughput per operand configuration, and the softmax <-> tensor-core hand-off ------
// One converged warp issues batches of 8 UMMAs (K = 128) under ONE lane election, the instruction stream between two
// UMMAs is empty (the mode is a template parameter): what is measured is the tensor core, not the issuing thread.
//   0  SS  A, B K-major      128 x 128      1  SS  128 x 64        2  SS  128 x 256       3  SS  B MN-major 128 x 128
//   4  TS (A in TMEM) B K-major 128 x 128   5  TS  B MN-major 128 x 128   6  TS B K-major 128 x 256   7  TS 128 x 64
//   8  SS 128 x 128, two accumulators interleaved       9  SS 128 x 128 batch, then TS MN-major batch (the forward's step)
//   10 hand-off round trip: 1 UMMA -> commit -> 4 warps tcgen05.ld x32 + tcgen05.st x16 -> arrive -> next UMMA
// out[0] = cycles, out[1] = UMMAs issued.
namespace prl { namespace {
template <int kMode>
__global__ void __launch_bounds__(192, 1) mma_bench_kernel(int iters, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_smem = base, b_smem = base + 2 * kT16;          // A: two [128 x 128 B] tiles, B: four
  const uint32_t bar0 = base + 6 * kT16, bar1 = bar0 + 8, slot = bar0 + 16;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (uint32_t i = threadIdx.x; i < (uint32_t)(6 * kT16) / 16; i += blockDim.x) sts_v4(base + i * 16, 0u, 0u, 0u, 0u);
  ptx::fence_proxy_async();
  if (threadIdx.x == 0) {
    ptx::mbar_init(bar0, 1);
    ptx::mbar_init(bar1, 4);
    ptx::fence_barrier_init();
  }
  if (warp == 1) { ptx::tmem_alloc(slot, 512); ptx::tmem_relinquish(); }
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(slot));
  if (warp == 1) {
    constexpr uint32_t id128 = ptx::make_idesc_bf16_f32(128, 128), id64 = ptx::make_idesc_bf16_f32(128, 64);
    constexpr uint32_t id256 = ptx::make_idesc_bf16_f32(128, 256), id128mn = id128 | (1u << 16);
    const uint32_t tb = ptx::warp_uniform(tmem_base);
    const long long t0 = clock64();
    long long n = 0;
    for (int it = 0; it < iters; ++it) {
      if (kMode == 10) {
        if (ptx::elect_one()) {
          ptx::mma_bf16_ss(tb, ptx::make_kmajor_sw128_desc(a_smem), ptx::make_kmajor_sw128_desc(b_smem), id128, 0u);
          ptx::tc_commit(bar0);
        }
        ptx::mbar_wait(bar1, (uint32_t)(it & 1));
        ptx::tc_fence_after_sync();
        ++n;
        continue;
      }
      if (ptx::elect_one()) {
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
          const uint64_t ak = ptx::make_kmajor_sw128_desc(a_smem + (uint32_t)((ks >> 2) * kT16)) + (uint64_t)(2 * (ks & 3));
          const uint64_t bk = ptx::make_kmajor_sw128_desc(b_smem + (uint32_t)((ks >> 2) * kT16)) + (uint64_t)(2 * (ks & 3));
          const uint64_t bmn = ptx::make_mnmajor_sw128_desc(b_smem, kT16) + (uint64_t)(128 * ks);
          const uint32_t at = tb + 256u + (uint32_t)(ks * 8);
          if (kMode == 0 || kMode == 9) ptx::mma_bf16_ss(tb, ak, bk, id128, ks > 0);
          if (kMode == 1) ptx::mma_bf16_ss(tb, ak, bk, id64, ks > 0);
          if (kMode == 2) ptx::mma_bf16_ss(tb, ak, bk, id256, ks > 0);
          if (kMode == 3) ptx::mma_bf16_ss(tb, ak, bmn, id128mn, ks > 0);
          if (kMode == 4) ptx::mma_bf16_ts(tb, at, bk, id128, ks > 0);
          if (kMode == 5) ptx::mma_bf16_ts(tb, at, bmn, id128mn, ks > 0);
          if (kMode == 6) ptx::mma_bf16_ts(tb, at, bk, id256, ks > 0);
          if (kMode == 7) ptx::mma_bf16_ts(tb, at, bk, id64, ks > 0);
          if (kMode == 8) ptx::mma_bf16_ss(tb + (uint32_t)((ks & 1) * 128), ak, bk, id128, ks > 1);
        }
        if (kMode == 9) {
#pragma unroll
          for (int ks = 0; ks < 8; ++ks)
            ptx::mma_bf16_ts(tb + 128u, tb + 256u + (uint32_t)(ks * 8), ptx::make_mnmajor_sw128_desc(b_smem, kT16) + (uint64_t)(128 * ks),
                             id128mn, ks > 0);
        }
      }
      n += kMode == 9 ? 16 : 8;
    }
    if (kMode != 10) {
      if (ptx::elect_one()) ptx::tc_commit(bar0);
      ptx::mbar_wait(bar0, 0);
    }
    const long long t1 = clock64();
    if (blockIdx.x == 0 && lane == 0) { out[0] = t1 - t0; out[1] = n; }
  } else if (warp >= 2 && kMode == 10) {
    const uint32_t lane_addr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    for (int it = 0; it < iters; ++it) {
      ptx::mbar_wait(bar0, (uint32_t)(it & 1));
      ptx::tc_fence_after_sync();
      uint32_t v[32], w[16];
      ptx::tmem_ld_32x32b_x32(lane_addr, v);
      ptx::tmem_ld_wait();
#pragma unroll
      for (int e = 0; e < 16; ++e) w[e] = v[2 * e] ^ v[2 * e + 1];
      ptx::tmem_st_32x32b_x16(lane_addr + 256u, w);
      ptx::tmem_st_wait();
      ptx::tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(bar1);
    }
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) { ptx::tc_fence_after_sync(); ptx::tmem_dealloc(tmem_base, 512); }
}

template <int kMode>
int launch_mma_bench(int iters, long long* out, cudaStream_t stream) {
  const int smem = 6 * kT16 + 1024 + 64;
  static SmemAttr attr = {};
  PRL_CUDA(ensure_smem(mma_bench_kernel<kMode>, smem, attr));
  mma_bench_kernel<kMode><<<(unsigned)num_sms(), 192, (size_t)smem, stream>>>(iters, out);
  PRL_LAUNCH_CHECK();
  return PRL_OK;
}
} }

extern "C" int prl_debug_mma_bench(int32_t mode, int32_t iters, int64_t* out2_device, prl_stream_t stream_) {
  PRL_CHECK_ARG(out2_device && iters >= 1 && mode >= 0 && mode <= 10, "prl_debug_mma_bench: bad argument");
  long long* out = (long long*)out2_device;
  cudaStream_t st = (cudaStream_t)stream_;
  switch (mode) {
    case 0: return prl::launch_mma_bench<0>(iters, out, st);
    case 1: return prl::launch_mma_bench<1>(iters, out, st);
    case 2: return prl::launch_mma_bench<2>(iters, out, st);
    case 3: return prl::launch_mma_bench<3>(iters, out, st);
    case 4: return prl::launch_mma_bench<4>(iters, out, st);
    case 5: return prl::launch_mma_bench<5>(iters, out, st);
    case 6: return prl::launch_mma_bench<6>(iters, out, st);
    case 7: return prl::launch_mma_bench<7>(iters, out, st);
    case 8: return prl::launch_mma_bench<8>(iters, out, st);
    case 9: return prl::launch_mma_bench<9>(iters, out, st);
    default: return prl::launch_mma_bench<10>(iters, out, st);
  }
}
