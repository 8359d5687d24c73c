// Hot path (3): in-flight learner -> sampler weight update as a one-shot P2P copy over NVLink/NVSwitch.
//
// Replaces WeightUpdateManager.send_weight_update (pipelinerl/finetune_loop.py:205-292: HTTP POST +
// 339 per-tensor ncclBroadcast calls) and WorkerExtension.receive_weight_update
// (pipelinerl/vllm1.py:110-127: per tensor torch.empty -> broadcast -> load_weights) together with
// the PyNccl bootstrap of pipelinerl/torch_utils.py:70-94.
//
// Learner and samplers share ONE arena layout (pipelinerl_b200/model.py), so an update is a byte
// copy.  Each sampler exports CUDA-IPC handles of its two arena buffers and a control block once;
// the learner maps them and every update is:
//   1. push_kernel: read the learner's bf16 arena slice ONCE from HBM and store it to up to 8 peer
//      buffers (16-byte vector stores through the NVLink aperture, 8 independent loads in flight per
//      thread), into the sampler's INACTIVE buffer — in-flight sequences keep decoding on the active
//      one, nothing is drained;
//   2. signal_kernel (stream-ordered after 1): system-scope release of {version, arrival count} in
//      every peer's control block.  When all learner ranks that own a slice have arrived the sampler
//      flips buffers at its next token-step boundary (engine.py), i.e. the stall is one graph switch.
// With Ng learner replicas each rank pushes 1/Ng of the bytes to all N samplers, so the NVLink egress
// of every learner GPU is used: time ~ N * bytes / Ng / link_bw.
//
// NVLink-bound: algorithmic bytes = arena bytes per sampler replica.
#include "prl_common.cuh"

namespace prl {
namespace {

constexpr int kMaxPeers = 8;
constexpr int kThreads = 512;
constexpr int kUnroll = 8;

struct PushParams {
  const uint4* src;
  uint4* dst[kMaxPeers];
  int n_dst;
  size_t n16;  // number of 16-byte words
};

__global__ void __launch_bounds__(kThreads) push_kernel(PushParams p) {
  const size_t stride = (size_t)gridDim.x * kThreads * kUnroll;
  for (size_t base = (size_t)blockIdx.x * kThreads * kUnroll; base < p.n16; base += stride) {
    uint4 v[kUnroll];
#pragma unroll
    for (int j = 0; j < kUnroll; ++j) {
      const size_t i = base + (size_t)j * kThreads + threadIdx.x;
      if (i < p.n16) v[j] = ld_stream_u4(p.src + i);
    }
    for (int d = 0; d < p.n_dst; ++d) {
      uint4* dst = p.dst[d];
#pragma unroll
      for (int j = 0; j < kUnroll; ++j) {
        const size_t i = base + (size_t)j * kThreads + threadIdx.x;
        if (i < p.n16) dst[i] = v[j];
      }
    }
  }
}

struct SignalParams {
  unsigned long long* ctrl[kMaxPeers];  // per peer: ctrl[0] = version, ctrl[1] = arrivals (monotonic)
  int n_dst;
  unsigned long long version;
};

__global__ void signal_kernel(SignalParams s) {
  const int d = threadIdx.x;
  if (d >= s.n_dst) return;
  __threadfence_system();
  atomicMax_system(&s.ctrl[d][0], s.version);
  __threadfence_system();
  atomicAdd_system(&s.ctrl[d][1], 1ull);
}

}  // namespace
}  // namespace prl

using namespace prl;

extern "C" int prl_ipc_alloc(size_t bytes, void** dptr) {
  PRL_CHECK_ARG(dptr && bytes > 0, "prl_ipc_alloc: bad argument");
  PRL_CUDA(cudaMalloc(dptr, bytes));
  PRL_CUDA(cudaMemset(*dptr, 0, bytes));
  return PRL_OK;
}
extern "C" int prl_ipc_free(void* dptr) {
  if (dptr) PRL_CUDA(cudaFree(dptr));
  return PRL_OK;
}
extern "C" int prl_ipc_export(const void* dptr, uint8_t handle[64]) {
  PRL_CHECK_ARG(dptr && handle, "prl_ipc_export: NULL argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  cudaIpcMemHandle_t h;
  PRL_CUDA(cudaIpcGetMemHandle(&h, const_cast<void*>(dptr)));
  memcpy(handle, &h, 64);
  return PRL_OK;
}
extern "C" int prl_ipc_open(const uint8_t handle[64], void** dptr) {
  PRL_CHECK_ARG(dptr && handle, "prl_ipc_open: NULL argument");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, 64);
  PRL_CUDA(cudaIpcOpenMemHandle(dptr, h, cudaIpcMemLazyEnablePeerAccess));
  return PRL_OK;
}
extern "C" int prl_ipc_close(void* dptr) {
  if (dptr) PRL_CUDA(cudaIpcCloseMemHandle(dptr));
  return PRL_OK;
}
extern "C" int prl_enable_peer_access(int32_t peer_device) {
  int cur = 0;
  PRL_CUDA(cudaGetDevice(&cur));
  if (cur == peer_device) return PRL_OK;
  int can = 0;
  PRL_CUDA(cudaDeviceCanAccessPeer(&can, cur, peer_device));
  PRL_CHECK_ARG(can, "prl_enable_peer_access: device %d cannot access device %d", cur, peer_device);
  cudaError_t e = cudaDeviceEnablePeerAccess(peer_device, 0);
  if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); return PRL_OK; }
  PRL_CUDA(e);
  return PRL_OK;
}

extern "C" int prl_weights_push(const void* src, void* const* dst, int32_t n_dst, size_t offset_bytes, size_t bytes,
                                int32_t max_ctas, prl_stream_t stream_) {
  PRL_CHECK_ARG(src && dst && n_dst >= 1 && n_dst <= kMaxPeers, "prl_weights_push: need 1..8 destinations");
  PRL_CHECK_ARG(offset_bytes % 16 == 0 && bytes % 16 == 0 && (uintptr_t)src % 16 == 0,
                "prl_weights_push: offset/size/pointers must be 16-byte aligned");
  if (bytes == 0) return PRL_OK;
  PushParams p;
  p.src = reinterpret_cast<const uint4*>(static_cast<const char*>(src) + offset_bytes);
  p.n_dst = n_dst;
  for (int d = 0; d < n_dst; ++d) {
    PRL_CHECK_ARG(dst[d] && (uintptr_t)dst[d] % 16 == 0, "prl_weights_push: destination %d NULL or unaligned", d);
    p.dst[d] = reinterpret_cast<uint4*>(static_cast<char*>(dst[d]) + offset_bytes);
  }
  p.n16 = bytes / 16;
  size_t want = (p.n16 + (size_t)kThreads * kUnroll - 1) / ((size_t)kThreads * kUnroll);
  int ctas = max_ctas > 0 ? max_ctas : num_sms() * 2;
  if ((size_t)ctas > want) ctas = (int)want;
  push_kernel<<<ctas, kThreads, 0, (cudaStream_t)stream_>>>(p);
  PRL_LAUNCH_CHECK();
  return PRL_OK;
}

extern "C" int prl_weights_signal(void* const* ctrl, int32_t n_dst, uint64_t version, prl_stream_t stream_) {
  PRL_CHECK_ARG(ctrl && n_dst >= 1 && n_dst <= kMaxPeers, "prl_weights_signal: need 1..8 destinations");
  SignalParams s;
  s.n_dst = n_dst;
  s.version = version;
  for (int d = 0; d < n_dst; ++d) {
    PRL_CHECK_ARG(ctrl[d], "prl_weights_signal: NULL control block");
    s.ctrl[d] = static_cast<unsigned long long*>(ctrl[d]);
  }
  signal_kernel<<<1, 32, 0, (cudaStream_t)stream_>>>(s);
  PRL_LAUNCH_CHECK();
  return PRL_OK;
}
