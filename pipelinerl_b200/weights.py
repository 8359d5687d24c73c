"""In-flight weight update (hot path 3): learner -> samplers over NVLink P2P.

Reference boundary kept: `WeightUpdateManager(...).send_weight_update(version)` with the side effects
that every sampler serves the new weights and `WeightUpdateSuccess(version)` is appended to the
`weight_update_request` topic (pipelinerl/finetune_loop.py:147-171, 205-292; consumers:
pipelinerl/state.py:41-47).  Mechanism (csrc/weight_push.cu): samplers export CUDA-IPC handles of
two arena buffers + a control block once; an update is one copy kernel per learner rank into each
sampler's INACTIVE buffer plus a system-scope signal; the sampler flips buffers between two token
steps.  No HTTP, no per-tensor loop, no pause of in-flight sequences.
"""
from __future__ import annotations

import ctypes as C
import time
from dataclasses import dataclass
from typing import Literal

import torch
from pydantic import BaseModel

from . import _lib
from .model import ModelConfig, ParamArena

TRAINER_TOPIC = "weight_update_request"


# ---- trainer -> world messages (same kinds/fields as finetune_loop.py:141-171) ------------------
class ParameterInfo(BaseModel):
    name: str
    shape: list[int]
    dtype: str


class WeightUpdateRequest(BaseModel):
    kind: Literal["weight_update_request"] = "weight_update_request"
    version: int
    parameters_info: list[ParameterInfo] = []
    timestamp: float = 0.0


class WeightUpdateSuccess(BaseModel):
    kind: Literal["weight_update_success"] = "weight_update_success"
    version: int
    timestamp: float = 0.0


class SamplesProcessed(BaseModel):
    kind: Literal["samples_processed"] = "samples_processed"
    samples_processed: int
    timestamp: float = 0.0


class TrainingDone(BaseModel):
    kind: Literal["training_done"] = "training_done"
    timestamp: float = 0.0


# ---- device buffers that can be shared across processes --------------------------------------------
class _RawCudaBuffer:
    """cudaMalloc'ed memory (IPC-exportable, unlike a slice of torch's caching allocator) exposed to torch
    through __cuda_array_interface__."""

    def __init__(self, ptr: int, nbytes: int, owner: bool, opened: bool = False):
        self.ptr, self.nbytes, self.owner, self.opened = ptr, nbytes, owner, opened
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3}

    def tensor(self, dtype: torch.dtype, device) -> torch.Tensor:
        t = torch.as_tensor(self, device=device)
        return t.view(dtype)

    def release(self) -> None:
        lib = _lib.load()
        if self.owner and self.ptr:
            lib.prl_ipc_free(self.ptr)
        elif self.opened and self.ptr:
            lib.prl_ipc_close(self.ptr)
        self.ptr = 0


def ipc_alloc(nbytes: int) -> _RawCudaBuffer:
    lib = _lib.load()
    p = C.c_void_p()
    _lib.check(lib.prl_ipc_alloc(nbytes, C.byref(p)))
    return _RawCudaBuffer(int(p.value), nbytes, owner=True)


def ipc_export(buf: _RawCudaBuffer) -> bytes:
    lib = _lib.load()
    h = C.create_string_buffer(64)
    _lib.check(lib.prl_ipc_export(buf.ptr, h))
    return h.raw


def ipc_open(handle: bytes, nbytes: int) -> _RawCudaBuffer:
    lib = _lib.load()
    p = C.c_void_p()
    _lib.check(lib.prl_ipc_open(C.create_string_buffer(handle, 64), C.byref(p)))
    return _RawCudaBuffer(int(p.value), nbytes, owner=False, opened=True)


@dataclass
class SamplerHandles:
    """What a sampler publishes once (replaces the TCPStore/PyNccl rendezvous of torch_utils.py:70-94)."""
    arena: tuple[bytes, bytes]
    ctrl: bytes
    nbytes: int
    device_index: int


def push_slice(nbytes: int, rank: int, n_learners: int) -> tuple[int, int]:
    """(offset, length) in bytes of the arena slice learner `rank` pushes; 16-byte aligned, exact cover."""
    n16 = nbytes // 16
    lo = n16 * rank // n_learners * 16
    hi = n16 * (rank + 1) // n_learners * 16 if rank + 1 < n_learners else nbytes
    return lo, hi - lo


class WeightReceiver:
    """Sampler side: two arena buffers, a control block, and the flip at a token-step boundary."""

    def __init__(self, cfg: ModelConfig, device, n_pushers: int = 1):
        self.cfg, self.dev = cfg, torch.device(device)
        torch.cuda.set_device(self.dev)
        from .model import ArenaLayout
        total = ArenaLayout.build(cfg).total
        self.nbytes = total * 2
        self._bufs = [ipc_alloc(self.nbytes), ipc_alloc(self.nbytes)]
        self._ctrl_buf = ipc_alloc(32)
        self.arenas = [ParamArena(cfg, self.dev, data=b.tensor(torch.bfloat16, self.dev)) for b in self._bufs]
        self.ctrl = self._ctrl_buf.tensor(torch.int64, self.dev)  # [version, arrivals, acked flips, -]
        self._ctrl_host = torch.zeros(4, dtype=torch.int64).pin_memory()
        self._poll_stream = torch.cuda.Stream(device=self.dev)
        self._poll_event: torch.cuda.Event | None = None
        self.active = 0
        self.n_pushers = n_pushers
        self._arrivals_seen = 0
        self.version = 0
        self.flips = 0
        self.last_flip_wall_s = 0.0

    @property
    def arena(self) -> ParamArena:
        return self.arenas[self.active]

    @property
    def inactive_index(self) -> int:
        return 1 - self.active

    def handles(self) -> SamplerHandles:
        return SamplerHandles((ipc_export(self._bufs[0]), ipc_export(self._bufs[1])), ipc_export(self._ctrl_buf),
                              self.nbytes, self.dev.index or 0)

    def maybe_flip(self, engine=None) -> bool:
        """Call between token steps.  Non-blocking: an 16-byte D2H copy on a side stream is polled; when all
        pushers have signalled, the engine switches to the freshly written buffer (its CUDA graph for that
        buffer is captured on first use).  In-flight sequences continue on their existing KV."""
        if self._poll_event is None:
            with torch.cuda.stream(self._poll_stream):
                self._ctrl_host.copy_(self.ctrl, non_blocking=True)
                self._poll_event = torch.cuda.Event()
                self._poll_event.record()
            return False
        if not self._poll_event.query():
            return False
        self._poll_event = None
        version, arrivals = int(self._ctrl_host[0]), int(self._ctrl_host[1])
        if arrivals - self._arrivals_seen < self.n_pushers:
            return False
        t0 = time.perf_counter()
        self._arrivals_seen += self.n_pushers
        self.active = 1 - self.active
        self.version = version
        self.arenas[self.active].version = version
        if engine is not None:
            engine.set_arena(self.arenas[self.active])
        self.flips += 1
        # acknowledge: the learner may now overwrite the buffer this sampler just stopped reading ... but only
        # after the token step that is possibly still in flight on it has finished (stream order)
        self.ctrl[2] = self.flips
        self.last_flip_wall_s = time.perf_counter() - t0
        return True

    def close(self) -> None:
        for b in self._bufs + [self._ctrl_buf]:
            b.release()


class WeightUpdateManager:
    """Learner side.  `samplers`: SamplerHandles of every sampler (or WeightReceiver objects when learner and
    sampler share a process, as in the single-GPU tests).  `rank`/`n_learners`: this learner pushes byte slice
    rank/n_learners of the arena to all samplers, so the push uses every learner GPU's NVLink egress."""

    def __init__(self, samplers: list, learner_arena: torch.Tensor, update_stream=None, rank: int = 0,
                 n_learners: int = 1, max_ctas: int = 0):
        self.lib = _lib.load()
        self.src = learner_arena
        self.update_stream = update_stream
        self.rank, self.n_learners, self.max_ctas = rank, n_learners, max_ctas
        self._opened: list[_RawCudaBuffer] = []
        self.peer_bufs: list[tuple[int, int]] = []
        self.peer_ctrl: list[int] = []
        nbytes = learner_arena.numel() * learner_arena.element_size()
        for s in samplers:
            if isinstance(s, WeightReceiver):
                if s.dev != learner_arena.device:
                    _lib.check(self.lib.prl_enable_peer_access(s.dev.index))
                self.peer_bufs.append((s._bufs[0].ptr, s._bufs[1].ptr))
                self.peer_ctrl.append(s._ctrl_buf.ptr)
                assert s.nbytes == nbytes
            else:
                assert s.nbytes == nbytes, "learner and sampler arenas must share one layout"
                a0, a1, c = ipc_open(s.arena[0], nbytes), ipc_open(s.arena[1], nbytes), ipc_open(s.ctrl, 32)
                self._opened += [a0, a1, c]
                self.peer_bufs.append((a0.ptr, a1.ptr))
                self.peer_ctrl.append(c.ptr)
        self.n = len(self.peer_bufs)
        self.target = [1] * self.n  # samplers start on buffer 0, so the first update lands in buffer 1
        self.sent = 0
        dev = learner_arena.device
        self._ctrl_views = [_RawCudaBuffer(p, 32, owner=False).tensor(torch.int64, dev) for p in self.peer_ctrl]
        self.offset, self.bytes = push_slice(nbytes, rank, n_learners)
        self.last_push_ms = 0.0

    def send_weight_update(self, version: int, stream: torch.cuda.Stream | None = None, wait: bool = True,
                           ack_timeout_s: float = 120.0) -> float:
        """Push this rank's slice to every sampler's inactive buffer and signal.  Returns device ms of the push."""
        st = stream or torch.cuda.current_stream()
        self.wait_for_acks(ack_timeout_s)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ms = 0.0
        with torch.cuda.stream(st):
            e0.record()
            for g0 in range(0, self.n, 8):
                idx = list(range(g0, min(self.n, g0 + 8)))
                dst = (C.c_void_p * len(idx))(*[self.peer_bufs[i][self.target[i]] for i in idx])
                ctrl = (C.c_void_p * len(idx))(*[self.peer_ctrl[i] for i in idx])
                _lib.check(self.lib.prl_weights_push(self.src.data_ptr(), dst, len(idx), self.offset, self.bytes,
                                                     self.max_ctas, int(st.cuda_stream)))
                _lib.check(self.lib.prl_weights_signal(ctrl, len(idx), version, int(st.cuda_stream)))
            e1.record()
        self.target = [1 - t for t in self.target]
        self.sent += 1
        if wait:
            e1.synchronize()
            ms = e0.elapsed_time(e1)
            self.last_push_ms = ms
        if self.update_stream is not None and self.rank == 0:
            self.update_stream.write(WeightUpdateSuccess(version=version, timestamp=time.time()))
        return ms

    def wait_for_acks(self, timeout_s: float = 120.0) -> None:
        """Update j overwrites the buffer the sampler was reading before its (j-1)-th flip: wait until every
        sampler has acknowledged j-1 flips (a P2P read of one word per sampler; normally already true because
        updates are an optimizer step apart)."""
        need = self.sent  # update j = sent + 1 needs flips 1 .. j-1 acknowledged
        if need <= 0:
            return
        t0 = time.time()
        for view in self._ctrl_views:
            while int(view[2].item()) < need:
                if time.time() - t0 > timeout_s:
                    raise TimeoutError(f"sampler did not acknowledge weight version flip {need}")
                time.sleep(0.001)

    def close(self) -> None:
        for b in self._opened:
            b.release()
