"""Optimizer of the trainer: fused AdamW over a flat parameter arena (csrc/adamw.cu).

Interface of the reference (pipelinerl/finetune/optim.py): `get_grouped_params` decides which
tensors are exempt from weight decay (names containing "bias" or "LayerNorm.weight", :8-22) and
`get_optimizer("adamw_torch", model, lr, wd)` returns the optimizer (:25-29).  Here the returned
object owns ONE contiguous arena per state (fp32 master, exp_avg, exp_avg_sq, gradient, bf16
shadow), re-points every parameter's .data/.grad at views of it, and performs

    global grad-norm -> clip -> AdamW -> bf16 re-cast

in two kernel launches without a host sync.  The bf16 shadow arena is exactly what the weight push
(hot path 3) copies to the samplers.
"""
from __future__ import annotations

import ctypes as C
from typing import Iterable

import torch

from .. import _lib

NO_DECAY_DEFAULT = ("bias", "LayerNorm.weight")


def get_grouped_params(model, weight_decay: float, no_decay: Iterable[str] = NO_DECAY_DEFAULT):
    """Two torch-style param groups; kept for callers that build a stock optimizer themselves."""
    with_wd, without_wd = [], []
    for name, p in model.named_parameters():
        (without_wd if any(tag in name for tag in no_decay) else with_wd).append(p)
    return [{"params": with_wd, "weight_decay": weight_decay}, {"params": without_wd, "weight_decay": 0.0}]


def _align(n: int, a: int = 64) -> int:
    return (n + a - 1) // a * a


class FusedAdamW:
    """AdamW(lr, betas=(0.9, 0.999), eps=1e-8, weight_decay) == torch.optim.AdamW numerics, one arena.

    named_params: iterable of (name, Parameter) all on one CUDA device.  Parameters may be fp32
    (then they ARE the master weights and `shadow_bf16` is an extra output) or bf16 (then a fp32
    master arena is created from them, the DeepSpeed-bf16 arrangement of the reference).
    """

    def __init__(self, named_params, lr: float, weight_decay: float = 0.01, betas=(0.9, 0.999), eps: float = 1e-8,
                 max_grad_norm: float | None = None, no_decay: Iterable[str] = NO_DECAY_DEFAULT,
                 grad_dtype: torch.dtype | None = None, keep_lo_residual: bool = False, lo_tail_for: str | None = None):
        """lo_tail_for: name of ONE parameter (the lm_head) whose bf16 residual lo = bf16(master - bf16(master)) is kept in a
        tail of the bf16 arena, right where the sampler's arena layout has `<name>_lo` (model.fused_shapes with fp32_head):
        the pushed bytes then carry the fp32-equivalent head, and the learner's own head reads the same pair."""
        named = [(n, p) for n, p in named_params if p.requires_grad]
        if not named:
            raise ValueError("FusedAdamW: no trainable parameters")
        dev = named[0][1].device
        if dev.type != "cuda":
            raise RuntimeError("FusedAdamW needs CUDA parameters: pipelinerl_b200 has no CPU fallback")
        self.lib = _lib.load()
        self.names = [n for n, _ in named]
        self.params = [p for _, p in named]
        self.lr, self.weight_decay, self.betas, self.eps = lr, weight_decay, betas, eps
        self.max_grad_norm = max_grad_norm
        self.step_count = 0
        pdt = self.params[0].dtype
        if any(p.dtype != pdt for p in self.params):
            raise ValueError("FusedAdamW: mixed parameter dtypes")
        if pdt not in (torch.float32, torch.bfloat16):
            raise ValueError(f"FusedAdamW: unsupported parameter dtype {pdt}")
        gdt = grad_dtype or pdt
        # tensors start on 64-element boundaries so that 16-byte vector accesses never straddle two tensors
        offsets, at = [], 0
        for p in self.params:
            offsets.append(at)
            at = _align(at + p.numel())
        self.n = at
        self.offsets = offsets
        self.master = torch.zeros(self.n, dtype=torch.float32, device=dev)
        self.exp_avg = torch.zeros_like(self.master)
        self.exp_avg_sq = torch.zeros_like(self.master)
        self.grad = torch.zeros(self.n, dtype=gdt, device=dev)
        self.lo_index = self.names.index(lo_tail_for) if lo_tail_for else None
        tail = _align(self.params[self.lo_index].numel()) if self.lo_index is not None else 0
        self.shadow_bf16 = torch.zeros(self.n + tail, dtype=torch.bfloat16, device=dev)   # [parameters | lo tail]
        self.shadow_lo = torch.zeros(self.n, dtype=torch.bfloat16, device=dev) if keep_lo_residual else None
        with torch.no_grad():
            for p, off in zip(self.params, offsets):
                k = p.numel()
                self.master[off:off + k].copy_(p.detach().reshape(-1).float())
                self.shadow_bf16[off:off + k].copy_(p.detach().reshape(-1))
                home = self.master if pdt == torch.float32 else self.shadow_bf16
                p.data = home[off:off + k].view(p.shape)
                if gdt == pdt:
                    p.grad = self.grad[off:off + k].view(p.shape)
                # else (bf16 parameters, fp32 gradient arena): torch forbids a .grad of another dtype; producers
                # that accumulate in fp32 (learner_body.NativeBody) write through grad_view(name) instead
        table = offsets + [self.n]
        self.tensor_offsets = torch.tensor(table, dtype=torch.int64, device=dev)
        flags = [1 if any(tag in n for tag in no_decay) else 0 for n in self.names]
        self.tensor_no_decay = torch.tensor(flags, dtype=torch.uint8, device=dev)
        self.workspace = torch.zeros(int(self.lib.prl_adamw_workspace_bytes()), dtype=torch.uint8, device=dev)
        self.grad_norm = torch.zeros(1, dtype=torch.float32, device=dev)
        self.param_groups = [{"lr": lr, "weight_decay": weight_decay, "params": self.params}]
        self.lo_view = None
        if self.lo_index is not None:
            k = self.params[self.lo_index].numel()
            self.lo_view = self.shadow_bf16[self.n:self.n + k].view(self.params[self.lo_index].shape)
            self._refresh_lo()

    def _refresh_lo(self) -> None:
        if self.lo_index is None:
            return
        off, k = self.offsets[self.lo_index], self.params[self.lo_index].numel()
        dst = (C.c_void_p * 1)(self.lo_view.data_ptr())
        _lib.check(self.lib.prl_bf16_residual(self.master.data_ptr() + off * 4, k, dst, 1, _lib.stream_ptr()))

    def grad_view(self, name: str) -> torch.Tensor:
        i = self.names.index(name)
        p = self.params[i]
        return self.grad[self.offsets[i]:self.offsets[i] + p.numel()].view(p.shape)

    def grad_views(self) -> dict[str, torch.Tensor]:
        return {n: self.grad[o:o + p.numel()].view(p.shape) for n, p, o in zip(self.names, self.params, self.offsets)}

    # torch.optim-like surface used by the trainer loop
    def zero_grad(self, set_to_none: bool = False) -> None:
        self.grad.zero_()

    def step(self, grad_scale: float = 1.0) -> torch.Tensor:
        """One optimizer step.  Returns the (device) pre-clip gradient norm, as clip_grad_norm_ does."""
        self.step_count += 1
        lr = self.param_groups[0]["lr"]
        a = _lib.AdamwArgs()
        a.n = self.n
        a.master = self.master.data_ptr()
        a.exp_avg = self.exp_avg.data_ptr()
        a.exp_avg_sq = self.exp_avg_sq.data_ptr()
        a.grad = self.grad.data_ptr()
        a.grad_is_bf16 = int(self.grad.dtype == torch.bfloat16)
        a.param_bf16 = self.shadow_bf16.data_ptr()
        a.param_bf16_lo = self.shadow_lo.data_ptr() if self.shadow_lo is not None else None
        a.tensor_offsets = self.tensor_offsets.data_ptr()
        a.tensor_no_decay = self.tensor_no_decay.data_ptr()
        a.n_tensors = len(self.params)
        a.lr, a.beta1, a.beta2, a.eps = lr, self.betas[0], self.betas[1], self.eps
        a.weight_decay = self.weight_decay
        a.step = self.step_count
        a.max_grad_norm = float(self.max_grad_norm) if self.max_grad_norm else 0.0
        a.grad_scale = grad_scale
        _lib.check(self.lib.prl_adamw_step(C.byref(a), self.grad_norm.data_ptr(), self.workspace.data_ptr(),
                                           self.workspace.numel(), _lib.stream_ptr()))
        self._refresh_lo()
        return self.grad_norm

    def state_dict(self) -> dict:
        return {"step": self.step_count, "master": self.master, "exp_avg": self.exp_avg,
                "exp_avg_sq": self.exp_avg_sq, "names": self.names, "offsets": self.offsets}

    def load_state_dict(self, sd: dict) -> None:
        self.step_count = int(sd["step"])
        self.master.copy_(sd["master"])
        self.exp_avg.copy_(sd["exp_avg"])
        self.exp_avg_sq.copy_(sd["exp_avg_sq"])
        self.shadow_bf16[:self.n].copy_(self.master)
        self._refresh_lo()


def get_optimizer(name: str, model, learning_rate: float, weight_decay: float, **kw) -> FusedAdamW:
    if name != "adamw_torch":
        raise ValueError(f"Unknown optimizer: {name} (only the reference default adamw_torch is on the hot path)")
    return FusedAdamW(model.named_parameters(), lr=learning_rate, weight_decay=weight_decay, **kw)


class LRSchedule:
    """Learning-rate schedule of the trainer: `get_scheduler(args.lr_scheduler_type, optimizer, args.num_warmup_steps,
    args.max_train_steps)` at pipelinerl/finetune_loop.py:394-399, i.e. transformers' LambdaLR schedules (the reference's
    defaults: cosine, 50 warm-up steps, conf/finetune/base.yaml:41-43).  The factor of step t (t = number of `step()`
    calls so far) multiplies the base rate found in `optimizer.param_groups[*]["lr"]` at construction; the fused
    optimizers read `param_groups[0]["lr"]` at every step, so no kernel argument changes."""

    KINDS = ("constant", "constant_with_warmup", "linear", "cosine")

    def __init__(self, kind: str, optimizer, num_warmup_steps: int = 0, num_training_steps: int | None = None):
        if kind not in self.KINDS:
            raise ValueError(f"Unknown lr_scheduler_type {kind!r} (supported: {', '.join(self.KINDS)})")
        if kind in ("linear", "cosine") and num_training_steps is None:
            raise ValueError(f"{kind} schedule needs num_training_steps")
        self.kind, self.opt = kind, optimizer
        self.warmup, self.total = int(num_warmup_steps), num_training_steps
        self.base_lrs = [g["lr"] for g in optimizer.param_groups]
        self.last_step = 0
        self._apply()

    def factor(self, t: int) -> float:
        import math
        if self.kind == "constant":
            return 1.0
        if t < self.warmup:
            return t / max(1, self.warmup)
        if self.kind == "constant_with_warmup":
            return 1.0
        if self.kind == "linear":
            return max(0.0, (self.total - t) / max(1, self.total - self.warmup))
        progress = (t - self.warmup) / max(1, self.total - self.warmup)
        return max(0.0, 0.5 * (1.0 + math.cos(math.pi * progress)))   # num_cycles = 0.5

    def _apply(self) -> None:
        f = self.factor(self.last_step)
        for g, base in zip(self.opt.param_groups, self.base_lrs):
            g["lr"] = base * f

    def step(self) -> None:
        self.last_step += 1
        self._apply()

    def get_last_lr(self) -> list[float]:
        return [g["lr"] for g in self.opt.param_groups]


def get_scheduler(name: str, optimizer, num_warmup_steps: int = 0, num_training_steps: int | None = None) -> LRSchedule:
    return LRSchedule(name, optimizer, num_warmup_steps, num_training_steps)


class ShardedFusedAdamW:
    """Data-parallel learners: gradient reduce-scatter + AdamW on a 1/Ng shard + bf16 parameter all-gather as one
    exchange step over NVLink peer memory (csrc/adamw.cu: shard_reduce / shard_update), with the fp32 optimizer
    state sharded Ng ways.  One process per GPU; `torch.distributed` is used for the one-off exchange of CUDA-IPC
    handles and for the host barriers between the two phases — there is no NCCL collective on the data path.

    Parameters must be bf16 (the reference's `load_as_bf16: True` + DeepSpeed-bf16 arrangement): their .data are
    views of this rank's bf16 parameter arena, their .grad views of its bf16 gradient arena.
    """

    def __init__(self, named_params, lr: float, weight_decay: float = 0.01, betas=(0.9, 0.999), eps: float = 1e-8,
                 max_grad_norm: float | None = None, no_decay: Iterable[str] = NO_DECAY_DEFAULT, group=None,
                 grad_accum_fp32: bool = False, lo_tail_for: str | None = None):
        import torch.distributed as dist
        from ..weights import ipc_alloc, ipc_export, ipc_open
        if not (dist.is_available() and dist.is_initialized()):
            raise RuntimeError("ShardedFusedAdamW needs an initialised torch.distributed process group")
        named = [(n, p) for n, p in named_params if p.requires_grad]
        dev = named[0][1].device
        if dev.type != "cuda":
            raise RuntimeError("ShardedFusedAdamW needs CUDA parameters: pipelinerl_b200 has no CPU fallback")
        if any(p.dtype != torch.bfloat16 for _, p in named):
            raise ValueError("ShardedFusedAdamW: parameters must be bf16")
        self.lib, self.dist, self.group = _lib.load(), dist, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        if self.world > 8:
            raise ValueError("ShardedFusedAdamW: at most 8 learner ranks (one box)")
        self.names = [n for n, _ in named]
        self.params = [p for _, p in named]
        self.lr, self.weight_decay, self.betas, self.eps, self.max_grad_norm = lr, weight_decay, betas, eps, max_grad_norm
        self.step_count = 0
        offsets, at = [], 0
        for p in self.params:
            offsets.append(at)
            at = _align(at + p.numel())
        self.n, self.offsets = at, offsets
        per = _align((self.n + self.world - 1) // self.world)
        self.lo, self.hi = min(self.n, per * self.rank), min(self.n, per * (self.rank + 1))
        self.lo_index = self.names.index(lo_tail_for) if lo_tail_for else None
        self.tail = _align(self.params[self.lo_index].numel()) if self.lo_index is not None else 0
        # IPC-shareable arenas of this rank
        self._grad_buf, self._shadow_buf = ipc_alloc(self.n * 2), ipc_alloc((self.n + self.tail) * 2)
        self._norm_buf = ipc_alloc(8 * 8)
        self.grad = self._grad_buf.tensor(torch.bfloat16, dev)
        self.shadow_bf16 = self._shadow_buf.tensor(torch.bfloat16, dev)
        shard = max(self.hi - self.lo, 1)
        self.master = torch.zeros(shard, dtype=torch.float32, device=dev)
        self.exp_avg, self.exp_avg_sq = torch.zeros_like(self.master), torch.zeros_like(self.master)
        self.gsum = torch.zeros_like(self.master)
        # micro-batch accumulation in fp32 (learner_body.NativeBody writes here through grad_views()); the exchange
        # itself stays bf16: one cast pass per optimizer step fills the IPC gradient arena
        self.grad_f32 = torch.zeros(self.n, dtype=torch.float32, device=dev) if grad_accum_fp32 else None
        with torch.no_grad():
            for p, off in zip(self.params, offsets):
                k = p.numel()
                self.shadow_bf16[off:off + k].copy_(p.detach().reshape(-1))
                p.data = self.shadow_bf16[off:off + k].view(p.shape)
                p.grad = self.grad[off:off + k].view(p.shape)
            self.master[: self.hi - self.lo].copy_(self.shadow_bf16[self.lo:self.hi].float())
        self.tensor_offsets = torch.tensor(offsets + [self.n], dtype=torch.int64, device=dev)
        self.tensor_no_decay = torch.tensor([1 if any(t in n for t in no_decay) else 0 for n in self.names],
                                            dtype=torch.uint8, device=dev)
        self.workspace = torch.zeros(int(self.lib.prl_adamw_workspace_bytes()), dtype=torch.uint8, device=dev)
        self.grad_norm = torch.zeros(1, dtype=torch.float32, device=dev)
        # exchange handles (once) and map the peers' arenas
        mine = (ipc_export(self._grad_buf), ipc_export(self._shadow_buf), ipc_export(self._norm_buf), self.n)
        gathered = [None] * self.world
        dist.all_gather_object(gathered, mine, group=group)
        self._peer_bufs = []
        self._grads, self._shadows, self._norms = [], [], []
        for r, (hg, hs, hn, n_r) in enumerate(gathered):
            if n_r != self.n:
                raise RuntimeError("ShardedFusedAdamW: ranks disagree on the arena size")
            if r == self.rank:
                self._grads.append(self._grad_buf.ptr); self._shadows.append(self._shadow_buf.ptr); self._norms.append(self._norm_buf.ptr)
            else:
                g, s, nn = ipc_open(hg, self.n * 2), ipc_open(hs, (self.n + self.tail) * 2), ipc_open(hn, 64)
                self._peer_bufs += [g, s, nn]
                self._grads.append(g.ptr); self._shadows.append(s.ptr); self._norms.append(nn.ptr)
        self.param_groups = [{"lr": lr, "weight_decay": weight_decay, "params": self.params}]
        self.last_phase_ms = (0.0, 0.0)
        self.lo_view = None
        if self.lo_index is not None:
            k = self.params[self.lo_index].numel()
            self.lo_view = self.shadow_bf16[self.n:self.n + k].view(self.params[self.lo_index].shape)
            self._refresh_lo()
            self._barrier()

    def _refresh_lo(self) -> None:
        """this rank's part of the head's fp32 master -> bf16 residual in EVERY rank's arena tail (P2P stores)"""
        if self.lo_index is None:
            return
        off, k = self.offsets[self.lo_index], self.params[self.lo_index].numel()
        a, b = max(self.lo, off), min(self.hi, off + k)
        if b <= a:
            return
        dsts = (C.c_void_p * self.world)(*[self._shadows[r] + (self.n + a - off) * 2 for r in range(self.world)])
        _lib.check(self.lib.prl_bf16_residual(self.master.data_ptr() + (a - self.lo) * 4, b - a, dsts, self.world,
                                              _lib.stream_ptr()))

    def zero_grad(self, set_to_none: bool = False) -> None:
        (self.grad_f32 if self.grad_f32 is not None else self.grad).zero_()

    def grad_views(self) -> dict[str, torch.Tensor]:
        src = self.grad_f32 if self.grad_f32 is not None else self.grad
        return {n: src[o:o + p.numel()].view(p.shape) for n, p, o in zip(self.names, self.params, self.offsets)}

    def _args(self) -> _lib.AdamwShardArgs:
        a = _lib.AdamwShardArgs()
        a.n, a.shard_begin, a.shard_end = self.n, self.lo, self.hi
        a.master, a.exp_avg, a.exp_avg_sq = self.master.data_ptr(), self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr()
        for r in range(self.world):
            a.grads[r], a.shadows[r], a.norm_tables[r] = self._grads[r], self._shadows[r], self._norms[r]
        a.n_peers, a.rank, a.grad_is_bf16 = self.world, self.rank, 1
        a.gsum_scratch = self.gsum.data_ptr()
        a.tensor_offsets, a.tensor_no_decay, a.n_tensors = self.tensor_offsets.data_ptr(), self.tensor_no_decay.data_ptr(), len(self.params)
        a.lr, a.beta1, a.beta2, a.eps = self.param_groups[0]["lr"], self.betas[0], self.betas[1], self.eps
        a.weight_decay, a.step = self.weight_decay, self.step_count
        a.max_grad_norm = float(self.max_grad_norm) if self.max_grad_norm else 0.0
        a.grad_scale = 1.0
        return a

    def _barrier(self) -> None:
        torch.cuda.current_stream().synchronize()
        self.dist.barrier(group=self.group)

    def step(self, grad_scale: float = 1.0) -> torch.Tensor:
        self.step_count += 1
        a = self._args()
        a.grad_scale = grad_scale
        st = _lib.stream_ptr()
        if self.grad_f32 is not None:
            self.grad.copy_(self.grad_f32)
        e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        self._barrier()                                   # every rank's backward has written its gradient arena
        e[0].record()
        _lib.check(self.lib.prl_adamw_sharded_reduce(C.byref(a), self.workspace.data_ptr(), self.workspace.numel(), st))
        e[1].record()
        self._barrier()                                   # every rank's norm partial is in every norm table
        e[2].record()
        _lib.check(self.lib.prl_adamw_sharded_update(C.byref(a), self.grad_norm.data_ptr(), st))
        self._refresh_lo()
        e[3].record()
        self._barrier()                                   # every shard of every parameter arena is written
        self.last_phase_ms = (e[0].elapsed_time(e[1]), e[2].elapsed_time(e[3]))
        return self.grad_norm

    def close(self) -> None:
        for b in self._peer_bufs:
            b.release()
        for b in (self._grad_buf, self._shadow_buf, self._norm_buf):
            b.release()
