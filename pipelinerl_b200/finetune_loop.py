"""Trainer stage: micro-batches -> rl_step -> backward -> (boundary) fused AdamW -> weight update.

Mirrors the control flow of pipelinerl/finetune_loop.py::rl_finetuning_worker (:562-1097) for one learner
rank with the pieces that matter to the hot path:
  * micro-batches accumulate until `samples_per_step` samples were seen (:674-713); sentinel batches
    contribute loss*0 (:784-786);
  * at the boundary: optimizer step with gradient clipping (:727-744), then `send_weight_update(version)`
    every `weight_update_interval` steps with version = samples trained on (:936-949);
  * `SamplesProcessed` / `WeightUpdateSuccess` / `TrainingDone` messages on topic `weight_update_request`.
The model is any torch module (rl_step's contract); the optimizer is FusedAdamW; no Accelerate/DeepSpeed.
"""
from __future__ import annotations

import time
from dataclasses import dataclass, field
from typing import Iterable

import torch

from .finetune.optim import FusedAdamW
from .finetune.rl import RLConfig, rl_step
from .finetune.types import PipelineBatchEncoding, TrainingMetrics
from .weights import SamplesProcessed, TrainingDone, WeightUpdateManager


@dataclass
class TrainerConfig:
    samples_per_step: int = 8              # train_batch_size * gradient_accumulation_passes (finetune_loop.py:629-631)
    learning_rate: float = 1e-6
    weight_decay: float = 0.01
    gradient_clipping_threshold: float | None = 0.3
    max_train_steps: int = 10
    # learning-rate schedule (finetune_loop.py:394-399).  The reference's production values are "cosine" with 50
    # warm-up steps over max_train_steps = 100000 (conf/finetune/base.yaml:41-53); the default here is the constant rate
    # because this dataclass's default max_train_steps is a 10-step smoke value
    lr_scheduler_type: str = "constant"
    num_warmup_steps: int = 0
    weight_update_interval: int = 1
    seq_parallel: int = 1                  # consecutive learner ranks sharing one packed row (finetune_loop.py:507-517)
    eos_token_id: int = 2                  # only used by the GPU-resident preprocess ("eos in input_ids" overflow rule)
    rl: RLConfig = field(default_factory=RLConfig)


def allreduce_gradients(flat_grad: torch.Tensor, group=None) -> None:
    """Learner data parallelism: ONE sum all-reduce of the flat gradient arena per optimizer step — the only
    collective on the hot path (the reference's DDP/ZeRO traffic at finetune_loop.py:716-755).
    A plain SUM is exact because every rank's loss is already normalised by the GLOBAL samples-per-step
    (rl/__init__.py:250, finetune_loop.py:644-646).  NCCL over NVLink on GPUs, gloo in the CPU tests."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM, group=group)


def build_seq_parallel_group(dp_group, seq_parallel: int):
    """Groups of `seq_parallel` CONSECUTIVE learner ranks (the reference's loop, finetune_loop.py:508-516); every learner
    rank calls this and gets the group it belongs to."""
    import torch.distributed as dist
    world = dist.get_world_size(dp_group)
    if world % seq_parallel != 0:
        raise ValueError(f"{world} learner ranks are not a multiple of seq_parallel={seq_parallel}")
    global_ranks = dist.get_process_group_ranks(dp_group) if dp_group is not None else list(range(world))
    me, mine = dist.get_rank(), None
    for leader in range(0, world, seq_parallel):
        ranks = [global_ranks[leader + i] for i in range(seq_parallel)]
        if me in ranks:      # only the members create their group (use_local_synchronization): samplers need not join
            mine = dist.new_group(ranks=ranks, use_local_synchronization=True)
    assert mine is not None
    return mine


class StepAccountant:
    """Sample accounting of one learner rank (pipelinerl/finetune_loop.py:626-646, 674-713).

    `samples_per_step` is the GLOBAL batch of an optimizer step (and the loss normaliser, rl_config.batch_size);
    every lead trainer owns `samples_per_step / world` of it.  After each micro-batch the cumulative local counts are
    summed over the ranks (the reference all-gathers them, :707-709) and the optimizer step happens when the global
    count EQUALS the target: the writer cuts micro-batches at that boundary (preprocess.py:620-622) and feeds sentinel
    batches to ranks that already hold their share (:600-607, trainer side :674-676)."""

    def __init__(self, samples_per_step: int, group=None, device=None, start_samples: int = 0, seq_parallel: int = 1):
        import torch.distributed as dist
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if self.world > 1 else 0
        # sequence parallelism: `seq_parallel` consecutive ranks hold slices of the SAME micro-batch and each counts its
        # samples, so the summed count is `seq_parallel` times the truth (reference finetune_loop.py:628,709-712)
        self.seq_parallel = seq_parallel
        if self.world % seq_parallel != 0:
            raise ValueError(f"{self.world} learner ranks are not a multiple of seq_parallel={seq_parallel}")
        leads = self.world // seq_parallel
        if samples_per_step % leads != 0:
            raise ValueError(f"samples_per_step={samples_per_step} is not divisible by the {leads} lead learner ranks")
        self.samples_per_step = samples_per_step
        self.samples_per_lead_per_step = samples_per_step // leads
        self.start_samples = start_samples
        self.local_samples = 0
        self.target_local = self.samples_per_lead_per_step
        self.target_total = samples_per_step
        self.device = device if (self.world > 1 and dist.get_backend(group) == "nccl") else "cpu"

    def expects_sentinel(self) -> bool:
        return self.local_samples == self.target_local

    def observe(self, n_samples: int, sentinel: bool) -> tuple[int, bool]:
        """account one micro-batch; returns (global samples so far incl. start_samples, do_optimizer_step)"""
        if self.expects_sentinel() and not sentinel:
            raise RuntimeError("this rank already holds its share of the optimizer step: expected a sentinel batch "
                               "(the writer must cut micro-batches at the step boundary, MicroBatchDealer)")
        if not sentinel:
            self.local_samples += n_samples
        total = self.local_samples
        if self.world > 1:
            import torch.distributed as dist
            counts = torch.tensor([self.local_samples], dtype=torch.int64, device=self.device)
            dist.all_reduce(counts, op=dist.ReduceOp.SUM, group=self.group)
            total = int(counts.item())
            if total % self.seq_parallel != 0:
                raise RuntimeError("sample counts of a sequence-parallel group disagree (finetune_loop.py:711)")
            total //= self.seq_parallel
        if total > self.target_total:
            raise RuntimeError(f"micro-batch overshoots the optimizer step ({total} > {self.target_total} samples): "
                               "pack with samples_per_step (pack_micro_batches / MicroBatchDealer)")
        do_step = total == self.target_total
        if do_step:
            self.target_local += self.samples_per_lead_per_step
            self.target_total += self.samples_per_step
        return self.start_samples + total, do_step


def prefetch_batches(batches: Iterable, maxsize: int = 1, pin_memory: bool = False):
    """The trainer's loader thread (pipelinerl/finetune_loop.py:494-505): a background thread pulls the next
    micro-batch from the `training_data` stream (parsing / tensor construction / optional pinning happen there) while
    the GPU works on the current one; `Queue(maxsize=1)` bounds the look-ahead exactly as the reference does.  An
    exception raised by the source is re-raised in the consumer (the reference forwards it through the queue, :104,133).
    Usage: `run_training(model, prefetch_batches(batches), cfg, ...)`."""
    import queue
    import threading
    q: "queue.Queue" = queue.Queue(maxsize=maxsize)
    done = object()

    def work():
        try:
            for b in batches:
                if pin_memory and hasattr(b, "pin_memory"):
                    b = b.pin_memory()
                q.put(b)
            q.put(done)
        except BaseException as e:  # noqa: BLE001  (forwarded to the consumer)
            q.put(e)
    t = threading.Thread(target=work, name="trainer-loader", daemon=True)
    t.start()
    while True:
        item = q.get()
        if item is done:
            return
        if isinstance(item, BaseException):
            raise item
        yield item


def run_training(model: torch.nn.Module, batches: Iterable[PipelineBatchEncoding], cfg: TrainerConfig,
                 weight_manager: WeightUpdateManager | None = None, message_writer=None,
                 device: torch.device | str = "cuda:0", dp_group=None) -> tuple[TrainingMetrics, list[dict]]:
    dev = torch.device(device)
    import torch.distributed as dist
    sharded = dist.is_available() and dist.is_initialized() and dist.get_world_size(dp_group) > 1 and \
        all(p.dtype == torch.bfloat16 for _, p in model.named_parameters())
    if sharded:
        # learner DP: P2P reduce-scatter + AdamW shard + P2P all-gather in one exchange step (optimizer state / Ng)
        from .finetune.optim import ShardedFusedAdamW
        native = hasattr(model, "bind")   # learner_model.NativeQwen2: fp32 gradient accumulation in the arena
        opt = ShardedFusedAdamW(model.named_parameters(), lr=cfg.learning_rate, weight_decay=cfg.weight_decay,
                                max_grad_norm=cfg.gradient_clipping_threshold, group=dp_group, grad_accum_fp32=native,
                                **(model.optimizer_kwargs() if hasattr(model, "optimizer_kwargs") else {}))
        if native:
            model.bind(opt)
    else:
        native = hasattr(model, "bind")
        opt = FusedAdamW(model.named_parameters(), lr=cfg.learning_rate, weight_decay=cfg.weight_decay,
                         max_grad_norm=cfg.gradient_clipping_threshold,
                         grad_dtype=torch.float32 if native else None,
                         **(model.optimizer_kwargs() if hasattr(model, "optimizer_kwargs") else {}))
        if native:
            model.bind(opt)
    if weight_manager is not None:
        weight_manager.src = opt.shadow_bf16
    from .finetune.optim import get_scheduler
    lr_schedule = get_scheduler(cfg.lr_scheduler_type, opt, cfg.num_warmup_steps, cfg.max_train_steps)
    rl_cfg = cfg.rl.model_copy(update={"batch_size": cfg.samples_per_step})   # GLOBAL batch normalises the loss
    tm = TrainingMetrics()
    history: list[dict] = []
    sp_group = build_seq_parallel_group(dp_group, cfg.seq_parallel) if cfg.seq_parallel > 1 else None
    acct = StepAccountant(cfg.samples_per_step, group=dp_group, device=dev, start_samples=tm.samples,
                          seq_parallel=cfg.seq_parallel)
    rank = acct.rank
    step_stats, t_step = [], time.time()
    opt.zero_grad()
    gpu_pre = None
    for batch in batches:
        if isinstance(batch, (bytes, bytearray, memoryview)):
            # binary micro-batch record (records.py): RL columns and the packed row are built ON THIS GPU
            if gpu_pre is None:
                from .records import GpuPreprocessor
                gpu_pre = GpuPreprocessor(dev, cfg.eos_token_id, divide_advantage_by_std=cfg.rl.divide_advantage_by_std)
            batch = gpu_pre.pack(batch)
        elif isinstance(batch, dict):
            batch = PipelineBatchEncoding.from_dict(batch)        # JSON documents of the stream (sentinels, legacy rows)
        batch = batch.to_device(dev)
        n_samples = 0 if batch.sentinel else (int(batch.seq_boundaries.numel()) - 1 - (1 if batch.padding else 0)
                                               if batch.is_packed else batch.input_ids.shape[0])
        samples_so_far, do_optimizer_step = acct.observe(n_samples, bool(batch.sentinel))
        loss, stats = rl_step(model, batch, tm.completed_steps, cfg.max_train_steps, rl_cfg, seq_parallel_group=sp_group)
        if batch.sentinel:
            loss = loss * 0.0
        loss.backward()
        tm.tokens += int(batch.input_ids.numel())
        if not batch.sentinel:
            tm.passes += 1
            step_stats.append(stats)
        if message_writer is not None and rank == 0:
            message_writer.write(SamplesProcessed(samples_processed=samples_so_far, timestamp=time.time()))
        if not do_optimizer_step:
            continue
        tm.samples = samples_so_far
        if not sharded:
            allreduce_gradients(opt.grad, dp_group)   # fp32-parameter path: plain SUM all-reduce, full AdamW per rank
        grad_norm = opt.step()
        tm.lr = float(opt.param_groups[0]["lr"])
        lr_schedule.step()
        opt.zero_grad()
        if hasattr(model, "after_optimizer_step"):
            model.after_optimizer_step()
        tm.completed_steps += 1
        tm.grad_norm = float(grad_norm.item())
        tm.train_loss = sum(s.get("loss", 0.0) for s in step_stats)
        pushed_ms = None
        if weight_manager is not None and tm.completed_steps % cfg.weight_update_interval == 0:
            pushed_ms = weight_manager.send_weight_update(tm.samples)
            tm.last_broadcasted_version = tm.samples
        history.append({"step": tm.completed_steps, "loss": tm.train_loss, "grad_norm": tm.grad_norm,
                        "samples": tm.samples, "sec_per_step": time.time() - t_step, "push_ms": pushed_ms})
        step_stats, t_step = [], time.time()
        if tm.completed_steps >= cfg.max_train_steps:
            break
    if message_writer is not None and rank == 0:
        message_writer.write(TrainingDone(timestamp=time.time()))
    return tm, history
