#!/usr/bin/env python
"""The actor-learner split of BASELINE.json configs[1]/[2] on one box: `--learners M` learner GPUs (data-parallel
NativeQwen2, gradients exchanged by ShardedFusedAdamW over NVLink peer memory) + the remaining ranks as samplers
(DecodeEngine token steps that never pause), one in-flight weight update after EVERY optimizer step with the arena's byte
range split across the learner ranks (every learner pushes its slice to every sampler).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tools/split_bench.py --learners M [--updates 2] [--context 8192] [--batch 64]

`run_split()` is also what bench.py calls under torchrun (components.pipeline), so the numbers land in the driver's
BENCH / SCALE records.  Reported by rank 0:
  rollout tokens/s summed over the samplers WHILE the learners train and push, trainer tokens/s and optimizer steps/s
  (batch stated), DP exchange ms (reduce-scatter / update+all-gather phases), push ms (max over learners), stall ms per
  update (slowest token step around a flip minus the median step, max over samplers), `bytes_identical` (every sampler's
  live arena == every learner's bf16 parameters after the last update) and `dp_equals_single` (tiny-model check run on
  the learner ranks in the same job: the sharded exchange reproduces the single-learner step)."""
from __future__ import annotations

import argparse
import json
import os
import sys
import time
from pathlib import Path

import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tools"))


def checksum(flat_bf16: torch.Tensor):
    flat = flat_bf16.view(torch.int16)
    c0 = c1 = 0
    step = 1 << 26
    for s0 in range(0, flat.numel(), step):
        v = flat[s0:s0 + step].to(torch.int64)
        c0 += int(v.sum())
        c1 += int((v * (torch.arange(s0, s0 + v.numel(), device=flat.device) % 8191 + 1)).sum())
    return [c0, c1]


def run_split(n_learners: int, updates: int = 2, context: int = 8192, batch: int = 64, tokens: int = 16384,
              micro: int = 2, model_name: str = "7b", max_wall_s: float = 200.0, check_dp: bool = True):
    """All ranks of the (already initialised, NCCL) job call this; rank 0 gets the result dict, the others None."""
    from pipelinerl_b200.engine import DecodeEngine
    from pipelinerl_b200.model import ModelConfig
    from pipelinerl_b200.weights import SamplerHandles, WeightReceiver, WeightUpdateManager
    rank, world = dist.get_rank(), dist.get_world_size()
    dev = torch.device(f"cuda:{int(os.environ.get('LOCAL_RANK', '0'))}")
    torch.cuda.set_device(dev)
    n_samplers = world - n_learners
    assert n_learners >= 1 and n_samplers >= 1
    cfg = ModelConfig.qwen2_5_7b(fp32_head=True) if model_name == "7b" else ModelConfig(
        vocab_size=1024, hidden_size=512, intermediate_size=1024, num_layers=2, num_q_heads=4, num_kv_heads=2, fp32_head=True)
    is_learner = rank < n_learners
    lgroup = dist.new_group(ranks=list(range(n_learners)))          # every rank must take part in group creation
    U = updates
    dp_check = None
    if is_learner and n_learners > 1 and check_dp:
        import train_bench
        try:
            dp_check = train_bench.check_dp(group=lgroup, own_process_group=False)
        except Exception as e:  # noqa: BLE001
            dp_check = {"ok": False, "error": f"{type(e).__name__}: {str(e)[:200]}"}

    my = None
    if is_learner:
        import train_bench
        from pipelinerl_b200.finetune.optim import FusedAdamW, ShardedFusedAdamW
        from pipelinerl_b200.finetune.rl import RLConfig, rl_step
        from pipelinerl_b200.learner_model import NativeQwen2
        try:
            torch.cuda.memory._set_allocator_settings("expandable_segments:True")
        except Exception:  # noqa: BLE001
            pass
        model = NativeQwen2(cfg, dev)
        if n_learners > 1:
            opt = ShardedFusedAdamW(model.named_parameters(), lr=1e-6, weight_decay=0.01, max_grad_norm=0.3,
                                    grad_accum_fp32=True, group=lgroup, **model.optimizer_kwargs())
        else:
            opt = FusedAdamW(model.named_parameters(), lr=1e-6, weight_decay=0.01, max_grad_norm=0.3, grad_dtype=torch.float32,
                         **model.optimizer_kwargs())
        model.bind(opt)
        # spend the HBM left after parameters / optimizer state (sharded over the learners) on kept gate_up outputs: every kept
        # layer skips the largest recompute GEMM of the backward (same rule as tools/train_bench.py, 20 GB of headroom)
        torch.cuda.synchronize()
        torch.cuda.empty_cache()
        free = torch.cuda.mem_get_info(dev)[0]
        kept_bytes = tokens * 2 * (cfg.num_layers * (cfg.qkv_size + cfg.q_size + cfg.hidden_size) + cfg.num_layers * cfg.hidden_size)
        per_layer = tokens * 2 * cfg.intermediate_size * 2
        model.body.keep_gate_up_layers = int(max(0, min(cfg.num_layers, (free - 20e9 - kept_bytes) // per_layer)))
        rcfg = RLConfig(batch_size=micro * n_learners)
        batches = [train_bench.synthetic_batch(cfg, tokens, 1, dev, 100 + rank * micro + i) for i in range(micro)]
        for b in batches:
            b.input_ids %= cfg.vocab_size
            b.labels = torch.where(b.labels >= 0, b.input_ids, b.labels)
    else:
        recv = WeightReceiver(cfg, dev, n_pushers=n_learners)
        recv.arenas[0].init_random(seed=42)
        room = 8192
        eng = DecodeEngine(cfg, recv.arena, max_batch=batch, max_seq_len=context + room, max_new_tokens=room,
                           device=dev, prefill_chunk=0)
        B, mb = eng.B, eng.max_blocks
        eng.block_table.copy_(torch.arange(1, 1 + B * mb, dtype=torch.int32, device=dev).view(B, mb))
        eng.free_pages.clear()
        eng.prompt_len.fill_(context); eng.positions.fill_(context); eng.seq_lens.fill_(context + 1)
        eng.max_new_t.fill_(room); eng.active.fill_(1); eng.ignore_eos = True
        h = recv.handles()
        my = (h.arena, h.ctrl, h.nbytes, h.device_index)
    gathered = [None] * world
    dist.all_gather_object(gathered, my)

    if is_learner:
        handles = [SamplerHandles(*g) for g in gathered[n_learners:]]
        mgr = WeightUpdateManager(handles, opt.shadow_bf16, rank=rank, n_learners=n_learners)
        dist.barrier()
        steps, pushes, exch = [], [], []
        ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731
        # update 1 = the initial weights (lets every sampler capture the CUDA graph of its second buffer); updates
        # 2 .. U+1 follow one optimizer step each
        pushes.append(mgr.send_weight_update(version=1))
        loss = torch.zeros(())
        for u in range(U):
            e0, e1 = ev(), ev()
            opt.zero_grad()
            e0.record()
            for b in batches:
                loss, _ = rl_step(model, b, u, 1000, rcfg)
                loss.backward()
            opt.step()
            model.after_optimizer_step()
            e1.record()
            torch.cuda.synchronize()
            steps.append(e0.elapsed_time(e1))
            exch.append(list(getattr(opt, "last_phase_ms", (0.0, 0.0))))
            pushes.append(mgr.send_weight_update(version=u + 2))
        mgr.wait_for_acks()
        time.sleep(0.3)
        torch.cuda.synchronize()
        mine = {"role": "learner", "step_ms": steps, "push_ms": pushes, "exchange_ms": exch,
                "checksum": checksum(opt.shadow_bf16), "arena_bytes": int(opt.shadow_bf16.numel() * 2),
                "push_bytes": int(mgr.bytes) * n_samplers, "loss": float(loss), "dp_check": dp_check,
                "peak_memory_GB": round(torch.cuda.max_memory_allocated() / 1e9, 1),
                "keep_gate_up_layers": model.body.keep_gate_up_layers}
    else:
        dist.barrier()
        times, flips_at = [], []
        for _ in range(3):
            eng.step()
        torch.cuda.synchronize()
        t_start = time.perf_counter()
        extra = 0
        while time.perf_counter() - t_start < max_wall_s:
            t0 = time.perf_counter()
            if recv.maybe_flip(eng):
                flips_at.append(len(times))
            eng.step()
            torch.cuda.current_stream().synchronize()
            times.append((time.perf_counter() - t0) * 1e3)
            if len(flips_at) >= U + 1:
                extra += 1
                if extra > 50:
                    break
        t_total = time.perf_counter() - t_start
        med = sorted(times)[len(times) // 2]
        stalls = [max(times[max(0, f - 1): f + 3]) - med for f in flips_at[1:]]   # flip 1 captures the 2nd buffer's graph
        torch.cuda.synchronize()
        mine = {"role": "sampler", "median_step_ms": med, "p99_step_ms": sorted(times)[int(len(times) * 0.99)],
                "n_steps": len(times), "window_s": t_total, "tokens_per_s_over_window": len(times) * eng.B / t_total,
                "stall_ms": stalls, "flips": len(flips_at), "version": recv.version, "checksum": checksum(recv.arena.data)}
    allr = [None] * world
    dist.all_gather_object(allr, mine)
    out = None
    if rank == 0:
        L, S = allr[:n_learners], allr[n_learners:]
        tok_step = micro * tokens * n_learners
        step_ms = max(sorted(l["step_ms"])[len(l["step_ms"]) // 2] for l in L)
        push_med = [sorted(l["push_ms"])[len(l["push_ms"]) // 2] for l in L]
        push_ms = max(push_med)
        all_stalls = [x for s in S for x in s["stall_ms"]]
        exch = [e for e in L[0]["exchange_ms"]]
        dpc = L[0]["dp_check"]
        out = {"split": f"{n_samplers} inference + {n_learners} learner GPUs", "n_samplers": n_samplers, "n_learners": n_learners,
               "model": "Qwen2.5-7B random-init" if model_name == "7b" else "tiny", "weight_updates": U,
               "update_interval_steps": 1,
               "rollout_tokens_per_s_while_training": round(sum(s["tokens_per_s_over_window"] for s in S), 1),
               "rollout_tokens_per_s_at_median_step": round(sum(batch / s["median_step_ms"] * 1e3 for s in S), 1),
               "sampler_median_step_ms": round(max(s["median_step_ms"] for s in S), 3),
               "sampler_p99_step_ms": round(max(s["p99_step_ms"] for s in S), 3),
               "trainer_ms_per_optimizer_step": round(step_ms, 1),
               "optimizer_steps_per_s": round(1e3 / step_ms, 4),
               "samples_per_optimizer_step": micro * n_learners, "tokens_per_optimizer_step": tok_step,
               "trainer_tokens_per_s": round(tok_step / step_ms * 1e3, 1),
               "dp_exchange_ms": {"reduce_scatter": round(max(e[0] for e in exch), 2) if exch else 0.0,
                                  "update_allgather": round(max(e[1] for e in exch), 2) if exch else 0.0},
               "push_ms": round(push_ms, 2), "push_ms_all_updates": [[round(x, 2) for x in l["push_ms"]] for l in L],
               "push_GBs_per_learner": round(L[0]["push_bytes"] / (push_med[0] / 1e3) / 1e9, 1),
               "stall_ms": round(max(all_stalls), 3) if all_stalls else None,
               "stall_ms_per_update": [[round(x, 3) for x in s["stall_ms"]] for s in S],
               "sampler_flips": [s["flips"] for s in S], "sampler_weight_version": [s["version"] for s in S],
               "bytes_identical": bool(all(s["checksum"] == L[0]["checksum"] for s in S)
                                       and all(l["checksum"] == L[0]["checksum"] for l in L)),
               "dp_equals_single": (bool(dpc["ok"]) if dpc is not None else None), "dp_check": dpc,
               "lm_head": "fp32-equivalent on both sides (hi + lo bf16 streams; the push carries both)",
               "learner_peak_memory_GB": max(l["peak_memory_GB"] for l in L),
               "learner_keep_gate_up_layers": [l["keep_gate_up_layers"] for l in L], "context": context, "batch_per_sampler": batch,
               "final_loss": L[0]["loss"]}
    # release IPC mappings / big buffers before the caller goes on
    dist.barrier()
    if is_learner:
        mgr.close()
        if hasattr(opt, "close"):
            opt.close()
        del model, opt, batches, mgr
    else:
        del eng
        recv.close()
    import gc
    gc.collect()
    torch.cuda.empty_cache()
    dist.barrier()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--learners", type=int, default=1)
    ap.add_argument("--model", default="7b")
    ap.add_argument("--updates", type=int, default=2)
    ap.add_argument("--context", type=int, default=8192)
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--tokens", type=int, default=16384)
    ap.add_argument("--micro", type=int, default=2)
    a = ap.parse_args()
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    os.environ.setdefault("NCCL_DEBUG", "WARN")
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    out = run_split(a.learners, a.updates, a.context, a.batch, a.tokens, a.micro, a.model)
    if out is not None:
        print(json.dumps(out), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
