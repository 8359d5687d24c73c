#!/usr/bin/env python
"""All three hot paths at once on Qwen2.5-7B (BASELINE.json configs[1] shape, reduced to the GPUs at hand):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tools/pipeline_bench.py [--updates 3] [--context 8192] [--batch 64]

rank 0      learner: NativeQwen2 + FusedAdamW; every optimizer step (2 micro-batches x 16 384 tokens through rl_step
            -> native backward -> fused AdamW) is followed by ONE in-flight weight update: the bf16 shadow arena is
            pushed over NVLink into every sampler's inactive buffer and signalled (hot path 3)
ranks 1..   samplers: DecodeEngine token steps (64 sequences, `context`-token KV) that never pause; a sampler flips
            to the new weights between two token steps when the signal arrives

Reported (rank 0, one JSON line): trainer ms/step and tokens/s, push ms, sampler tokens/s over the whole window
(training + pushes going on), the per-update stall (slowest token step around a flip minus the median step), the
slow-down of token steps while a push is writing into the sampler's HBM, and that after the last update the sampler's
live arena is byte-identical to the learner's parameters.
"""
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
from pipelinerl_b200.engine import DecodeEngine  # noqa: E402
from pipelinerl_b200.model import ModelConfig  # noqa: E402
from pipelinerl_b200.weights import SamplerHandles, WeightReceiver, WeightUpdateManager  # noqa: E402


def checksum(flat_bf16: torch.Tensor):
    flat = flat_bf16.view(torch.int16)
    c0 = c1 = 0
    step = 1 << 26
    for s0 in range(0, flat.numel(), step):
        v = flat[s0:s0 + step].to(torch.int64)
        c0 += int(v.sum())
        c1 += int((v * (torch.arange(s0, s0 + v.numel(), device=flat.device) % 8191 + 1)).sum())
    return [c0, c1]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="7b")
    ap.add_argument("--updates", type=int, default=3)
    ap.add_argument("--context", type=int, default=8192)
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--tokens", type=int, default=16384)
    ap.add_argument("--micro", type=int, default=2)
    args = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    os.environ.setdefault("NCCL_DEBUG", "WARN")
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    dist.init_process_group("nccl", device_id=dev)
    assert world >= 2, "needs one learner and at least one sampler"
    cfg = ModelConfig.qwen2_5_7b() if args.model == "7b" else ModelConfig(
        vocab_size=1024, hidden_size=512, intermediate_size=1024, num_layers=2, num_q_heads=4, num_kv_heads=2)
    is_learner = rank == 0
    U = args.updates

    my = None
    if is_learner:
        import train_bench
        from pipelinerl_b200.finetune.optim import FusedAdamW
        from pipelinerl_b200.finetune.rl import RLConfig, rl_step
        from pipelinerl_b200.learner_model import NativeQwen2
        try:
            torch.cuda.memory._set_allocator_settings("expandable_segments:True")
        except Exception:  # noqa: BLE001
            pass
        model = NativeQwen2(cfg, dev)
        opt = FusedAdamW(model.named_parameters(), lr=1e-6, weight_decay=0.01, max_grad_norm=0.3, grad_dtype=torch.float32)
        model.bind(opt)
        rcfg = RLConfig(batch_size=args.micro)
        batches = [train_bench.synthetic_batch(cfg, args.tokens, 1, dev, 100 + i) for i in range(args.micro)]
        for b in batches:
            b.input_ids %= cfg.vocab_size
            b.labels = torch.where(b.labels >= 0, b.input_ids, b.labels)
    else:
        recv = WeightReceiver(cfg, dev, n_pushers=1)
        recv.arenas[0].init_random(seed=42)
        room = 8192
        eng = DecodeEngine(cfg, recv.arena, max_batch=args.batch, max_seq_len=args.context + room, max_new_tokens=room,
                           device=dev, prefill_chunk=0)
        B, mb = eng.B, eng.max_blocks
        eng.block_table.copy_(torch.arange(1, 1 + B * mb, dtype=torch.int32, device=dev).view(B, mb))
        eng.free_pages.clear()
        eng.prompt_len.fill_(args.context); eng.positions.fill_(args.context); eng.seq_lens.fill_(args.context + 1)
        eng.max_new_t.fill_(room); eng.active.fill_(1); eng.ignore_eos = True
        h = recv.handles()
        my = (h.arena, h.ctrl, h.nbytes, h.device_index)
    gathered = [None] * world
    dist.all_gather_object(gathered, my)

    if is_learner:
        handles = [SamplerHandles(*g) for g in gathered[1:]]
        mgr = WeightUpdateManager(handles, opt.shadow_bf16, rank=0, n_learners=1)
        dist.barrier()
        steps, pushes = [], []
        ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731
        # update 1 = the initial weights (also lets every sampler capture the CUDA graph of its second buffer),
        # updates 2 .. U+1 follow one optimizer step each
        pushes.append(mgr.send_weight_update(version=1))
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
            pushes.append(mgr.send_weight_update(version=u + 2))
        mgr.wait_for_acks()
        time.sleep(0.3)
        torch.cuda.synchronize()
        mine = {"step_ms": steps, "push_ms": pushes, "checksum": checksum(opt.shadow_bf16),
                "arena_bytes": int(opt.shadow_bf16.numel() * 2), "loss": float(loss)}
    else:
        dist.barrier()
        times, flips_at, t_wall0 = [], [], time.perf_counter()
        for _ in range(3):
            eng.step()
        torch.cuda.synchronize()
        t_start = time.perf_counter()
        extra = 0
        while len(times) < 200000:
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
        stalls = []
        for f in flips_at[1:]:                      # the first flip captures the second buffer's CUDA graph
            stalls.append(max(times[max(0, f - 1): f + 3]) - med)
        slow = sorted(times)[int(len(times) * 0.99)]
        torch.cuda.synchronize()
        mine = {"median_step_ms": med, "p99_step_ms": slow, "n_steps": len(times), "window_s": t_total,
                "tokens_per_s_over_window": len(times) * eng.B / t_total, "stall_ms": stalls,
                "first_flip_ms_incl_graph_capture": (max(times[flips_at[0]: flips_at[0] + 2]) - med) if flips_at else None,
                "version": recv.version, "checksum": checksum(recv.arena.data)}
    allr = [None] * world
    dist.all_gather_object(allr, mine)
    if rank == 0:
        L, S = allr[0], allr[1:]
        tok_step = args.micro * args.tokens
        step_ms = sorted(L["step_ms"])[len(L["step_ms"]) // 2]
        out = {"bench": "pipeline_all_hot_paths", "model": "Qwen2.5-7B" if args.model == "7b" else "tiny",
               "gpus": f"1 learner + {world - 1} sampler(s)", "updates": U,
               "learner": {"ms_per_optimizer_step": round(step_ms, 1), "trainer_tokens_per_s": round(tok_step / step_ms * 1e3, 1),
                           "tokens_per_step": tok_step, "step_ms_all": [round(x, 1) for x in L["step_ms"]],
                           "push_ms_all": [round(x, 2) for x in L["push_ms"]],
                           "push_GBs": round(L["arena_bytes"] * (world - 1) / (sorted(L["push_ms"])[len(L["push_ms"]) // 2] / 1e3) / 1e9, 1),
                           "final_loss": L["loss"]},
               "samplers": [{"median_step_ms": round(s["median_step_ms"], 3), "p99_step_ms": round(s["p99_step_ms"], 3),
                             "rollout_tokens_per_s_over_window": round(s["tokens_per_s_over_window"], 1),
                             "rollout_tokens_per_s_at_median_step": round(args.batch / s["median_step_ms"] * 1e3, 1),
                             "token_steps": s["n_steps"], "stall_ms_per_update": [round(x, 3) for x in s["stall_ms"]],
                             "first_flip_ms_incl_graph_capture": s["first_flip_ms_incl_graph_capture"],
                             "weight_version": s["version"]} for s in S],
               "sampler_arena_identical_to_learner": all(s["checksum"] == L["checksum"] for s in S),
               "context": args.context, "batch": args.batch}
        print(json.dumps(out), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
