#!/usr/bin/env python
"""Weight-update stall measurement (hot path 3) on one multi-GPU box.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tools/push_bench.py --learners Ng [--model 7b|tiny] [--updates 5]

Ranks [0, Ng) are learners (each holds the full bf16 arena), ranks [Ng, N) are samplers running a
DecodeEngine token-step loop.  Measured:
  * ours      : one-shot P2P push (each learner rank pushes 1/Ng of the bytes to every sampler's inactive
                buffer) + flip at a token-step boundary.  push_ms = device time on the learner (max over
                ranks); stall_ms = how much longer the sampler's slowest step around the flip was than its
                median step (the sampler never pauses).
  * baseline A: the reference's mechanism (finetune_loop.py:279-286 / vllm1.py:110-127): sampler paused,
                one NCCL broadcast per HF parameter tensor (339 for Qwen2.5-7B) into a temporary, then a copy
                into the fused parameter (load_weights); stall_ms = pause -> resume wall time.
Rank 0 prints one JSON line.  NVLink-bound: algorithmic bytes = arena bytes per sampler.
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

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from pipelinerl_b200.engine import DecodeEngine  # noqa: E402
from pipelinerl_b200.model import ModelConfig, ParamArena  # noqa: E402
from pipelinerl_b200.weights import SamplerHandles, WeightReceiver, WeightUpdateManager  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--learners", type=int, default=1)
    ap.add_argument("--model", default="7b")
    ap.add_argument("--updates", type=int, default=5)
    ap.add_argument("--context", type=int, default=2048)
    ap.add_argument("--batch", type=int, default=64)
    args = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    dist.init_process_group("nccl", device_id=dev)
    ng = args.learners
    n_s = world - ng
    assert n_s >= 1 and ng >= 1
    cfg = ModelConfig.qwen2_5_7b() if args.model == "7b" else ModelConfig.tiny()
    is_learner = rank < ng

    # ---- set-up: samplers publish IPC handles, learners open them ----
    recv = eng = None
    if is_learner:
        arena = ParamArena(cfg, dev).init_random(seed=42)
        my = None
    else:
        recv = WeightReceiver(cfg, dev, n_pushers=ng)
        recv.arenas[0].init_random(seed=42)
        room = 4096
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
    mgr = None
    if is_learner:
        handles = [SamplerHandles(*g) for g in gathered[ng:]]
        mgr = WeightUpdateManager(handles, arena.data, rank=rank, n_learners=ng)
    nbytes = arena.nbytes() if is_learner else recv.nbytes

    def sampler_loop(stop_after_flips: int, use_flip: bool):
        """Run token steps until `stop_after_flips` weight flips happened; per-step wall times."""
        times, flips_at = [], []
        for _ in range(3):
            eng.step()
        torch.cuda.synchronize()
        while len(flips_at) < stop_after_flips and len(times) < 20000:
            t0 = time.perf_counter()
            if use_flip and recv.maybe_flip(eng):
                flips_at.append(len(times))
            eng.step()
            torch.cuda.current_stream().synchronize()
            times.append((time.perf_counter() - t0) * 1e3)
        return times, flips_at

    # ---- ours ----
    dist.barrier()
    if is_learner:
        pushes = []
        for u in range(args.updates + 1):
            time.sleep(0.15)
            arena.data[:1024].add_(1)  # new weights every update
            ms = mgr.send_weight_update(version=u + 1)
            pushes.append(ms)
        ours = {"push_ms": pushes[1:]}
    else:
        # first flip captures the second buffer's CUDA graph; measure the following ones
        times, flips_at = sampler_loop(args.updates + 1, True)
        med = sorted(times)[len(times) // 2]
        stalls = []
        for f in flips_at[1:]:
            w = times[max(0, f - 1): f + 3]
            stalls.append(max(w) - med)
        ours = {"median_step_ms": med, "stall_ms": stalls, "first_flip_ms_incl_graph_capture":
                (max(times[flips_at[0]: flips_at[0] + 2]) - med) if flips_at else None, "version": recv.version}
    # byte-exactness across GPUs/processes: the sampler's live arena must hash like the learner's
    torch.cuda.synchronize()
    live = arena.data if is_learner else recv.arena.data
    flat = live.view(torch.int16)
    c0 = c1 = 0
    step = 1 << 26
    for s0 in range(0, flat.numel(), step):
        v = flat[s0:s0 + step].to(torch.int64)
        c0 += int(v.sum())
        c1 += int((v * (torch.arange(s0, s0 + v.numel(), device=dev) % 8191 + 1)).sum())
    ours["checksum"] = [c0, c1]
    all_ours = [None] * world
    dist.all_gather_object(all_ours, ours)

    # ---- baseline A: per-tensor NCCL broadcast, sampler paused ----
    dist.barrier()
    src_arena = arena if is_learner else recv.arena
    names = list(src_arena.hf_state_dict())
    group_ranks = [0] + list(range(ng, world))
    grp = dist.new_group(group_ranks, backend="nccl")
    base = None
    if rank in group_ranks:
        sd = src_arena.hf_state_dict()
        torch.cuda.synchronize()
        dist.barrier(group=grp)
        t0 = time.perf_counter()
        for name in names:
            t = sd[name]
            if rank == 0:
                dist.broadcast(t.contiguous(), src=0, group=grp)
            else:
                buf = torch.empty(t.shape, dtype=t.dtype, device=dev)       # vllm1.py:120
                dist.broadcast(buf, src=0, group=grp)                       # vllm1.py:121
                t.copy_(buf)                                                # load_weights, vllm1.py:122
        torch.cuda.synchronize()
        base = {"stall_ms": (time.perf_counter() - t0) * 1e3, "tensors": len(names)}
    all_base = [None] * world
    dist.all_gather_object(all_base, base)

    if rank == 0:
        push = [max(o["push_ms"][i] for o in all_ours[:ng]) for i in range(args.updates)]
        stall = [s for o in all_ours[ng:] for s in o["stall_ms"]]
        med_step = [o["median_step_ms"] for o in all_ours[ng:]]
        push_ms = sorted(push)[len(push) // 2]
        out = {"bench": "weight_update", "model": args.model, "learners": ng, "samplers": n_s, "arena_bytes": nbytes,
               "ours": {"push_ms_median": round(push_ms, 3), "push_ms_all": [round(p, 3) for p in push],
                        "egress_GBs_per_learner": round(nbytes / ng * n_s / (push_ms / 1e3) / 1e9, 1),
                        "stall_ms_max": round(max(stall), 3) if stall else None,
                        "stall_ms_median": round(sorted(stall)[len(stall) // 2], 3) if stall else None,
                        "sampler_median_step_ms": [round(m, 3) for m in med_step],
                        "first_flip_ms_incl_graph_capture": [o["first_flip_ms_incl_graph_capture"] for o in all_ours[ng:]]},
               "baseline_per_tensor_nccl_paused": {"stall_ms": [round(b["stall_ms"], 2) for b in all_base if b][1:],
                                                   "tensors": all_base[0]["tensors"]},
               "bytes_identical_on_all_ranks": len({tuple(o["checksum"]) for o in all_ours}) == 1,
               "nvlink_peak_GBs_per_direction": 770.0}
        print(json.dumps(out), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
