#!/usr/bin/env python
"""Tensor-parallel sampler engine (TP=2) on a 2-GPU box: parity check on a tiny model, or the token-step bench for
random-init Qwen2.5-32B (BASELINE config 4) with 64 sequences x 8192-token synthetic context.

    torchrun --nproc-per-node 2 tools/tp_bench.py --check
    torchrun --nproc-per-node 2 tools/tp_bench.py [--model 32b] [--steps 20]
"""
import argparse
import json
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from pipelinerl_b200.engine import SamplingParams  # noqa: E402
from pipelinerl_b200.model import ModelConfig, ParamArena, shard_fused_weights  # noqa: E402
from pipelinerl_b200.tp_engine import TPDecodeEngine  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--model", default="32b")
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--context", type=int, default=8192)
    ap.add_argument("--batch", type=int, default=64)
    args = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    assert world == 2
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    os.environ["NCCL_DEBUG"] = "WARN"
    dist.init_process_group("nccl", device_id=dev)

    if args.check:
        from oracle.decode_oracle import OracleQwen2
        from tests.helpers import tiny_cfg, tiny_weights
        cfg = tiny_cfg("gqa2")
        w = tiny_weights(cfg)
        loc = shard_fused_weights(cfg, w, rank, 2)
        arena = ParamArena(cfg.shard(2), dev)
        for name in arena.names():
            arena.view(name).copy_(loc[name].to(torch.bfloat16))
        ok = True
        detail = {}
        for use_graph in (False, True):
            eng = TPDecodeEngine(cfg, arena, rank, 2, max_batch=4, max_seq_len=192, max_new_tokens=24, device=dev,
                                 use_cuda_graph=use_graph)
            g = torch.Generator().manual_seed(5)
            prompts = [torch.randint(0, cfg.vocab_size, (n,), generator=g).tolist() for n in (9, 70, 33, 1, 64)]
            outs = eng.generate(prompts, SamplingParams(max_tokens=12, greedy=True))
            tag = "graph" if use_graph else "eager"
            worst, mism = 0.0, 0
            if rank == 0:
                orc = OracleQwen2(cfg, w)
                for pr, r in zip(prompts, outs):
                    logits = orc.forward(torch.tensor(pr))[-1]
                    for tok, lp in zip(r.output_ids, r.output_logprobs):
                        ref = torch.log_softmax(logits, -1)
                        top2 = torch.topk(logits, 2).values
                        if float(top2[0] - top2[1]) > 5e-2:
                            mism += int(int(torch.argmax(logits)) != tok)
                        worst = max(worst, abs(lp - float(ref[tok])))
                        logits = orc.forward(torch.tensor([tok]))[-1]
                    orc.reset()
                ok &= (mism == 0 and worst <= 3e-2)
                detail[tag] = {"max_logprob_err": round(worst, 4), "greedy_mismatches": mism}
            # both ranks sampled the same ids (same RNG stream, same merged partials)
            mine = torch.tensor([t for r in outs for t in r.output_ids], device=dev)
            both = [torch.zeros_like(mine) for _ in range(2)]
            dist.all_gather(both, mine)
            ok &= bool(torch.equal(both[0], both[1]))
            if rank == 0:
                detail[tag]["ranks_equal_greedy"] = bool(torch.equal(both[0], both[1]))
            # sampling (temperature 1) also agrees across ranks
            outs2 = eng.generate(prompts[:2], SamplingParams(max_tokens=8, temperature=1.0))
            mine = torch.tensor([t for r in outs2 for t in r.output_ids], device=dev)
            both = [torch.zeros_like(mine) for _ in range(2)]
            dist.all_gather(both, mine)
            ok &= bool(torch.equal(both[0], both[1]))
            if rank == 0:
                detail[tag]["ranks_equal_sampled"] = bool(torch.equal(both[0], both[1]))
            eng.close()
        flag = torch.tensor([1.0 if ok else 0.0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if rank == 0:
            print(json.dumps({"bench": "tp2_check", "ok": bool(flag.item() == 1.0), "detail": detail}), flush=True)
        dist.destroy_process_group()
        return

    cfg = ModelConfig.qwen2_5_32b() if args.model == "32b" else ModelConfig.qwen2_5_7b()
    loc_cfg = cfg.shard(2)
    arena = ParamArena(loc_cfg, dev).init_random(seed=42 + rank)
    room = 128 + args.steps * 2
    eng = TPDecodeEngine(cfg, arena, rank, 2, max_batch=args.batch, max_seq_len=args.context + room, max_new_tokens=room,
                         device=dev)
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    flat, step = eng.kv_cache, 1 << 28
    for s in range(0, flat.numel(), step):
        n = min(step, flat.numel() - s)
        flat[s:s + n] = (torch.randn(n, generator=g, device=dev, dtype=torch.float32) * 0.5).to(torch.bfloat16)
    B, mb = eng.B, eng.max_blocks
    eng.block_table.copy_(torch.arange(1, 1 + B * mb, dtype=torch.int32, device=dev).view(B, mb))
    eng.free_pages.clear()
    eng.prompt_len.fill_(args.context); eng.positions.fill_(args.context); eng.seq_lens.fill_(args.context + 1)
    eng.max_new_t.fill_(room); eng.active.fill_(1)
    eng.tokens.copy_(torch.randint(0, 151643, (B,), generator=torch.Generator().manual_seed(1000)).int())
    eng.temperature, eng.greedy, eng.ignore_eos = 1.0, False, True
    for _ in range(3):
        eng.step()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        eng.step()
    e1.record()
    torch.cuda.synchronize(); dist.barrier()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = t.item() / args.steps
    # both ranks must have sampled identical tokens
    mine = eng.sampled.clone()
    both = [torch.zeros_like(mine) for _ in range(2)]
    dist.all_gather(both, mine)
    same = bool(torch.equal(both[0], both[1]))
    if rank == 0:
        w_bytes = arena.nbytes() - loc_cfg.vocab_size * loc_cfg.hidden_size * 2   # embeddings are gathered, not streamed
        kv_bytes = loc_cfg.num_layers * B * (args.context + 1) * 2 * loc_cfg.num_kv_heads * 128 * 2
        peak = 6566.1
        try:
            peak = float(json.loads((Path(__file__).resolve().parent.parent / "MEASURED_PEAKS.json").read_text())["hbm_gbs"])
        except Exception:
            pass
        gbs = (w_bytes + kv_bytes) / (ms / 1e3) / 1e9
        print(json.dumps({"bench": "tp2_token_step", "model": "Qwen2.5-32B" if args.model == "32b" else "Qwen2.5-7B",
                          "tp": 2, "batch": B, "context": args.context, "ms_per_step": round(ms, 4),
                          "tokens_per_s": round(B / (ms / 1e3), 1), "per_rank_algorithmic_GB": round((w_bytes + kv_bytes) / 1e9, 2),
                          "per_rank_GBs": round(gbs, 1), "frac_of_hbm_peak": round(gbs / peak, 4),
                          "ranks_sampled_identical_tokens": same,
                          "nvlink_partial_bytes_per_step_per_dir": loc_cfg.num_layers * 2 * eng.split_k["o"] * B * loc_cfg.hidden_size * 4}),
              flush=True)
    eng.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
