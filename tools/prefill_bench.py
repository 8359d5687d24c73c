#!/usr/bin/env python
"""Prefill cost of the sampler (Qwen2.5-7B, 8192-token synthetic prompts): one prompt, and a GRPO group of 8
attempts of the same prompt with page-granular prefix sharing.  One JSON line."""
import json
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from pipelinerl_b200.engine import DecodeEngine, SamplingParams  # noqa: E402
from pipelinerl_b200.model import ModelConfig, ParamArena  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    cfg = ModelConfig.qwen2_5_7b()
    arena = ParamArena(cfg, dev).init_random(seed=42)
    P = 8192
    chunk = int(next((a.split("=")[1] for a in sys.argv if a.startswith("--chunk=")), 1024))
    eng = DecodeEngine(cfg, arena, max_batch=16, max_seq_len=P + 128, max_new_tokens=64, device=dev, prefill_chunk=chunk)
    g = torch.Generator().manual_seed(1)
    prompts = [torch.randint(8, 151643, (P,), generator=g).tolist() for _ in range(3)]
    sp = SamplingParams(max_tokens=4, greedy=True, ignore_eos=True)
    out = {}

    def timed(fn):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) * 1e3
    # warm-up (kernel attributes, buffers)
    eng.add_request(prompts[0][:1100], sp)
    eng.run_prefill()
    while eng.slot_req:
        eng.step(); eng.harvest()
    eng._evict_prefixes(10 ** 9)
    # one prompt
    eng.add_request(prompts[1], sp)
    ms1 = timed(eng.run_prefill)
    while eng.slot_req:
        eng.step(); eng.harvest()
    eng._evict_prefixes(10 ** 9)
    # a group of 8 attempts
    before = dict(eng.stats)
    for _ in range(8):
        eng.add_request(prompts[2], sp)
    ms8 = timed(eng.run_prefill)
    d = {k: eng.stats[k] - before[k] for k in eng.stats}
    flops = 2 * (cfg.num_params() - 2 * cfg.vocab_size * cfg.hidden_size) * (P - 1) + 2 * (P - 1) ** 2 * cfg.hidden_size * cfg.num_layers
    out = {"bench": "prefill", "model": "Qwen2.5-7B", "prompt_tokens": P, "chunk": chunk,
           "single_prompt_ms": round(ms1, 2), "single_prompt_tokens_per_s": round((P - 1) / ms1 * 1e3),
           "single_prompt_TFLOPs": round(flops / ms1 / 1e9, 1),
           "group_of_8_ms": round(ms8, 2), "group_prefill_tokens": d["prefill_tokens"],
           "group_prefix_hit_tokens": d["prefix_hit_tokens"],
           "group_effective_tokens_per_s": round(8 * (P - 1) / ms8 * 1e3)}
    print(json.dumps(out), flush=True)
    if "--profile" in sys.argv:
        from torch.profiler import ProfilerActivity, profile
        eng._evict_prefixes(10 ** 9)
        while eng.slot_req:
            eng.step(); eng.harvest()
        eng._evict_prefixes(10 ** 9)
        eng.add_request(prompts[0], sp)
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            eng.run_prefill()
            torch.cuda.synchronize()
        print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=12, max_name_column_width=60), file=sys.stderr)


if __name__ == "__main__":
    main()
