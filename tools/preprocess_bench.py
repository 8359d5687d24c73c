#!/usr/bin/env python
"""CPU feeder of hot path 2 (rows a4/a5): reference pandas/python-list pipeline vs this repo's host code on
identical inputs (groups of 8 samples x 16384 tokens).  Run in the authoring container (needs
/root/reference for the reference arm; elsewhere only ours + the oracle port are timed).

    python tools/preprocess_bench.py [--groups 4]
"""
import argparse
import copy
import json
import os
import sys
import time
import types
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


class Tok:
    eos_token_id = 7
    padding_side = "right"


def make_samples(n_groups, attempts=8, prompt=8192, gen=8192, vocab=151643):
    rng = np.random.default_rng(0)
    out = []
    for g in range(n_groups):
        p = rng.integers(8, vocab, size=prompt).tolist()
        for a in range(attempts):
            gen_ids = rng.integers(8, vocab, size=gen).tolist()
            out.append({"input_ids": p + gen_ids, "labels": [-100] * prompt + gen_ids,
                        "logprobs": (-rng.random(gen) * 12).tolist(), "ref_logprobs": (-rng.random(gen) * 12).tolist(),
                        "reward": float(rng.random() < 0.5), "group_id": f"g{g}", "rollout_index": a, "step_index": 0,
                        "finished": False, "model_version": 1})
    return out


def run_ours(samples, seq_length):
    from pipelinerl_b200.finetune.data import collate_packed, preprocess_fn
    from pipelinerl_b200.finetune.rl import RLConfig, populate_rl_data
    cfg = RLConfig()
    t0 = time.perf_counter()
    entries = []
    for s in samples:
        e = preprocess_fn(s, Tok(), seq_length, is_rl=True)
        for k in ("group_id", "rollout_index", "step_index", "finished", "model_version"):
            e[k] = s[k]
        entries.append(e)
    entries = populate_rl_data(entries, Tok.eos_token_id, cfg)
    batches = [collate_packed([e], Tok(), 1) for e in entries]
    return time.perf_counter() - t0, batches


def run_reference(samples, seq_length):
    import transformers  # noqa: F401
    sys.path.insert(0, "/root/reference")
    for name, attrs in (("omegaconf", {"DictConfig": dict, "ListConfig": list, "OmegaConf": object}),
                        ("accelerate", {"Accelerator": object})):
        m = types.ModuleType(name)
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules.setdefault(name, m)
    from pipelinerl.finetune import data as rd
    from pipelinerl.finetune import rl as rr
    cfg = rr.RLConfig()
    t0 = time.perf_counter()
    entries = []
    for s in samples:
        e = rd.preprocess_fn(dict(s), Tok(), seq_length, is_rl=True)
        for k in ("group_id", "rollout_index", "step_index", "finished", "model_version"):
            e[k] = s[k]
        entries.append(e)
    entries = rr.populate_rl_data(entries, Tok.eos_token_id, cfg)
    batches = [rd.collate_packed([e], Tok(), 1) for e in entries]
    return time.perf_counter() - t0, batches


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--groups", type=int, default=4)
    args = ap.parse_args()
    samples = make_samples(args.groups)
    n_tok = sum(len(s["input_ids"]) for s in samples)
    t_ours, b_ours = run_ours(copy.deepcopy(samples), 20000)
    out = {"bench": "preprocess_feeder", "samples": len(samples), "tokens": n_tok, "cores_used": 1,
           "host_cores": os.cpu_count(),
           "ours": {"seconds": round(t_ours, 3), "samples_per_s": round(len(samples) / t_ours, 2),
                    "tokens_per_s": round(n_tok / t_ours)}}
    if Path("/root/reference").exists():
        t_ref, b_ref = run_reference(copy.deepcopy(samples), 20000)
        import torch
        same = all(torch.equal(getattr(a, k), getattr(b, k)) for a, b in zip(b_ours, b_ref)
                   for k in ("input_ids", "labels", "position_ids", "segment_ids", "rewards", "advantages",
                             "old_logprobs", "ref_logprobs", "group_tokens", "num_labels", "overflow", "seq_boundaries"))
        out["reference"] = {"seconds": round(t_ref, 3), "samples_per_s": round(len(samples) / t_ref, 2),
                            "tokens_per_s": round(n_tok / t_ref)}
        out["outputs_bit_identical"] = bool(same)
        out["speedup"] = round(t_ref / t_ours, 2)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
