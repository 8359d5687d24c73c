"""Oracle for the CPU feeder of hot path (2): RL field preparation, leave-one-out
advantages and sequence packing.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Plain Python loops + numpy; no
pandas.  This is also the "port" CPU baseline bench.py times.

Follows:
  prepare_rl_fields   pipelinerl/finetune/rl/__init__.py:573-594
  populate_rl_data    pipelinerl/finetune/rl/__init__.py:453-570
  collate_packed      pipelinerl/finetune/data.py:215-283
  sentinel example    pipelinerl/finetune/utils.py:60-78
  sentinel batch      pipelinerl/finetune/utils.py:17-57
Pinned by tests/test_oracle_golden.py against tests/golden/preprocess_cases.json
(outputs of the reference functions themselves).
"""
from __future__ import annotations

import math

import numpy as np

IGNORE = -100
RL_COLUMNS = ["overflow", "group_tokens", "num_labels", "rewards", "advantages", "old_logprobs", "ref_logprobs"]


def prepare_rl_fields(input_ids, labels, reward, old_logprobs, ref_logprobs):
    n = len(labels)
    n_target = sum(1 for x in labels if x != IGNORE)
    if n_target != len(old_logprobs):
        raise AssertionError(f"Target tokens: {n_target}, old logprobs: {len(old_logprobs)}")
    return {
        "input_ids": list(input_ids), "labels": list(labels), "attention_mask": [1] * len(input_ids),
        "rewards": [reward] * n,
        "advantages": [0.0] * n,
        "old_logprobs": [0] * (n - len(old_logprobs)) + list(old_logprobs),
        "ref_logprobs": [0] * (n - len(ref_logprobs)) + list(ref_logprobs),
        "overflow": [0] * n,
        "group_tokens": [0] * n,
        "num_labels": [1 if x != IGNORE else 0 for x in labels],
    }


def _sample_std(xs):
    """pandas .std(): ddof=1, NaN for a single value."""
    k = len(xs)
    if k < 2:
        return math.nan
    mu = sum(xs) / k
    return math.sqrt(sum((x - mu) ** 2 for x in xs) / (k - 1))


def populate_rl_data(entries, eos_token_id, divide_advantage_by_std=True):
    """entries: dicts with input_ids, labels, rewards (per token), group_id, rollout_index, step_index,
    optional finish_reason / finished.  Adds advantages, group_tokens, overflow, num_labels in place."""
    rollout_tokens = {}
    for e in entries:
        key = (e["group_id"], e["rollout_index"])
        rollout_tokens[key] = rollout_tokens.get(key, 0) + len(e["input_ids"])
    per_group = {}
    for (g, _), v in rollout_tokens.items():
        per_group.setdefault(g, []).append(v)
    group_tokens = {g: sum(v) / len(v) for g, v in per_group.items()}

    step_rewards = {}
    for e in entries:
        step_rewards.setdefault((e["group_id"], e["step_index"]), []).append(e["rewards"][0])

    for e in entries:
        rs = step_rewards[(e["group_id"], e["step_index"])]
        r0 = e["rewards"][0]
        cnt = len(rs)
        loo = (sum(rs) - r0) / (cnt - 1) if cnt > 1 else r0
        if divide_advantage_by_std:
            sd = _sample_std(rs)
            sd = 0.0 if math.isnan(sd) else sd
            e["advantages"] = [(r - loo) / (sd + 1e-4) for r in e["rewards"]]
        else:
            e["advantages"] = [(r - loo) for r in e["rewards"]]
        n = len(e["input_ids"])
        fr = e.get("finish_reason")
        over = None
        if isinstance(fr, str):
            fr = fr.strip().lower()
            if fr == "length":
                over = 1.0
            elif fr in ("stop", "content_filter"):
                over = 0.0
        if over is None:
            if e.get("finished"):
                over = 0.0
            else:
                over = 0.0 if eos_token_id in e["input_ids"] else 1.0
        e["overflow"] = [over] * len(e["overflow"])
        e["group_tokens"] = [group_tokens[e["group_id"]]] * n
        e["num_labels"] = [sum(1 for x in e["labels"] if x != IGNORE)] * n
    return entries


def sentinel_example(n_tokens, eos_token_id, model_version=0):
    return {
        "input_ids": [eos_token_id] * n_tokens, "attention_mask": [1] * n_tokens, "labels": [IGNORE] * n_tokens,
        "position_ids": list(range(n_tokens)), "rewards": [0.0] * n_tokens, "advantages": [0.0] * n_tokens,
        "ref_logprobs": [0.0] * n_tokens, "old_logprobs": [0.0] * n_tokens, "group_tokens": [1.0] * n_tokens,
        "num_labels": [1.0] * n_tokens, "overflow": [0.0] * n_tokens, "model_version": model_version,
    }


def collate_packed(examples, eos_token_id, seq_parallel=1):
    """-> dict of numpy arrays with the reference's dtypes ([1,T] int64 / float32, seq_boundaries int32)."""
    examples = list(examples)
    total = sum(len(e["input_ids"]) for e in examples)
    padding = 0
    if total % seq_parallel != 0:
        padding = seq_parallel - total % seq_parallel
        examples.append(sentinel_example(padding, eos_token_id, max(e["model_version"] for e in examples)))
        total += padding
    ids = np.empty((1, total), np.int64)
    labels = np.empty((1, total), np.int64)
    pos = np.empty((1, total), np.int64)
    seg = np.empty((1, total), np.int64)
    bounds = np.zeros(len(examples) + 1, np.int32)
    cols = {k: [] for k in RL_COLUMNS if k in examples[0]}
    at = 0
    for i, e in enumerate(examples):
        n = len(e["input_ids"])
        ids[0, at:at + n] = e["input_ids"]
        pos[0, at:at + n] = np.arange(n)
        seg[0, at:at + n] = i
        lab = np.asarray(e["labels"], np.int64).copy()
        if i > 0 and n > 0:
            lab[0] = IGNORE
        labels[0, at:at + n] = lab
        for k in cols:
            cols[k].extend(e[k])
        at += n
        bounds[i + 1] = at
    out = {"input_ids": ids, "labels": labels, "attention_mask": np.ones((1, total), np.int64),
           "position_ids": pos, "segment_ids": seg, "seq_boundaries": bounds}
    for k, v in cols.items():
        out[k] = np.asarray(v, np.float32).reshape(1, -1)
    out["model_version"] = min(e.get("model_version", 0) for e in examples)
    out["is_packed"] = True
    out["padding"] = padding
    out["sentinel"] = False
    return out


def sentinel_batch(eos_token_id=2, model_version=0):
    n = 8
    z = np.zeros((1, n), np.float32)
    o = np.ones((1, n), np.float32)
    return {
        "input_ids": np.full((1, n), eos_token_id, np.int64), "attention_mask": np.ones((1, n), np.int64),
        "labels": np.full((1, n), IGNORE, np.int64), "position_ids": np.arange(n, dtype=np.int64)[None],
        "segment_ids": np.zeros((1, n), np.int64), "rewards": z.copy(), "advantages": z.copy(),
        "ref_logprobs": z.copy(), "old_logprobs": z.copy(), "group_tokens": o.copy(), "num_labels": o.copy(),
        "overflow": z.copy(), "seq_boundaries": np.array([0, n], np.int32), "model_version": model_version,
        "sentinel": True, "is_packed": True, "padding": 0,
    }
