"""Preprocess stage: rollout groups -> RL columns -> packed micro-batches.

Host code mirroring pipelinerl/preprocess.py: `preprocess_dataset` (:145-189: prepare_rl_fields per sample,
populate_rl_data per chunk of whole groups, ref logprobs copied from old logprobs when kl_coef == 0, :160-161)
and the writer's packing rule (:598-637: greedily add samples to the current micro-batch while the token
count stays <= seq_length; `collate_packed`; round-robin over trainer ranks; sentinel batches equalise the
micro-batch counts per optimizer step).
"""
from __future__ import annotations

from typing import Any

from .finetune.data import collate_packed, preprocess_fn
from .finetune.rl import RLConfig, populate_rl_data
from .finetune.types import PipelineBatchEncoding


def preprocess_dataset(samples: list[dict[str, Any]], tokenizer, seq_length: int, rl_config: RLConfig) -> list[dict]:
    """samples: TrainingText.model_dump() dicts of WHOLE groups."""
    entries = []
    for s in samples:
        s = dict(s)
        if not s.get("ref_logprobs"):
            s["ref_logprobs"] = list(s["logprobs"])
        enc = preprocess_fn(s, tokenizer, seq_length, is_rl=True)
        meta = s.get("metadata", {})
        enc["group_id"] = s["group_id"]
        enc["rollout_index"] = meta.get("rollout_index", 0)
        enc["step_index"] = meta.get("step_index", 0)
        enc["finished"] = s.get("finished", False)
        enc["model_version"] = meta.get("model_version", 0)
        entries.append(enc)
    entries = populate_rl_data(entries, tokenizer.eos_token_id, rl_config)
    entries = [e for e in entries if len(e["input_ids"]) <= seq_length]
    if rl_config.filter_zero_advantage_groups:
        entries, _ = filter_zero_advantage_groups(entries)
    return entries


def filter_zero_advantage_groups(entries: list[dict], epsilon: float = 1e-6) -> tuple[list[dict], int]:
    """Drop every group none of whose samples carries an advantage with |a| > epsilon (all attempts got the same
    reward: no learning signal).  Applied by the preprocessor when `RLConfig.filter_zero_advantage_groups` is set
    (pipelinerl/preprocess.py:316-353, call site :548-552).  Groups keep their first-seen order, samples their order
    inside the group; returns (kept entries, number of samples dropped)."""
    by_group: dict[Any, list[dict]] = {}
    for e in entries:
        by_group.setdefault(e["group_id"], []).append(e)
    kept: list[dict] = []
    dropped = 0
    for members in by_group.values():
        if any(abs(a) > epsilon for m in members for a in m["advantages"]):
            kept.extend(members)
        else:
            dropped += len(members)
    return kept, dropped


def pack_micro_batches(entries: list[dict], tokenizer, seq_length: int, seq_parallel: int = 1,
                       pin_memory: bool = False) -> list[PipelineBatchEncoding]:
    batches, cur, cur_len = [], [], 0
    for e in entries:
        n = len(e["input_ids"])
        if cur and cur_len + n > seq_length:
            batches.append(collate_packed(cur, tokenizer, seq_parallel, pin_memory=pin_memory))
            cur, cur_len = [], 0
        cur.append(e)
        cur_len += n
    if cur:
        batches.append(collate_packed(cur, tokenizer, seq_parallel, pin_memory=pin_memory))
    return batches
