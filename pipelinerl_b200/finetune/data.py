"""Sample -> micro-batch layout (the integer, bit-exact contract between preprocessor and trainer).

Same function names and results as the reference (pipelinerl/finetune/data.py):
  preprocess_fn   :111-160   pre-tokenised samples only (the actor always supplies input_ids/labels)
  collate_packed  :215-283   one [1, T] row, position_ids restart per sample, segment_ids = sample index,
                             first label of every non-first sample masked, int32 seq_boundaries,
                             pad-to-seq_parallel with a sentinel sample
  collate         :163-212   padded [B, L] batch (L rounded up to a multiple of 16)
Built with numpy concatenation instead of per-sample torch slicing; output tensors can be pinned for
the host->device copy.
"""
from __future__ import annotations

from typing import Any

import numpy as np
import torch

from .rl import RL_DATA_COLUMNS, prepare_rl_fields
from .types import PipelineBatchEncoding
from .utils import create_sentinel_example

MASKED_TOKEN_ID = -100


def preprocess_fn(entry: dict[str, Any], tokenizer, seq_length: int, is_rl: bool = False) -> dict[str, Any]:
    if not entry.get("input_ids"):
        raise NotImplementedError(
            "text-only samples need a tokenizer pass; the actor of this framework always publishes input_ids/labels")
    encoding = {
        "input_ids": entry["input_ids"],
        "labels": entry["labels"],
        "attention_mask": [1] * len(entry["input_ids"]),
    }
    if is_rl:
        encoding = prepare_rl_fields(encoding, entry["reward"], entry["logprobs"], entry["ref_logprobs"])
    for key in ("pixel_values", "image_thw"):
        if key in entry:
            encoding[key] = entry[key]
    return encoding


def collate_packed(examples: list[dict[str, Any]], tokenizer, seq_parallel: int,
                   label_pad_value: int = MASKED_TOKEN_ID, pin_memory: bool = False) -> PipelineBatchEncoding:
    lengths = [len(ex["input_ids"]) for ex in examples]
    total = sum(lengths)
    padding = 0
    if total % seq_parallel != 0:
        padding = seq_parallel - total % seq_parallel
        version = max(ex["model_version"] for ex in examples)
        examples = examples + [create_sentinel_example(padding, tokenizer=tokenizer, model_version=version)]
        lengths.append(padding)
        total += padding

    bounds = np.zeros(len(examples) + 1, dtype=np.int64)
    np.cumsum(lengths, out=bounds[1:])
    input_ids = np.concatenate([np.asarray(ex["input_ids"], dtype=np.int64) for ex in examples]) if total else \
        np.zeros(0, np.int64)
    labels = np.concatenate([np.asarray(ex["labels"], dtype=np.int64) for ex in examples]) if total else \
        np.zeros(0, np.int64)
    starts = bounds[:-1]
    nonempty = np.asarray(lengths) > 0
    # the first token of every sample but the first has no in-sample predecessor: never a target
    mask_at = starts[1:][nonempty[1:]]
    labels[mask_at] = label_pad_value
    segment_ids = np.repeat(np.arange(len(examples), dtype=np.int64), lengths)
    position_ids = np.arange(total, dtype=np.int64) - np.repeat(starts, lengths)

    fields: dict[str, Any] = {
        "input_ids": input_ids[None], "labels": labels[None], "attention_mask": np.ones((1, total), np.int64),
        "position_ids": position_ids[None], "segment_ids": segment_ids[None],
    }
    for key in RL_DATA_COLUMNS:
        if key not in examples[0]:
            continue
        parts = []
        for ex in examples:
            v = ex[key]
            parts.append(np.asarray(v, dtype=np.float32) if isinstance(v, (list, tuple, np.ndarray))
                         else np.asarray([v], dtype=np.float32))
        fields[key] = np.concatenate(parts)[None]
    tensors = {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in fields.items()}
    if pin_memory:
        tensors = {k: v.pin_memory() for k, v in tensors.items()}
    return PipelineBatchEncoding(
        **tensors,
        model_version=min(ex.get("model_version", 0) for ex in examples),
        is_packed=True,
        seq_boundaries=torch.from_numpy(bounds.astype(np.int32)),
        padding=padding,
    )


def collate(examples: list[dict[str, Any]], tokenizer, label_mask_value: int = MASKED_TOKEN_ID,
            pad_to_multiple_of: int = 16) -> PipelineBatchEncoding:
    width = max(len(ex["input_ids"]) for ex in examples)
    if width % pad_to_multiple_of:
        width += pad_to_multiple_of - width % pad_to_multiple_of
    right = getattr(tokenizer, "padding_side", "right") == "right"
    out: dict[str, Any] = {}
    for key in examples[0]:
        if key == "model_version":
            continue
        column = [ex[key] for ex in examples]
        if any(isinstance(v, (str, dict)) for v in column):
            continue
        if any(isinstance(x, (str, dict)) for v in column if isinstance(v, list) for x in v):
            continue
        is_float = key in RL_DATA_COLUMNS
        fill = label_mask_value if key == "labels" else (0.0 if is_float else 0)
        rows = []
        for v in column:
            if v is None:
                continue
            v = v if isinstance(v, list) else [v]
            pad = [fill] * (width - len(v))
            rows.append(v + pad if right else pad + v)
        out[key] = torch.tensor(rows)
    out["model_version"] = min(ex.get("model_version", 0) for ex in examples)
    out["is_packed"] = False
    return PipelineBatchEncoding(**out)
