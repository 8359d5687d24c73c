"""Sentinel micro-batches / samples that keep data-parallel learners in lock-step.

Same results as pipelinerl/finetune/utils.py:17-78: a sentinel is 8 (or n) EOS tokens with every label
masked, unit group_tokens/num_labels and zero everything else, so that rl_step yields exactly zero
loss while every rank still runs forward/backward.
"""
from __future__ import annotations

from typing import Callable

import torch

from .types import PipelineBatchEncoding

SENTINEL_LENGTH = 8
_DEFAULT_EOS = 2


def create_sentinel_batch(device, tokenizer=None, model_version=0) -> PipelineBatchEncoding:
    eos = getattr(tokenizer, "eos_token_id", _DEFAULT_EOS) if tokenizer else _DEFAULT_EOS
    n = SENTINEL_LENGTH
    zeros = torch.zeros(1, n, dtype=torch.float32)
    ones = torch.ones(1, n, dtype=torch.float32)
    batch = PipelineBatchEncoding(
        input_ids=torch.full((1, n), eos, dtype=torch.long),
        attention_mask=torch.ones(1, n, dtype=torch.long),
        labels=torch.full((1, n), -100, dtype=torch.long),
        position_ids=torch.arange(n, dtype=torch.long).reshape(1, n),
        segment_ids=torch.zeros(1, n, dtype=torch.long),
        rewards=zeros.clone(), advantages=zeros.clone(), ref_logprobs=zeros.clone(), old_logprobs=zeros.clone(),
        group_tokens=ones.clone(), num_labels=ones.clone(), overflow=zeros.clone(),
        seq_boundaries=torch.tensor([0, n], dtype=torch.int32),
        model_version=model_version, sentinel=True, is_packed=True,
    )
    return batch


def create_sentinel_example(n_tokens: int, tokenizer=None, model_version=0) -> dict:
    eos = tokenizer.eos_token_id
    return {
        "input_ids": [eos] * n_tokens,
        "attention_mask": [1] * n_tokens,
        "labels": [-100] * n_tokens,
        "position_ids": list(range(n_tokens)),
        "rewards": [0.0] * n_tokens,
        "advantages": [0.0] * n_tokens,
        "ref_logprobs": [0.0] * n_tokens,
        "old_logprobs": [0.0] * n_tokens,
        "group_tokens": [1.0] * n_tokens,
        "num_labels": [1.0] * n_tokens,
        "overflow": [0.0] * n_tokens,
        "model_version": model_version,
    }


def dummy_eval_callback(config_name: str) -> Callable:
    return lambda *a, **k: {}
