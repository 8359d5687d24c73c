"""Batch container handed from the preprocessor to the trainer.

Mirrors the interface of the reference's pydantic `PipelineBatchEncoding`
(pipelinerl/finetune/types.py:46-180): same field names, dtypes (int64 token
columns, fp32 RL columns, int32 seq_boundaries), `to_device`, `from_dict`,
`make_slices`, `model_dump`.  It is a plain Python class: the trainer re-creates
one per micro-batch and pydantic validation of 16 K-token tensors is measurable
host time the hot path does not need.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any

import numpy as np
import torch

LONG_FIELDS = ("input_ids", "attention_mask", "labels", "position_ids", "segment_ids", "image_grid_thw")
FLOAT_FIELDS = ("rewards", "advantages", "ref_logprobs", "old_logprobs", "group_tokens", "num_labels", "overflow",
                "pixel_values")
ROW_FIELDS = ("input_ids", "attention_mask", "labels", "position_ids", "segment_ids", "rewards", "advantages",
              "ref_logprobs", "old_logprobs", "group_tokens", "overflow", "num_labels")
META_DEFAULTS = {"model_version": None, "sentinel": False, "padding": 0, "is_packed": False}
REQUIRED = ("input_ids", "attention_mask", "labels", "rewards", "advantages", "ref_logprobs", "old_logprobs",
            "group_tokens", "num_labels", "overflow", "model_version")


def _as_tensor(value, dtype):
    if value is None:
        return None
    if isinstance(value, torch.Tensor):
        return value.to(dtype)
    if isinstance(value, (list, tuple, np.ndarray)):
        return torch.as_tensor(np.asarray(value), dtype=dtype) if isinstance(value, np.ndarray) \
            else torch.tensor(value, dtype=dtype)
    raise ValueError(f"Unsupported type for tensor field: {type(value)}")


@dataclass
class TrainingMetrics:
    """Counters the trainer persists (reference: finetune/types.py:27-43)."""
    epoch: int = 0
    passes: int = 0
    completed_steps: int = 0
    samples: int = 0
    tokens: int = 0
    samples_too_old_to_queue: int = 0
    samples_too_old_to_train: int = 0
    last_broadcasted_version: int = 0
    train_loss: float = 1e9
    eval_loss: float = 1e9
    dev_loss: float = 1e9
    grad_norm: float = 0.0
    best_eval_loss: float = 1e9
    best_completed_steps: int = 0
    lr: float = 0.0
    time_waiting_for_data: float = 0.0


class PipelineBatchEncoding:
    model_fields = (*LONG_FIELDS[:5], *FLOAT_FIELDS[:7], "model_version", "sentinel", "padding", "is_packed",
                    "seq_boundaries", "pixel_values", "image_grid_thw")

    def __init__(self, **kw: Any):
        missing = [k for k in REQUIRED if k not in kw]
        if missing:
            raise ValueError(f"PipelineBatchEncoding: missing fields {missing}")
        unknown = [k for k in kw if k not in self.model_fields]
        if unknown:
            raise ValueError(f"PipelineBatchEncoding: unknown fields {unknown}")
        for name in LONG_FIELDS:
            setattr(self, name, _as_tensor(kw.get(name), torch.long))
        for name in FLOAT_FIELDS:
            setattr(self, name, _as_tensor(kw.get(name), torch.float32))
        self.seq_boundaries = _as_tensor(kw.get("seq_boundaries"), torch.int32)
        for name, default in META_DEFAULTS.items():
            setattr(self, name, kw.get(name, default))
        self.model_version = int(self.model_version)
        self.sentinel = bool(self.sentinel)
        self.is_packed = bool(self.is_packed)
        self.padding = int(self.padding)
        self.model_extra: dict[str, Any] = {}

    # -- reference-compatible helpers ---------------------------------------------
    def to_device(self, device) -> "PipelineBatchEncoding":
        for name in self.model_fields:
            v = getattr(self, name)
            if isinstance(v, torch.Tensor):
                setattr(self, name, v.to(device, non_blocking=True))
        return self

    def pin_memory(self) -> "PipelineBatchEncoding":
        for name in self.model_fields:
            v = getattr(self, name)
            if isinstance(v, torch.Tensor) and v.device.type == "cpu":
                setattr(self, name, v.pin_memory())
        return self

    @classmethod
    def from_dict(cls, data: dict[str, Any], **defaults) -> "PipelineBatchEncoding":
        merged = {**defaults, **data}
        known = {k: v for k, v in merged.items() if k in cls.model_fields}
        inst = cls(**known)
        inst.model_extra.update({k: v for k, v in merged.items() if k not in cls.model_fields})
        return inst

    def model_dump(self) -> dict[str, Any]:
        return {name: getattr(self, name) for name in self.model_fields}

    def make_slices(self, num_slices: int) -> list["PipelineBatchEncoding"]:
        """Equal contiguous slices of a packed row for sequence parallelism (types.py:145-180)."""
        if self.position_ids is None or self.input_ids.shape[0] > 1:
            raise ValueError("Cannot a batch that is not properly packed")
        length = self.input_ids.shape[1]
        if length < num_slices:
            raise ValueError(f"Cannot slice batch of size {length} into {num_slices} slices")
        if length % num_slices != 0:
            raise ValueError(f"Sequence length {length} is not divisible by number of slices {num_slices}")
        step = length // num_slices
        out = []
        for i in range(num_slices):
            lo, hi = i * step, (i + 1) * step
            kw = {}
            for name in ROW_FIELDS:
                v = getattr(self, name)
                kw[name] = v[:, lo:hi] if v is not None else None
            kw.update(model_version=self.model_version, sentinel=self.sentinel, is_packed=self.is_packed,
                      padding=self.padding, seq_boundaries=self.seq_boundaries, pixel_values=self.pixel_values,
                      image_grid_thw=self.image_grid_thw)
            out.append(PipelineBatchEncoding(**kw))
        return out

    def __repr__(self) -> str:
        shape = tuple(self.input_ids.shape) if self.input_ids is not None else None
        return (f"PipelineBatchEncoding(input_ids={shape}, is_packed={self.is_packed}, sentinel={self.sentinel}, "
                f"model_version={self.model_version}, padding={self.padding})")
