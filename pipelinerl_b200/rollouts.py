"""Rollout record types of the plugin surface (same names/fields as pipelinerl/rollouts.py:6-97)."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any, Sequence

from pydantic import BaseModel, Field


class BaseMetrics(BaseModel):
    reward: float
    success: bool
    no_error: bool
    no_answer: bool


class TrainingText(BaseModel):
    """One trainable sample: prompt + generated tokens with the sampler's logprobs."""
    model_config = {"arbitrary_types_allowed": True}

    text: str
    n_predicted: int
    reward: float = 0.0
    logprobs: list[float] = Field(default_factory=list)
    ref_logprobs: list[float] = Field(default_factory=list)
    input_ids: list[int] = Field(default_factory=list)
    labels: list[int] = Field(default_factory=list)
    group_id: str | None = None
    finished: bool = False
    prompt_tokens: int = 0
    output_tokens: int = 0
    visual_features: dict[str, Any] | None = None
    metadata: dict = Field(default_factory=dict)

    @property
    def prompt_text(self) -> str:
        return self.text[: -self.n_predicted]

    @property
    def output_text(self) -> str:
        return self.text[-self.n_predicted:]


class RolloutResult(BaseModel):
    training_texts: list[TrainingText]
    metrics: BaseMetrics
    latency: float
    model_version: int | None = None
    dataset_name: str | None = None
    group_id: str | None = None
    domain: str | None = None


@dataclass(frozen=True)
class TrainingTextSummary:
    prompt_tokens: list[int]
    output_tokens: list[int]
    overflow: bool
    num_turns: int


def apply_rollout_reward(training_texts: Sequence[TrainingText], reward: float) -> list[TrainingText]:
    out = list(training_texts)
    for t in out:
        t.reward = reward
    return out


def rollout_has_overflow(training_texts: Sequence[TrainingText]) -> bool:
    return any(not t.finished for t in training_texts)


def summarize_training_texts(training_texts: Sequence[TrainingText]) -> TrainingTextSummary:
    ts = list(training_texts)
    return TrainingTextSummary([t.prompt_tokens for t in ts], [t.output_tokens for t in ts], rollout_has_overflow(ts),
                               len(ts))
