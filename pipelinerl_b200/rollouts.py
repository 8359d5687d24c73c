"""Records a rollout plugin hands back to the actor.

The field NAMES are the plugin contract of the reference (pipelinerl/rollouts.py:6-68) and are what
`make_training_text`, the `actor` topic and the preprocessor read; everything else here is this
repository's own arrangement.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any, Iterable

from pydantic import BaseModel, ConfigDict, Field


def rollout_has_overflow(training_texts: Iterable["TrainingText"]) -> bool:
    """True when any turn of the rollout was cut by the token limit (finished == False)."""
    for text in training_texts:
        if not text.finished:
            return True
    return False


def apply_rollout_reward(training_texts: Iterable["TrainingText"], reward: float) -> list["TrainingText"]:
    """Stamp one scalar reward on every turn of a rollout."""
    stamped = []
    for text in training_texts:
        text.reward = reward
        stamped.append(text)
    return stamped


class _Record(BaseModel):
    model_config = ConfigDict(arbitrary_types_allowed=True)


class BaseMetrics(_Record):
    """Outcome of one rollout; domains extend it with their own counters."""
    reward: float
    success: bool
    no_error: bool
    no_answer: bool


class TrainingText(_Record):
    """One LLM call turned into a trainable sample.

    token-level payload (what the trainer consumes, exact ids and sampler logprobs):
        input_ids = prompt ids + generated ids, labels = -100 over the prompt then the generated ids,
        logprobs / ref_logprobs aligned with the generated ids
    bookkeeping: text (+ n_predicted characters at its end are the completion), reward, group_id,
        finished (False = cut by max_tokens), prompt_tokens / output_tokens, metadata (model_version,
        rollout_index, step_index are added by the actor), visual_features for VLMs.
    """
    # token-level payload
    input_ids: list[int] = Field(default_factory=list)
    labels: list[int] = Field(default_factory=list)
    logprobs: list[float] = Field(default_factory=list)
    ref_logprobs: list[float] = Field(default_factory=list)
    # text view
    text: str
    n_predicted: int
    # bookkeeping
    reward: float = 0.0
    group_id: str | None = None
    finished: bool = False
    prompt_tokens: int = 0
    output_tokens: int = 0
    metadata: dict = Field(default_factory=dict)
    visual_features: dict[str, Any] | None = None

    @property
    def output_text(self) -> str:
        return self.text[-self.n_predicted:]   # same slicing (and n_predicted == 0 quirk) as the reference

    @property
    def prompt_text(self) -> str:
        return self.text[: -self.n_predicted]


class RolloutResult(_Record):
    """What `generate_rollout(cfg, llm, problem, session)` returns; the actor fills the optional fields."""
    training_texts: list[TrainingText]
    metrics: BaseMetrics
    latency: float
    model_version: int | None = None
    group_id: str | None = None
    dataset_name: str | None = None
    domain: str | None = None


@dataclass(frozen=True)
class TrainingTextSummary:
    prompt_tokens: list[int]
    output_tokens: list[int]
    overflow: bool
    num_turns: int


def summarize_training_texts(training_texts: Iterable[TrainingText]) -> TrainingTextSummary:
    turns = list(training_texts)
    return TrainingTextSummary(prompt_tokens=[t.prompt_tokens for t in turns],
                               output_tokens=[t.output_tokens for t in turns],
                               overflow=rollout_has_overflow(turns), num_turns=len(turns))
