"""Synthetic single-turn domain for the BASELINE configs 2-4 (random-init model, synthetic prompts).

Same plugin surface as a real domain (cf. pipelinerl/domains/math/rollouts.py:58-149): one LLM call per
rollout on a fixed token-id prompt, reward ~ Bernoulli(0.5) seeded per (problem, attempt) as SURVEY §8d
prescribes (random-init weights never solve anything, so the reward must be synthetic).
"""
from __future__ import annotations

import time
import zlib

import numpy as np

from ..async_llm import llm_async_generate, make_training_text
from ..llm import Prompt, TrainableLLM
from ..rollouts import BaseMetrics, RolloutResult

DOMAIN = "synthetic"
_counter: dict[int, int] = {}


async def generate_synthetic_rollout(cfg, llm: TrainableLLM, problem: dict, session=None) -> RolloutResult:
    t0 = time.time()
    attempt = _counter.get(problem["id"], 0)
    _counter[problem["id"]] = attempt + 1
    prompt = Prompt(messages=[{"role": "user", "content": f"problem {problem['id']}"}], token_ids=problem["prompt_ids"])
    call = await llm_async_generate(llm, prompt, session)
    text = make_training_text(llm, call)
    seed = zlib.crc32(f"{problem['id']}/{attempt}".encode())
    reward = float(np.random.default_rng(2000 + seed).random() < 0.5)
    text.reward = reward
    return RolloutResult(training_texts=[text], latency=time.time() - t0, dataset_name=problem["dataset"], domain=DOMAIN,
                         metrics=BaseMetrics(reward=reward, success=reward > 0, no_error=True, no_answer=False))


def load_problems(dataset_names: list[str], n_problems: int = 64, prompt_tokens: int = 8192,
                  vocab_limit: int = 151_643) -> list[dict]:
    out = []
    for name in dataset_names:
        for p in range(n_problems):
            rng = np.random.default_rng(1000 + p + (0 if name == "train" else 10_000_000))
            out.append({"id": p, "dataset": name, "domain": DOMAIN,
                        "prompt_ids": rng.integers(8, vocab_limit, size=prompt_tokens).tolist()})
    return out
