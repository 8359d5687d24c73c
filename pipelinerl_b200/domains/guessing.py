"""Canonical plugin example: the 13-turn number-guessing game (reference: pipelinerl/domains/guessing/guessing.py).

Plugin surface kept verbatim:
    async generate_guessing_rollout(cfg, llm: TrainableLLM, problem: dict, session) -> RolloutResult
    load_problems(dataset_names: list[str]) -> list[dict]
Rewards: 2 - turn/10 on a correct guess, -2 + turn/10 when the output has no <answer>N</answer>; every turn's
LLM call becomes one TrainingText carrying the rollout reward.
"""
from __future__ import annotations

import re
import time

from ..async_llm import llm_async_generate, make_training_text
from ..llm import Prompt, TrainableLLM
from ..rollouts import BaseMetrics, RolloutResult

DOMAIN = "guessing"
MAX_TURNS = 13
_ANSWER = re.compile(r"<answer>(\d+)</answer>")
_TASK = ("You must guess a number between 1 and 1024. Output the answer as <answer>number</answer>."
         " After each guess I will tell you if your answer is higher or lower than the target number.")


def _history_message(target: int, guesses: list[int]) -> str:
    lines = [f"Your {len(guesses)} previous guesses:"]
    for g in guesses:
        lines.append(f"{g}, which is {'lower' if g < target else 'higher'} than the target number.")
    # the reference builds this message with a `for ... else` whose else-branch always runs (guessing.py:41-46), so
    # every history message ends with this line; kept verbatim: the prompt text is part of the plugin's behaviour
    lines.append("<wrong output>")
    return "\n".join(lines)


async def generate_guessing_rollout(cfg, llm: TrainableLLM, problem: dict, session=None) -> RolloutResult:
    base = [{"role": "system", "content": "You are a helpful assistant"}, {"role": "user", "content": _TASK}]
    started = time.time()
    calls, guesses = [], []
    reward, success, error = 0.0, False, False
    for turn in range(MAX_TURNS):
        messages = list(base)
        if turn > 0:
            messages.append({"role": "user", "content": _history_message(problem["answer"], guesses)})
        call = await llm_async_generate(llm, Prompt(messages=messages), session)
        calls.append(call)
        found = _ANSWER.search(call.output.content or "")
        if not found:
            reward, error = -2 + turn / 10, True
            break
        guess = int(found.group(1))
        if guess == problem["answer"]:
            reward, success = 2 - turn / 10, True
            break
        guesses.append(guess)
    texts = [make_training_text(llm, c) for c in calls]
    for t in texts:
        t.reward = reward
    return RolloutResult(training_texts=texts,
                         metrics=BaseMetrics(reward=reward, success=success, no_error=not error, no_answer=error),
                         latency=time.time() - started, dataset_name=problem["dataset"], domain=DOMAIN)


def load_problems(dataset_names: list[str]) -> list[dict]:
    n, c = 1024, 191
    out = []
    for name in dataset_names:
        if name == "train":
            out += [{"answer": (2 * i * c) % n + 1, "dataset": "train", "domain": DOMAIN} for i in range(512)]
        elif name == "test":
            out += [{"answer": ((2 * i + 1) * c) % n + 1, "dataset": "test", "domain": DOMAIN} for i in range(512)]
    return out
