"""Actor stage: fan rollouts out over the sampler engines and publish finished groups.

Host code mirroring pipelinerl/actor.py:114-286 (`schedule_rollouts`): exactly `attempts` rollouts per
problem form a group; each new rollout goes to the LLM with the fewest running rollouts; results carry the
weight version that was live when the rollout STARTED and the group id; a complete group is handed on.
The shared-memory queues / worker processes of the reference collapse into one asyncio loop because the
engines are in-process; the topic written is the same (`actor`, actor.py:648-652).
"""
from __future__ import annotations

import asyncio
import importlib
import math
import time
from typing import Any, Callable, Iterable

from .llm import TrainableLLM
from .rollouts import RolloutResult
from .streams import SingleStreamSpec, write_to_streams


def get_method(path: str | Callable) -> Callable:
    """Resolve `cfg.actor.rollout_policy` / `cfg.dataset_loader` dotted paths (hydra.utils.get_method, actor.py:141)."""
    if callable(path):
        return path
    module, _, name = path.rpartition(".")
    return getattr(importlib.import_module(module), name)


class LagBudget:
    """How many rollout GROUPS the actor may have submitted so far (the `max_lag` throttle of the training actor loop,
    pipelinerl/actor.py:509-534 for the arithmetic, :551-577 for its use): initially `ceil(max_lag / attempts)` groups
    of head start plus one update's worth, `ceil(total_update_size / attempts)` more every time a NEW weight version is
    observed (once per observation, whatever the jump), where total_update_size is `weight_update_interval` rounded up
    to whole optimizer batches (`train_batch_size * gradient_accumulation_passes`).  max_lag=None disables the throttle."""

    def __init__(self, max_lag: int | None, attempts: int, train_batch_size: int, gradient_accumulation_passes: int,
                 weight_update_interval: int, get_model_version: Callable[[], int]):
        self.get_model_version = get_model_version
        self.submitted = 0
        self.last_version = get_model_version()
        if max_lag is None:
            self.groups_per_update, self.can_submit = None, math.inf
            return
        total_batch = train_batch_size * gradient_accumulation_passes
        total_update = math.ceil(weight_update_interval / total_batch) * total_batch
        self.groups_per_update = math.ceil(total_update / attempts)
        self.lag_groups = math.ceil(max_lag / attempts)
        self.can_submit = self.lag_groups + self.groups_per_update

    def observe(self) -> None:
        v = self.get_model_version()
        if v > self.last_version:
            if self.groups_per_update is not None:
                self.can_submit += self.groups_per_update
            self.last_version = v

    def try_submit(self) -> bool:
        """True (and counted) when one more group may be submitted now."""
        self.observe()
        if self.submitted >= self.can_submit:
            return False
        self.submitted += 1
        return True


async def schedule_rollouts(cfg: Any, attempts: int, problems: Iterable[dict], llms: list[TrainableLLM],
                            rollout_policy: str | Callable, on_group: Callable[[list[RolloutResult]], None],
                            get_model_version: Callable[[], int] = lambda: 0, max_rollouts_per_llm: int = 64,
                            scheduler_name: str = "actor", max_retries: int = 3,
                            lag_budget: "LagBudget | None" = None) -> dict:
    policy = get_method(rollout_policy)
    active = [0] * len(llms)
    groups: dict[int, list[RolloutResult]] = {}
    stats = {"started": 0, "finished": 0, "groups": 0, "output_tokens": 0, "t0": time.time()}
    tasks: set[asyncio.Task] = set()

    async def one(problem: dict, group_id: int, rollout_index: int, li: int):
        version = get_model_version()
        full_gid = f"{scheduler_name}_{group_id}"
        attempt = 0
        try:
            while True:
                try:
                    res: RolloutResult = await policy(cfg, llms[li], problem, None)
                    break
                except (asyncio.TimeoutError, TimeoutError):
                    attempt += 1
                    if attempt > max_retries:
                        raise
                    await asyncio.sleep(min(30.0, 2.0 ** attempt * 0.1))
            res.model_version = version
            res.group_id = full_gid
            for step_index, t in enumerate(res.training_texts):
                t.group_id = full_gid
                t.metadata.update(model_version=version, rollout_index=rollout_index, step_index=step_index)
                stats["output_tokens"] += t.output_tokens
            groups[group_id].append(res)
            if len(groups[group_id]) == attempts:
                stats["groups"] += 1
                on_group(groups.pop(group_id))
        finally:
            active[li] -= 1
            stats["finished"] += 1

    gid = 0
    for problem in problems:
        while lag_budget is not None and not lag_budget.try_submit():
            await asyncio.sleep(0.01)        # throttled by max_lag: wait for the next weight version (actor.py:566)
        groups[gid] = []
        for r in range(attempts):
            while min(active) >= max_rollouts_per_llm:
                done, _ = await asyncio.wait(tasks, return_when=asyncio.FIRST_COMPLETED)
                tasks -= done
                for d in done:
                    d.result()
            li = active.index(min(active))
            active[li] += 1
            stats["started"] += 1
            t = asyncio.create_task(one(problem, gid, r, li))
            tasks.add(t)
        gid += 1
    if tasks:
        done, _ = await asyncio.wait(tasks)
        for d in done:
            d.result()
    stats["seconds"] = time.time() - stats.pop("t0")
    stats["output_tokens_per_second"] = stats["output_tokens"] / max(stats["seconds"], 1e-9)
    return stats


def publish_groups_to_stream(exp_path, topic: str = "actor"):
    """-> (writer context, on_group callback) that writes `[TrainingText.model_dump(), ...]` per group."""
    spec = SingleStreamSpec(exp_path=exp_path, topic=topic)
    writer = write_to_streams(spec)

    def on_group(group: list[RolloutResult]) -> None:
        w.write([t.model_dump() for r in group for t in r.training_texts])

    w = writer.__enter__()
    return writer, on_group
