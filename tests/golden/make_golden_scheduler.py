"""Golden for row a3: the reference's `schedule_rollouts` coroutine (pipelinerl/actor.py:114-286) EXECUTED with a scripted
rollout policy, scripted queues and a scripted trainer state.  The function's source is cut out of the reference file
with ast and exec'd in a namespace that provides the names it uses (importing pipelinerl.actor pulls in wandb, hydra,
uvloop ... which are not installed here); the body that runs is the reference's.

    python tests/golden/make_golden_scheduler.py      (authoring container only)

Recorded (tests/golden/scheduler_case.json): the launch order with the LLM each rollout was routed to (least-busy
routing, :268-273), and for every finished group the stamping the downstream stages rely on (:207-219): group id
"<scheduler>_<n>", model_version, rollout_index, step_index on every training text.
"""
import ast
import asyncio
import json
import logging
import random
import sys
import time
import types
from pathlib import Path
from queue import Empty

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
OUT = Path(__file__).resolve().parent
SRC = Path("/root/reference/pipelinerl/actor.py")

SCENARIO = dict(attempts=4, n_llms=3, llm_max_rollouts=64, problems=[{"answer": 11, "dataset": "train"},
                                                                    {"answer": 22, "dataset": "train"},
                                                                    {"answer": 33, "dataset": "train"}],
                model_version=7, scheduler_name="actor0", turns=2)


def reference_schedule_rollouts(policy):
    import aiohttp
    sys.path.insert(0, "/root/reference")
    tree = ast.parse(SRC.read_text())
    fn = next(n for n in tree.body if isinstance(n, ast.AsyncFunctionDef) and n.name == "schedule_rollouts")

    class RetryableAbortedCompletionError(TimeoutError):
        pass
    hydra = types.SimpleNamespace(utils=types.SimpleNamespace(get_method=lambda path: policy))
    ns = dict(asyncio=asyncio, aiohttp=aiohttp, time=time, random=random, logger=logging.getLogger("ref_actor"), Empty=Empty,
              hydra=hydra, calculate_train_steps=lambda ft, interrupt: 10 ** 9,
              RetryableAbortedCompletionError=RetryableAbortedCompletionError,
              DictConfig=dict, SharedMemoryQueue=object, TrainerState=object, TrainableLLM=object)
    exec(compile(ast.Module(body=[fn], type_ignores=[]), str(SRC), "exec"), ns)
    return ns["schedule_rollouts"]


def make_policy(rollouts_mod, launches, turns):
    async def policy(cfg, llm, problem, session):
        launches.append({"answer": problem["answer"], "llm": llm.name})
        await asyncio.sleep(0.05)
        texts = [rollouts_mod.TrainingText(text=f"p{problem['answer']}t{t}", n_predicted=1, input_ids=[1, 2], labels=[-100, 2],
                                           logprobs=[-0.5], output_tokens=1, prompt_tokens=1) for t in range(turns)]
        return rollouts_mod.RolloutResult(training_texts=texts, latency=0.05, dataset_name=problem["dataset"],
                                          metrics=rollouts_mod.BaseMetrics(reward=1.0, success=True, no_error=True,
                                                                           no_answer=False))
    return policy


def summarize(groups):
    out = []
    for g in groups:
        rolls = sorted(g, key=lambda r: r.training_texts[0].metadata["rollout_index"])
        out.append({"group_id": rolls[0].group_id, "n": len(rolls),
                    "rollouts": [{"model_version": r.model_version, "group_id": r.group_id,
                                  "texts": [{"group_id": t.group_id, "metadata": dict(t.metadata), "text": t.text}
                                            for t in r.training_texts]} for r in rolls]})
    return sorted(out, key=lambda d: d["group_id"])


async def main_async():
    sys.path.insert(0, "/root/reference")
    import pipelinerl.rollouts as ref_rollouts   # reference classes (pydantic only)
    sc = SCENARIO
    launches, groups = [], []
    fn = reference_schedule_rollouts(make_policy(ref_rollouts, launches, sc["turns"]))
    pending = list(sc["problems"])

    class PQ:
        def get(self, block=False):
            if not pending:
                raise Empty
            return pending.pop(0)

    class RQ:
        def put(self, item):
            if isinstance(item, Exception):
                raise item
            groups.append(item)

        def max_actual_entry_size(self):
            return 0
    state = types.SimpleNamespace(samples_processed=0, propagated_weight_version=sc["model_version"])
    cfg = types.SimpleNamespace(actor=types.SimpleNamespace(rollout_policy="scripted", llm_max_rollouts=sc["llm_max_rollouts"]),
                                finetune=types.SimpleNamespace(interrupt_train_steps=-1, train_batch_size=1,
                                                               gradient_accumulation_passes=1))
    llms = [types.SimpleNamespace(name=f"llm{i}") for i in range(sc["n_llms"])]

    async def finish_when_done():
        while len(groups) < len(sc["problems"]):
            await asyncio.sleep(0.01)
        state.samples_processed = 10 ** 12
    random.seed(0)
    await asyncio.gather(fn(cfg, sc["attempts"], PQ(), RQ(), state, llms, sc["scheduler_name"]), finish_when_done())
    rec = {"scenario": sc, "launches": launches, "groups": summarize(groups)}
    (OUT / "scheduler_case.json").write_text(json.dumps(rec, indent=1))
    print("launch routing:", [l["llm"] for l in launches])
    print("groups:", [(g["group_id"], g["n"]) for g in rec["groups"]])


if __name__ == "__main__":
    asyncio.new_event_loop().run_until_complete(main_async())
