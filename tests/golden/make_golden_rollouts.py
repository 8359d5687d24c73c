"""Golden for the rollout record types (pipelinerl/rollouts.py:6-110): the reference classes and helpers are imported
(pydantic only) and their dumps / derived values recorded; tests build the same objects with pipelinerl_b200.rollouts.

    python tests/golden/make_golden_rollouts.py      (authoring container only)
"""
import json
import sys
from pathlib import Path

OUT = Path(__file__).resolve().parent
TEXTS = [dict(text="prompt words answer", n_predicted=6, input_ids=[1, 2, 3, 4], labels=[-100, -100, 3, 4], logprobs=[-0.5, -1.0],
              finished=True, prompt_tokens=2, output_tokens=2, metadata={"model_version": 4}),
         dict(text="cut off", n_predicted=3, input_ids=[5, 6], labels=[-100, 6], logprobs=[-2.0], finished=False,
              prompt_tokens=1, output_tokens=1),
         dict(text="zero predicted", n_predicted=0, finished=True)]


def main():
    sys.path.insert(0, "/root/reference")
    import pipelinerl.rollouts as ref
    texts = [ref.TrainingText(**t) for t in TEXTS]
    rec = {"texts": TEXTS,
           "dumps": [t.model_dump() for t in texts],
           "prompt_text": [t.prompt_text for t in texts], "output_text": [t.output_text for t in texts],
           "has_overflow_all": ref.rollout_has_overflow(texts), "has_overflow_first": ref.rollout_has_overflow(texts[:1]),
           "after_reward": [t.reward for t in ref.apply_rollout_reward([ref.TrainingText(**t) for t in TEXTS], 1.25)]}
    s = ref.summarize_training_texts(texts)
    rec["summary"] = {"prompt_tokens": s.prompt_tokens, "output_tokens": s.output_tokens, "overflow": s.overflow,
                      "num_turns": getattr(s, "num_turns", None)}
    rr = ref.RolloutResult(training_texts=texts[:1], latency=0.5,
                           metrics=ref.BaseMetrics(reward=1.0, success=True, no_error=True, no_answer=False))
    rec["rollout_result_dump"] = rr.model_dump()
    (OUT / "rollouts_cases.json").write_text(json.dumps(rec, indent=1))
    print(rec["summary"], rec["has_overflow_all"], rec["after_reward"])


if __name__ == "__main__":
    main()
