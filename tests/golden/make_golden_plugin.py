"""Plugin-surface drop-in recording: the REFERENCE's example plugin (pipelinerl/domains/guessing/guessing.py, loaded
unmodified from /root/reference) runs on THIS package's LLM handle: its `from pipelinerl.async_llm import
llm_async_generate, make_training_text`, `from pipelinerl.llm import Prompt, TrainableLLM` and `from pipelinerl.rollouts
import BaseMetrics, RolloutResult` are resolved to pipelinerl_b200's modules (what "switching the package" means for a
plugin author).  The sampler behind it is a scripted fake registered under an inproc:// address.

    python tests/golden/make_golden_plugin.py      (authoring container only)

Stored in tests/golden/plugin_guessing.json: per scenario the messages of every LLM call the reference plugin made and
the RolloutResult it returned; tests/test_host_logic.py::test_guessing_plugin_matches_reference_plugin replays the same
scripts through pipelinerl_b200.domains.guessing and compares.
"""
from __future__ import annotations

import asyncio
import importlib.util
import json
import sys
import types
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
from tests.helpers import ScriptedSampler, ScriptedTokenizer  # noqa: E402

OUT = Path(__file__).resolve().parent
SCENARIOS = [
    dict(name="found_on_4th_guess", answer=383, script=[512, 256, 384, 383]),
    dict(name="first_guess_correct", answer=7, script=[7]),
    dict(name="malformed_on_3rd_turn", answer=100, script=[512, 50, None]),
    dict(name="never_found", answer=1000, script=[1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13]),
]


def load_reference_plugin():
    import pipelinerl_b200.async_llm as my_async
    import pipelinerl_b200.llm as my_llm
    import pipelinerl_b200.rollouts as my_rollouts
    pkg = types.ModuleType("pipelinerl")
    pkg.__path__ = []
    sys.modules["pipelinerl"] = pkg
    sys.modules["pipelinerl.async_llm"] = my_async
    sys.modules["pipelinerl.llm"] = my_llm
    sys.modules["pipelinerl.rollouts"] = my_rollouts
    om = types.ModuleType("omegaconf")
    om.DictConfig = dict
    sys.modules.setdefault("omegaconf", om)
    spec = importlib.util.spec_from_file_location("ref_guessing", "/root/reference/pipelinerl/domains/guessing/guessing.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def run_scenario(plugin_fn, sc):
    from pipelinerl_b200.llm import TrainableLLM
    sampler = ScriptedSampler("golden", sc["script"])
    llm = TrainableLLM(base_url=sampler.base_url, model_name="scripted", tokenizer_name="scripted",
                       parameters={"max_tokens": 8, "temperature": 1.0}, collect_logprobs=True)
    llm.tokenizer = ScriptedTokenizer()
    try:
        res = asyncio.new_event_loop().run_until_complete(
            plugin_fn({}, llm, {"answer": sc["answer"], "dataset": "train", "domain": "guessing"}, None))
    finally:
        sampler.close()
    return {"calls": sampler.prompts_seen,
            "result": {"metrics": res.metrics.model_dump(), "dataset_name": res.dataset_name, "domain": res.domain,
                       "training_texts": [{k: getattr(t, k) for k in ("text", "n_predicted", "input_ids", "labels",
                                                                      "logprobs", "reward", "finished",
                                                                      "prompt_tokens", "output_tokens")}
                                          for t in res.training_texts]}}


def main():
    ref = load_reference_plugin()
    out = []
    for sc in SCENARIOS:
        rec = run_scenario(ref.generate_guessing_rollout, sc)
        out.append({"scenario": sc, **rec})
        print(sc["name"], "turns", len(rec["calls"]), "reward", rec["result"]["metrics"]["reward"])
    out.append({"load_problems": {"train_first": ref.load_problems(["train"])[:3], "test_first": ref.load_problems(["test"])[:3],
                                  "n": len(ref.load_problems(["train", "test"]))}})
    (OUT / "plugin_guessing.json").write_text(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
