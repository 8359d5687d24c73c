"""Product host code (no GPU needed) against the reference's own outputs: bit-exact layout contract."""
import copy
import json

import numpy as np
import pytest
import torch

from pipelinerl_b200.finetune.data import collate, collate_packed, preprocess_fn
from pipelinerl_b200.finetune.rl import RLConfig, populate_rl_data, prepare_rl_fields
from pipelinerl_b200.finetune.types import PipelineBatchEncoding
from pipelinerl_b200.finetune.utils import create_sentinel_batch
from tests.helpers import GOLDEN, load_rl_case


class Tok:
    eos_token_id = 7
    padding_side = "right"


@pytest.fixture(scope="module")
def cases():
    return json.loads((GOLDEN / "preprocess_cases.json").read_text())


def _entries(case):
    cfg = RLConfig(**case["config"])
    entries = []
    for s in copy.deepcopy(case["raw_samples"]):
        enc = preprocess_fn(s, Tok(), seq_length=10_000, is_rl=True)
        for k in ("group_id", "rollout_index", "step_index", "finished"):
            enc[k] = s[k]
        if "finish_reason" in s:
            enc["finish_reason"] = s["finish_reason"]
        enc["model_version"] = s["metadata"]["model_version"]
        entries.append(enc)
    return populate_rl_data(entries, Tok.eos_token_id, cfg), cfg


@pytest.mark.parametrize("name", ["loo_std", "loo_nostd", "loo_std_sp4"])
def test_populate_and_pack_match_reference(cases, name):
    case = cases[name]
    entries, _ = _entries(case)
    for got, want in zip(entries, case["entries"]):
        for k, w in want.items():
            if isinstance(w, list) and w and isinstance(w[0], float):
                np.testing.assert_allclose(got[k], w, rtol=1e-12, atol=0, err_msg=k)
            else:
                assert got[k] == w, k
    batch = collate_packed(entries, Tok(), seq_parallel=case["seq_parallel"])
    for k, w in case["batch"].items():
        g = getattr(batch, k)
        if isinstance(g, torch.Tensor):
            want_t = torch.tensor(w, dtype=g.dtype)
            assert g.shape == want_t.shape, k
            assert torch.equal(g, want_t), f"{k} differs"
        else:
            assert g == w, k
    assert batch.seq_boundaries.dtype == torch.int32 and batch.input_ids.dtype == torch.int64
    assert batch.rewards.dtype == torch.float32


def test_prepare_rl_fields_alignment_and_assert():
    enc = {"input_ids": [1, 2, 3, 4], "labels": [-100, -100, 3, 4]}
    out = prepare_rl_fields(enc, 0.5, [-1.0, -2.0], [-1.5, -2.5])
    assert out["old_logprobs"] == [0, 0, -1.0, -2.0] and out["ref_logprobs"] == [0, 0, -1.5, -2.5]
    assert out["rewards"] == [0.5] * 4 and out["num_labels"] == [0, 0, 1, 1]
    with pytest.raises(AssertionError):
        prepare_rl_fields({"input_ids": [1, 2], "labels": [-100, 2]}, 1.0, [-1.0, -1.0], [-1.0, -1.0])


def test_sentinel_batch_matches_reference():
    arrs, _ = load_rl_case("sentinel")
    sb = create_sentinel_batch("cpu", tokenizer=Tok(), model_version=5)
    for k in ("input_ids", "attention_mask", "labels", "position_ids", "segment_ids", "rewards", "advantages",
              "ref_logprobs", "old_logprobs", "group_tokens", "num_labels", "overflow", "seq_boundaries"):
        assert torch.equal(getattr(sb, k), torch.from_numpy(arrs[k])), k
    assert sb.sentinel and sb.is_packed and sb.model_version == 5


def test_collate_unpacked_matches_reference(cases):
    arrs, _ = load_rl_case("unpacked")
    # rebuild the same entries the generator used (seed 77 samples are stored only via the batch), so
    # check structural properties + padding rules on a fresh case instead
    case = cases["loo_nostd"]
    entries, _ = _entries(case)
    keep = ["input_ids", "labels", "attention_mask", "rewards", "advantages", "old_logprobs", "ref_logprobs",
            "overflow", "group_tokens", "num_labels", "model_version"]
    b = collate([{k: e[k] for k in keep} for e in entries[:3]], Tok())
    L = b.input_ids.shape[1]
    assert L % 16 == 0 and b.input_ids.shape[0] == 3 and not b.is_packed
    n0 = len(entries[0]["input_ids"])
    assert b.labels[0, n0:].eq(-100).all() and b.input_ids[0, n0:].eq(0).all() and b.rewards[0, n0:].eq(0).all()
    assert b.attention_mask[0, :n0].eq(1).all() and b.attention_mask[0, n0:].eq(0).all()
    assert arrs["input_ids"].shape[1] % 16 == 0  # same rule in the reference fixture


def test_make_slices_and_errors(cases):
    case = cases["loo_std_sp4"]
    entries, _ = _entries(case)
    batch = collate_packed(entries, Tok(), seq_parallel=4)
    T = batch.input_ids.shape[1]
    assert T % 4 == 0
    parts = batch.make_slices(4)
    assert torch.equal(torch.cat([p.input_ids for p in parts], dim=1), batch.input_ids)
    assert torch.equal(torch.cat([p.advantages for p in parts], dim=1), batch.advantages)
    assert all(p.model_version == batch.model_version and p.padding == batch.padding for p in parts)
    with pytest.raises(ValueError):
        batch.make_slices(T + 1)
    with pytest.raises(ValueError):
        batch.make_slices(7 if T % 7 else 9)


def test_empty_and_single_sample_groups():
    cfg = RLConfig(divide_advantage_by_std=True)
    assert populate_rl_data([], 7, cfg) == []
    e = prepare_rl_fields({"input_ids": [9, 9, 7], "labels": [-100, 9, 7]}, 1.0, [-0.1, -0.2], [-0.1, -0.2])
    e.update(group_id="g", rollout_index=0, step_index=0, finished=True)
    out = populate_rl_data([e], 7, cfg)[0]
    assert out["advantages"] == [0.0, 0.0, 0.0]  # single member: baseline = own reward, std NaN -> 0
    assert out["group_tokens"] == [3.0] * 3 and out["num_labels"] == [2] * 3 and out["overflow"] == [0.0] * 3


def test_native_body_segment_bounds_from_position_ids():
    """Packed rows: every restart of position_ids opens a new causal segment (collate_packed layout, data.py:215-283)."""
    import torch
    from pipelinerl_b200.learner_body import NativeBody
    pos = torch.tensor([0, 1, 2, 0, 1, 0, 0, 1, 2, 3])
    assert NativeBody.segment_bounds(pos) == [(0, 3), (3, 5), (5, 6), (6, 10)]
    assert NativeBody.segment_bounds(torch.arange(7)) == [(0, 7)]
    # a row that does not start at position 0 (sequence-parallel slice): the head of the row is its own segment
    assert NativeBody.segment_bounds(torch.tensor([5, 6, 0, 1])) == [(0, 2), (2, 4)]


def test_make_training_text_matches_reference():
    """Row a2 (integer path, bit-exact): tests/golden/training_text_cases.json holds the outputs of the reference's
    make_training_text (pipelinerl/async_llm.py:215-346) on a locally built chat tokenizer — stop / length / eos in the
    content / chat-template kwargs / tools + tool calls / bos stripping.  Same inputs through this package's function."""
    import json
    from pipelinerl_b200.async_llm import make_training_text
    from pipelinerl_b200.llm import LLMCall, LLMOutput, Prompt, TokenLogprob, TrainableLLM
    from tests.helpers import GOLDEN, tiny_chat_tokenizer
    tok = tiny_chat_tokenizer()
    cases = json.loads((GOLDEN / "training_text_cases.json").read_text())
    assert len(cases) >= 6
    for item in cases:
        c, want = item["case"], item["expected"]
        llm = TrainableLLM(base_url="inproc://none", model_name="tiny", tokenizer_name="tiny",
                           parameters={"max_tokens": 8}, collect_logprobs=True,
                           chat_template_kwargs=c.get("chat_template_kwargs"))
        llm.tokenizer = tok
        tcs = [{"id": t["id"], "function": {"name": t["name"], "arguments": t["arguments"]}}
               for t in c.get("tool_calls", [])] or None
        call = LLMCall(prompt=Prompt(messages=c["messages"], tools=c.get("tools")),
                       output=LLMOutput(content=c["content"], tool_calls=tcs),
                       prompt_length_tokens=c["prompt_len"], output_length_tokens=c["out_len"],
                       llm_info={"finish_reason": c["finish_reason"]} if c["finish_reason"] else {},
                       logprobs=[TokenLogprob(logprob=-0.25 * (i + 1), token_id=t) for i, t in enumerate(c["gen"])])
        got = make_training_text(llm, call)
        for k, v in want.items():
            assert getattr(got, k) == v, (c["name"], k, getattr(got, k), v)


def test_guessing_plugin_matches_reference_plugin():
    """Plugin surface: tests/golden/plugin_guessing.json was recorded by running the REFERENCE's example plugin
    (pipelinerl/domains/guessing/guessing.py, unmodified) on this package's TrainableLLM / llm_async_generate /
    make_training_text / RolloutResult with a scripted sampler (make_golden_plugin.py).  This package's own port of the
    plugin must send the same prompts turn by turn and return the same RolloutResult; load_problems must agree."""
    import asyncio
    import json
    from pipelinerl_b200.domains.guessing import generate_guessing_rollout, load_problems
    from pipelinerl_b200.llm import TrainableLLM
    from tests.helpers import GOLDEN, ScriptedSampler, ScriptedTokenizer
    rec = json.loads((GOLDEN / "plugin_guessing.json").read_text())
    scenarios = [r for r in rec if "scenario" in r]
    assert len(scenarios) >= 4
    for r in scenarios:
        sc = r["scenario"]
        sampler = ScriptedSampler("replay", sc["script"])
        llm = TrainableLLM(base_url=sampler.base_url, model_name="scripted", tokenizer_name="scripted",
                           parameters={"max_tokens": 8, "temperature": 1.0}, collect_logprobs=True)
        llm.tokenizer = ScriptedTokenizer()
        try:
            res = asyncio.new_event_loop().run_until_complete(
                generate_guessing_rollout({}, llm, {"answer": sc["answer"], "dataset": "train", "domain": "guessing"}, None))
        finally:
            sampler.close()
        assert sampler.prompts_seen == r["calls"], sc["name"]          # identical prompts, turn by turn
        want = r["result"]
        assert res.metrics.model_dump() == want["metrics"], sc["name"]
        assert (res.dataset_name, res.domain) == (want["dataset_name"], want["domain"])
        assert len(res.training_texts) == len(want["training_texts"])
        for got, w in zip(res.training_texts, want["training_texts"]):
            for k, v in w.items():
                assert getattr(got, k) == v, (sc["name"], k)
    lp = [r for r in rec if "load_problems" in r][0]["load_problems"]
    assert load_problems(["train"])[:3] == lp["train_first"] and load_problems(["test"])[:3] == lp["test_first"]
    assert len(load_problems(["train", "test"])) == lp["n"]


def test_filter_zero_advantage_groups_matches_reference():
    """preprocess.filter_zero_advantage_groups against the reference function executed on the same entries
    (tests/golden/filter_zero_advantage_cases.json, make_golden_filter.py): same samples kept, in the same order."""
    import json
    from pipelinerl_b200.preprocess import filter_zero_advantage_groups
    from tests.helpers import GOLDEN
    cases = json.loads((GOLDEN / "filter_zero_advantage_cases.json").read_text())
    assert len(cases) == 4
    for c in cases:
        kept, dropped = filter_zero_advantage_groups([dict(e) for e in c["entries"]])
        assert [e["uid"] for e in kept] == c["kept_uids"]
        assert dropped == c["dropped"]


def test_schedule_rollouts_matches_reference_scheduler():
    """Row a3: tests/golden/scheduler_case.json was recorded by EXECUTING the reference's schedule_rollouts
    (pipelinerl/actor.py:114-286) with a scripted policy (make_golden_scheduler.py).  Same script through this package's
    scheduler: identical least-busy routing of every rollout, identical group ids and identical stamping of
    model_version / rollout_index / step_index on every training text."""
    import asyncio
    import json
    from pipelinerl_b200.actor import schedule_rollouts
    from pipelinerl_b200.rollouts import BaseMetrics, RolloutResult, TrainingText
    from tests.helpers import GOLDEN
    rec = json.loads((GOLDEN / "scheduler_case.json").read_text())
    sc = rec["scenario"]
    launches, groups = [], []

    async def policy(cfg, llm, problem, session):
        launches.append({"answer": problem["answer"], "llm": llm.name})
        await asyncio.sleep(0.05)
        texts = [TrainingText(text=f"p{problem['answer']}t{t}", n_predicted=1, input_ids=[1, 2], labels=[-100, 2],
                              logprobs=[-0.5], output_tokens=1, prompt_tokens=1) for t in range(sc["turns"])]
        return RolloutResult(training_texts=texts, latency=0.05, dataset_name=problem["dataset"],
                             metrics=BaseMetrics(reward=1.0, success=True, no_error=True, no_answer=False))

    class L:
        def __init__(self, name):
            self.name = name
    llms = [L(f"llm{i}") for i in range(sc["n_llms"])]
    stats = asyncio.new_event_loop().run_until_complete(
        schedule_rollouts({}, sc["attempts"], sc["problems"], llms, policy, groups.append,
                          get_model_version=lambda: sc["model_version"], max_rollouts_per_llm=sc["llm_max_rollouts"],
                          scheduler_name=sc["scheduler_name"]))
    assert launches == rec["launches"]
    assert stats["started"] == stats["finished"] == len(rec["launches"]) and stats["groups"] == len(rec["groups"])
    got = []
    for g in groups:
        rolls = sorted(g, key=lambda r: r.training_texts[0].metadata["rollout_index"])
        got.append({"group_id": rolls[0].group_id, "n": len(rolls),
                    "rollouts": [{"model_version": r.model_version, "group_id": r.group_id,
                                  "texts": [{"group_id": t.group_id, "metadata": dict(t.metadata), "text": t.text}
                                            for t in r.training_texts]} for r in rolls]})
    assert sorted(got, key=lambda d: d["group_id"]) == rec["groups"]


@pytest.mark.parametrize("kind,warmup,total", [("cosine", 50, 1000), ("cosine", 0, 7), ("linear", 5, 40),
                                               ("constant_with_warmup", 3, None), ("constant", 0, None)])
def test_lr_schedule_matches_transformers_get_scheduler(kind, warmup, total):
    """Row a7: the reference builds its schedule with transformers.get_scheduler (finetune_loop.py:394-399; cosine, 50
    warm-up steps by default).  finetune.optim.get_scheduler must produce the same learning rate at every step."""
    import torch
    import transformers
    from pipelinerl_b200.finetune.optim import get_scheduler

    class Opt:
        def __init__(self, lr):
            self.param_groups = [{"lr": lr}, {"lr": lr * 0.5}]
    mine_opt = Opt(1e-6)
    mine = get_scheduler(kind, mine_opt, warmup, total)
    p = torch.nn.Parameter(torch.zeros(1))
    ref_opt = torch.optim.SGD([{"params": [p], "lr": 1e-6}, {"params": [torch.nn.Parameter(torch.zeros(1))], "lr": 5e-7}])
    ref = transformers.get_scheduler(kind, ref_opt, num_warmup_steps=warmup, num_training_steps=total)
    n = (total or 60) + 5
    for step in range(n):
        got = [g["lr"] for g in mine_opt.param_groups]
        want = [g["lr"] for g in ref_opt.param_groups]
        assert all(abs(a - b) <= 1e-12 * max(1.0, abs(b)) + 1e-18 for a, b in zip(got, want)), (step, got, want)
        ref_opt.step()
        ref.step()
        mine.step()


def test_lag_budget_arithmetic_and_throttle():
    """a3 lag budget (pipelinerl/actor.py:509-534, 551-577): max_lag 16 samples, 8 attempts, batches of 4 x 2 samples,
    weight update every 8 samples -> 2 head-start groups + 1 group per update; one more group per OBSERVED new version."""
    import asyncio
    from pipelinerl_b200.actor import LagBudget, schedule_rollouts
    from pipelinerl_b200.rollouts import BaseMetrics, RolloutResult, TrainingText
    version = {"v": 0}
    b = LagBudget(max_lag=16, attempts=8, train_batch_size=4, gradient_accumulation_passes=2, weight_update_interval=8,
                  get_model_version=lambda: version["v"])
    assert (b.lag_groups, b.groups_per_update, b.can_submit) == (2, 1, 3)
    assert [b.try_submit() for _ in range(4)] == [True, True, True, False]
    version["v"] = 5                      # a jump of several versions still adds ONE update's worth (actor.py:551-556)
    assert [b.try_submit() for _ in range(2)] == [True, False]
    # weight_update_interval is rounded up to whole optimizer batches: 9 samples -> 16 -> 2 groups of 8
    b2 = LagBudget(8, 8, 4, 2, 9, lambda: 0)
    assert (b2.lag_groups, b2.groups_per_update, b2.can_submit) == (1, 2, 3)
    assert LagBudget(None, 8, 4, 2, 8, lambda: 0).try_submit()

    # inside the scheduler: with a budget of 3 groups the 4th problem is only started after the version moved
    version["v"] = 0
    budget = LagBudget(16, 2, 1, 2, 2, lambda: version["v"])      # lag 8 groups?  ceil(16/2)=8 + 1
    budget.can_submit = 3
    started = []

    async def policy(cfg, llm, problem, session):
        started.append(problem["answer"])
        await asyncio.sleep(0.01)
        return RolloutResult(training_texts=[TrainingText(text="x", n_predicted=1)], latency=0.0,
                             metrics=BaseMetrics(reward=0.0, success=False, no_error=True, no_answer=False))

    async def go():
        async def bump():
            await asyncio.sleep(0.2)
            assert sorted(set(started)) == [0, 1, 2]      # three groups ran, the fourth is being held back
            version["v"] = 1
        llm = type("L", (), {"name": "l"})()
        await asyncio.gather(schedule_rollouts({}, 2, [{"answer": i} for i in range(4)], [llm], policy, lambda g: None,
                                               get_model_version=lambda: version["v"], lag_budget=budget), bump())
    asyncio.new_event_loop().run_until_complete(go())
    assert sorted(set(started)) == [0, 1, 2, 3]


def test_rollout_records_match_reference_classes():
    """Row a2 record types: tests/golden/rollouts_cases.json was produced by the reference's own TrainingText /
    RolloutResult / helpers (pipelinerl/rollouts.py:6-110, make_golden_rollouts.py).  Same inputs through
    pipelinerl_b200.rollouts: identical dumps (field names, defaults), prompt/output text slicing incl. the n_predicted == 0
    quirk, overflow rule, reward stamping, summary."""
    import json
    from pipelinerl_b200.rollouts import (BaseMetrics, RolloutResult, TrainingText, apply_rollout_reward,
                                          rollout_has_overflow, summarize_training_texts)
    from tests.helpers import GOLDEN
    rec = json.loads((GOLDEN / "rollouts_cases.json").read_text())
    texts = [TrainingText(**t) for t in rec["texts"]]
    assert [t.model_dump() for t in texts] == rec["dumps"]
    assert [t.prompt_text for t in texts] == rec["prompt_text"]
    assert [t.output_text for t in texts] == rec["output_text"]
    assert rollout_has_overflow(texts) == rec["has_overflow_all"]
    assert rollout_has_overflow(texts[:1]) == rec["has_overflow_first"]
    assert [t.reward for t in apply_rollout_reward([TrainingText(**t) for t in rec["texts"]], 1.25)] == rec["after_reward"]
    s = summarize_training_texts(texts)
    assert {"prompt_tokens": s.prompt_tokens, "output_tokens": s.output_tokens, "overflow": s.overflow,
            "num_turns": s.num_turns} == rec["summary"]
    rr = RolloutResult(training_texts=texts[:1], latency=0.5,
                       metrics=BaseMetrics(reward=1.0, success=True, no_error=True, no_answer=False))
    assert rr.model_dump() == rec["rollout_result_dump"]


# ---- a5 writer: dealing micro-batches to trainer ranks (reference loop executed: make_golden_dealer.py) ----
@pytest.mark.parametrize("name", ["one_trainer", "two_trainers", "three_trainers_sentinels", "four_trainers_sp2"])
def test_micro_batch_dealer_matches_reference_writer_loop(name):
    from collections import deque

    from pipelinerl_b200.preprocess import MicroBatchDealer
    case = json.loads((GOLDEN / "dealer_cases.json").read_text())[name]
    spec = case["spec"]
    writes = []
    dealer = MicroBatchDealer(Tok(), spec["seq_length"], spec["num_trainers"], spec["samples_per_lead_per_step"],
                              write=lambda rank, b: writes.append((rank, b)), seq_parallel=spec["seq_parallel"])
    queue, flags = deque(), []
    for chunk in copy.deepcopy(case["arrivals"]):
        queue.extend(chunk)
        while queue:
            before = (len(queue), len(writes))
            flags.append(dealer.deal(queue))
            if (len(queue), len(writes)) == before:
                break
    assert flags == case["batch_done_flags"]
    assert len(writes) == len(case["writes"])
    for (rank, b), want in zip(writes, case["writes"]):
        assert rank == want["rank"]
        for k, w in want["batch"].items():
            g = getattr(b, k)
            if isinstance(g, torch.Tensor):
                assert torch.equal(g, torch.tensor(w, dtype=g.dtype)), (name, k)
            else:
                assert g == w, (name, k)
    assert dealer.published_samples == case["published_samples"]
    assert {str(k): v for k, v in dealer.samples_per_trainer.items()} == case["samples_per_trainer"]
    assert len(dealer.current_batch) == case["left_in_current_batch"]
    # every lead trainer got the same number of micro-batches per completed step (collectives stay aligned)
    per_rank = {}
    for rank, _ in writes:
        per_rank[rank] = per_rank.get(rank, 0) + 1
    if dealer.trainer_id == 0:
        assert len(set(per_rank.values())) == 1


# ---- f4: checkpoint / resume in safetensors ----
def test_model_checkpoint_roundtrip_hf_names(tmp_path):
    from pipelinerl_b200.finetune.checkpoints import load_model_weights, save_model_only
    from pipelinerl_b200.model import fused_shapes
    from tests.helpers import tiny_cfg, tiny_weights
    cfg = tiny_cfg("gqa2")
    w = {k: v.to(torch.bfloat16) for k, v in tiny_weights(cfg).items()}
    save_model_only(tmp_path / "ckpt", cfg, w.items())
    assert (tmp_path / "ckpt" / "model.safetensors").exists() and not (tmp_path / "ckpt~temp").exists()
    cfgj = json.loads((tmp_path / "ckpt" / "config.json").read_text())
    assert cfgj["num_key_value_heads"] == cfg.num_kv_heads and cfgj["model_type"] == "qwen2"
    from safetensors.torch import load_file
    sd = load_file(str(tmp_path / "ckpt" / "model.safetensors"))
    assert "model.layers.0.self_attn.k_proj.weight" in sd and "model.layers.1.mlp.up_proj.weight" in sd
    assert sd["model.layers.0.self_attn.k_proj.weight"].shape == (cfg.kv_size, cfg.hidden_size)
    back = load_model_weights(tmp_path / "ckpt", cfg)
    for name, _ in fused_shapes(cfg):
        assert torch.equal(back[name], w[name]), name
    # a second save swaps directories atomically and leaves no temp / old siblings behind
    save_model_only(tmp_path / "ckpt", cfg, w.items())
    assert sorted(p.name for p in tmp_path.iterdir()) == ["ckpt"]
    # HF's own Qwen2ForCausalLM opens the directory (what the reference's load_model does, checkpoints.py:151-222)
    from transformers import AutoModelForCausalLM
    m = AutoModelForCausalLM.from_pretrained(str(tmp_path / "ckpt"), torch_dtype=torch.bfloat16)
    assert torch.equal(m.model.layers[1].mlp.down_proj.weight.data, w["layers.1.down_proj.weight"])


def test_tokenizer_is_never_silently_substituted():
    """a missing / misspelt tokenizer must raise; the synthetic tokenizer is used only when asked for by name or
    injected (the reference resolves `tokenizer_name` through AutoTokenizer, llm.py:366-374)."""
    from pipelinerl_b200.llm import SyntheticTokenizer, TrainableLLM
    assert isinstance(TrainableLLM("inproc://x", "m", tokenizer_name="synthetic").load_tokenizer(), SyntheticTokenizer)
    tok = SyntheticTokenizer()
    assert TrainableLLM("inproc://x", "m", tokenizer=tok).load_tokenizer() is tok
    with pytest.raises(Exception):
        TrainableLLM("inproc://x", "no-such-model-anywhere/at-all").load_tokenizer()


def test_unsupported_sampling_parameters_fail_loudly_in_process():
    """top_p / top_k / stop / n are rejected by the in-process client exactly as http_shim answers 400 for them"""
    import asyncio

    from pipelinerl_b200.async_llm import llm_async_generate
    from pipelinerl_b200.llm import Prompt, SyntheticTokenizer, TrainableLLM
    for bad in ({"top_p": 0.9}, {"top_k": 20}, {"stop": ["\n"]}, {"n": 2}, {"repetition_penalty": 1.2}):
        llm = TrainableLLM("inproc://none", "m", parameters={"max_tokens": 4, **bad}, tokenizer=SyntheticTokenizer())
        with pytest.raises(ValueError):
            asyncio.run(llm_async_generate(llm, Prompt(messages=[{"role": "user", "content": "hi"}])))


def test_lag_budget_matches_the_reference_actor_loop():
    """tests/golden/lag_budget_cases.json = the reference's own statements (the `max_lag` arithmetic of pipelinerl/actor.py:509-534,
    the weight-version block and the `blocked_by_lag` test of its loop, :551-577) executed on scripted ticks
    (make_golden_lag_budget.py).  LagBudget must allow exactly the same submissions and hold the same budget after every tick."""
    import json
    import math
    from pathlib import Path
    from pipelinerl_b200.actor import LagBudget
    doc = json.loads((Path(__file__).parent / "golden" / "lag_budget_cases.json").read_text())
    assert len(doc["cases"]) >= 5
    for name, case in doc["cases"].items():
        version = {"v": 0}
        lb = LagBudget(case["config"]["max_lag"], case["config"]["attempts"], case["config"]["train_batch_size"],
                       case["config"]["gradient_accumulation_passes"], case["config"]["weight_update_interval"],
                       lambda: version["v"])
        assert lb.groups_per_update == case["groups_per_update"], name
        total = 0
        for (v, available), want in zip(case["ticks"], case["after_tick"]):
            version["v"] = v
            lb.observe()
            n = 0
            for _ in range(available):
                if not lb.try_submit():
                    break
                n += 1
            total += n
            assert (n, total) == (want["submitted_in_tick"], want["submitted_total"]), (name, v, available, n, want)
            assert (None if lb.can_submit == math.inf else lb.can_submit) == want["can_submit_before_update"], name
