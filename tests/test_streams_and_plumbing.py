"""Host plumbing (CPU): stream topics, trainer messages, push-slice partitioning incl. a world_size-2 gloo run."""
import json
import os
import subprocess
import sys
import threading
from pathlib import Path

import pytest

from pipelinerl_b200 import streams
from pipelinerl_b200.weights import WeightUpdateSuccess, push_slice

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(autouse=True)
def files_backend():
    streams.reset_streams_backend()
    streams.set_streams_backend("files")
    yield
    streams.reset_streams_backend()


def test_stream_roundtrip_and_layout(tmp_path):
    spec = streams.SingleStreamSpec(exp_path=tmp_path, topic="actor")
    with streams.write_to_streams(spec) as w:
        w.write({"a": 1})
        w.write(WeightUpdateSuccess(version=7))
    assert (tmp_path / "streams" / "actor" / "0" / "0" / "0.jsonl").exists()
    with streams.read_stream(spec) as r:
        got = r.read_available()
    assert got[0] == {"a": 1} and got[1]["kind"] == "weight_update_success" and got[1]["version"] == 7


def test_blocking_tail_sees_later_writes(tmp_path):
    spec = streams.SingleStreamSpec(exp_path=tmp_path, topic="training_data", partition=1)
    seen = []

    def reader():
        with streams.read_stream(spec) as r:
            for x in r.read():
                seen.append(x)
                if len(seen) == 3:
                    return
    with streams.write_to_streams(spec) as w:
        w.write({"i": 0})
        t = threading.Thread(target=reader)
        t.start()
        w.write({"i": 1})
        w.write({"i": 2})
        t.join(timeout=10)
    assert [x["i"] for x in seen] == [0, 1, 2]


def test_round_robin_partitions(tmp_path):
    spec = streams.StreamRangeSpec(exp_path=tmp_path, topic="training_data", partition_range=(0, 3))
    with streams.write_to_streams(spec) as w:
        for i in range(7):
            w.write({"i": i})
        w.write({"i": 99}, partition=2)
    counts = []
    for p in range(3):
        with streams.read_stream(streams.SingleStreamSpec(exp_path=tmp_path, topic="training_data", partition=p)) as r:
            counts.append([x["i"] for x in r.read_available()])
    assert counts == [[0, 3, 6], [1, 4], [2, 5, 99]]


def test_backend_must_be_set(tmp_path):
    streams.reset_streams_backend()
    with pytest.raises(ValueError):
        streams.read_stream(streams.SingleStreamSpec(exp_path=tmp_path, topic="x"))
    with pytest.raises(ValueError):
        streams.set_streams_backend("kafka")


@pytest.mark.parametrize("nbytes,n", [(16, 1), (15_230_000_000 // 16 * 16, 2), (4096, 3), (1600, 7), (32, 4)])
def test_push_slices_cover_arena_exactly(nbytes, n):
    at = 0
    for r in range(n):
        off, ln = push_slice(nbytes, r, n)
        assert off == at and off % 16 == 0 and ln % 16 == 0
        at += ln
    assert at == nbytes


def test_world_size_2_gloo_slices_and_max_reduce(tmp_path):
    """The N>1 host logic (slice ownership, max-over-ranks timing) under torch.distributed gloo, world_size 2."""
    script = tmp_path / "w2.py"
    script.write_text(
        "import os, sys, json, torch, torch.distributed as dist\n"
        f"sys.path.insert(0, {str(ROOT)!r})\n"
        "from pipelinerl_b200.weights import push_slice\n"
        "dist.init_process_group('gloo')\n"
        "r, w = dist.get_rank(), dist.get_world_size()\n"
        "off, ln = push_slice(1 << 20, r, w)\n"
        "t = torch.tensor([off, ln, 10.0 + r], dtype=torch.float64)\n"
        "allt = [torch.zeros(3, dtype=torch.float64) for _ in range(w)]\n"
        "dist.all_gather(allt, t)\n"
        "mx = torch.tensor([10.0 + r], dtype=torch.float64); dist.all_reduce(mx, op=dist.ReduceOp.MAX)\n"
        "if r == 0: print(json.dumps({'slices': [a.tolist() for a in allt], 'max': mx.item()}))\n"
        "dist.destroy_process_group()\n")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29571")
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29571", str(script)],
                         capture_output=True, text=True, env=env, timeout=240)
    assert res.returncode == 0, res.stderr[-2000:]
    out = json.loads([l for l in res.stdout.splitlines() if l.startswith("{")][-1])
    (o0, l0, _), (o1, l1, _) = out["slices"]
    assert o0 == 0 and o1 == l0 and l0 + l1 == (1 << 20) and out["max"] == 11.0


def test_dp_gradient_allreduce_gloo_world2(tmp_path):
    """Learner DP exchange step: SUM all-reduce of the flat gradient arena (gloo, world_size 2)."""
    script = tmp_path / "dp.py"
    script.write_text(
        "import sys, json, torch, torch.distributed as dist\n"
        f"sys.path.insert(0, {str(ROOT)!r})\n"
        "from pipelinerl_b200.finetune_loop import allreduce_gradients\n"
        "dist.init_process_group('gloo')\n"
        "r = dist.get_rank()\n"
        "g = torch.arange(10, dtype=torch.float32) * (r + 1)\n"
        "allreduce_gradients(g)\n"
        "if r == 0: print(json.dumps(g.tolist()))\n"
        "dist.destroy_process_group()\n")
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29577", str(script)],
                         capture_output=True, text=True, timeout=240)
    assert res.returncode == 0, res.stderr[-2000:]
    got = json.loads([l for l in res.stdout.splitlines() if l.startswith("[")][-1])
    assert got == [3.0 * i for i in range(10)]


def test_dealer_feeds_two_gloo_learner_ranks_in_lockstep(tmp_path):
    """Writer + trainer accounting under torch.distributed gloo, world_size 2: MicroBatchDealer deals the golden
    `two_trainers` arrivals through the file streams (topic training_data, partition = rank); each rank reads ITS
    partition and runs StepAccountant.  Both ranks must take their optimizer steps at the same micro-batch index, on
    exactly `samples_per_step` global samples, with sentinels only where the rank already holds its share."""
    script = tmp_path / "deal.py"
    script.write_text(
        "import sys, json, copy, torch, torch.distributed as dist\n"
        "from collections import deque\n"
        f"sys.path.insert(0, {str(ROOT)!r})\n"
        "from pipelinerl_b200 import streams\n"
        "from pipelinerl_b200.preprocess import MicroBatchDealer\n"
        "from pipelinerl_b200.finetune_loop import StepAccountant\n"
        "from pipelinerl_b200.finetune.types import PipelineBatchEncoding\n"
        "class Tok:\n    eos_token_id = 7\n    padding_side = 'right'\n"
        "dist.init_process_group('gloo')\n"
        "r, w = dist.get_rank(), dist.get_world_size()\n"
        f"case = json.load(open({str(ROOT / 'tests' / 'golden' / 'dealer_cases.json')!r}))['two_trainers']\n"
        "spec = case['spec']\n"
        f"streams.set_streams_backend('files')\n"
        f"exp = {str(tmp_path)!r}\n"
        "if r == 0:\n"
        "    out = streams.StreamRangeSpec(exp_path=exp, topic='training_data', partition_range=(0, w))\n"
        "    with streams.write_to_streams(out) as wr:\n"
        "        dealer = MicroBatchDealer(Tok(), spec['seq_length'], w, spec['samples_per_lead_per_step'],\n"
        "                                  write=lambda rank, b: wr.write(b, rank))\n"
        "        q = deque()\n"
        "        for chunk in case['arrivals']:\n"
        "            q.extend(copy.deepcopy(chunk))\n"
        "            while q:\n"
        "                before = (len(q), dealer.published_samples, dealer.trainer_id)\n"
        "                dealer.deal(q)\n"
        "                if (len(q), dealer.published_samples, dealer.trainer_id) == before: break\n"
        "dist.barrier()\n"
        "with streams.read_stream(streams.SingleStreamSpec(exp_path=exp, topic='training_data', partition=r)) as rd:\n"
        "    mine = [PipelineBatchEncoding.from_dict(d) for d in rd.read_available()]\n"
        "n_mb = torch.tensor([len(mine)]); dist.all_reduce(n_mb, op=dist.ReduceOp.MIN)\n"
        "acct = StepAccountant(spec['samples_per_lead_per_step'] * w)\n"
        "steps, sent = [], 0\n"
        "for i, b in enumerate(mine[:int(n_mb)]):\n"
        "    n = 0 if b.sentinel else int(b.seq_boundaries.numel()) - 1 - (1 if b.padding else 0)\n"
        "    sent += int(bool(b.sentinel))\n"
        "    total, do_step = acct.observe(n, bool(b.sentinel))\n"
        "    if do_step: steps.append((i, total))\n"
        "allsteps = [None] * w; dist.all_gather_object(allsteps, (steps, sent, len(mine)))\n"
        "if r == 0: print(json.dumps(allsteps))\n"
        "dist.destroy_process_group()\n")
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29579", str(script)],
                         capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stderr[-3000:]
    (s0, sent0, n0), (s1, sent1, n1) = json.loads([l for l in res.stdout.splitlines() if l.startswith("[")][-1])
    assert s0 == s1 and len(s0) >= 2                      # same micro-batch index, same global sample count
    assert [t for _, t in s0] == [6 * (k + 1) for k in range(len(s0))]
    assert sent0 + sent1 >= 1 and abs(n0 - n1) <= 1


def test_sequence_parallel_ranks_share_one_sample_count(tmp_path):
    """seq_parallel = 2 on a gloo world of 2: both ranks hold slices of the SAME micro-batch and each counts its samples;
    the accountant divides the summed count by seq_parallel (reference finetune_loop.py:628,709-712) and both ranks step at
    the same micro-batch.  The dealer cuts a micro-batch into `seq_parallel` slices for ranks lead .. lead + sp - 1."""
    script = tmp_path / "sp.py"
    script.write_text(
        "import sys, json, torch, torch.distributed as dist\n"
        f"sys.path.insert(0, {str(ROOT)!r})\n"
        "from pipelinerl_b200.finetune_loop import StepAccountant, build_seq_parallel_group\n"
        "dist.init_process_group('gloo')\n"
        "r, w = dist.get_rank(), dist.get_world_size()\n"
        "g = build_seq_parallel_group(None, 2)\n"
        "assert dist.get_world_size(g) == 2\n"
        "acct = StepAccountant(6, seq_parallel=2)\n"
        "assert acct.samples_per_lead_per_step == 6\n"
        "steps = []\n"
        "for i, n in enumerate([2, 3, 1, 4, 2]):\n"
        "    total, do_step = acct.observe(n, False)\n"
        "    if do_step: steps.append((i, total))\n"
        "out = [None] * w; dist.all_gather_object(out, steps)\n"
        "if r == 0: print(json.dumps(out))\n"
        "dist.destroy_process_group()\n")
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29581", str(script)],
                         capture_output=True, text=True, timeout=240)
    assert res.returncode == 0, res.stderr[-3000:]
    s0, s1 = json.loads([l for l in res.stdout.splitlines() if l.startswith("[")][-1])
    assert s0 == s1 == [[2, 6], [4, 12]]


def test_sequence_parallel_segment_description():
    """NativeBody.sp_segments: a slice of a packed row -> (local start, length, position of the first local query, global
    row of the sequence's first key) per local segment, checked against the row's own sample boundaries"""
    import torch
    from pipelinerl_b200.learner_body import NativeBody
    lens = [130, 17, 300, 1, 64]
    pos = torch.cat([torch.arange(l) for l in lens])
    T = pos.numel()
    starts = [sum(lens[:i]) for i in range(len(lens))]
    for sp in (1, 2, 4):
        Tl = T // sp
        for r in range(sp):
            a = r * Tl
            qs, ql, p0, kvs, mq, mkv = NativeBody.sp_segments(pos[a:a + Tl], a, "cpu")
            assert int(ql.sum()) == Tl and mq == int(ql.max()) and mkv == int((p0 + ql).max())
            for s, l, p, k in zip(qs.tolist(), ql.tolist(), p0.tolist(), kvs.tolist()):
                first = max(x for x in starts if x <= a + s)
                assert p == a + s - first and k == first
                assert pos[a + s:a + s + l].tolist() == list(range(p, p + l))


def test_trainer_state_follows_topic(tmp_path):
    from pipelinerl_b200.state import TrainerState
    from pipelinerl_b200.weights import SamplesProcessed, TrainingDone, TRAINER_TOPIC
    spec = streams.SingleStreamSpec(exp_path=tmp_path, topic=TRAINER_TOPIC)
    with streams.write_to_streams(spec) as w:
        w.write(SamplesProcessed(samples_processed=5))
        st = TrainerState(tmp_path)
        st.start_listening()
        assert st.wait_for_processed_samples() == 5
        w.write(WeightUpdateSuccess(version=16))
        assert st.wait_for_model_version() == 16
        w.write(TrainingDone())
        assert st.wait_for_training_done(timeout=5) and st.training_done
        st.stop()


def test_file_streams_interoperate_with_the_reference_backend(tmp_path):
    """tests/golden/streams_tree.json = the tree the REFERENCE's file backend wrote (make_golden_streams.py).  (1) this
    package's writer, given the same writes, produces the same files (same relative paths, JSON-equal lines in the same
    order: round-robin and explicit partitions, pydantic models, numpy / tensor payloads); (2) this package's reader,
    pointed at the reference-written tree, yields what the reference's own reader yielded."""
    import json
    import numpy as np
    import torch
    from pydantic import BaseModel
    from pipelinerl_b200 import streams
    from tests.helpers import GOLDEN
    rec = json.loads((GOLDEN / "streams_tree.json").read_text())

    class Success(BaseModel):
        kind: str = "weight_update_success"
        version: int
        timestamp: float

    class WithTensor(BaseModel):
        model_config = {"arbitrary_types_allowed": True}
        name: str
        values: torch.Tensor
    g1 = [{"text": "a b", "input_ids": [1, 2, 3], "logprobs": [-0.5, -0.25], "reward": 1.0, "group_id": "actor0_0",
           "metadata": {"model_version": 7, "rollout_index": 0, "step_index": 0}, "finished": True}]
    g2 = [{"text": "é ü", "input_ids": [4], "logprobs": [-1.5], "reward": 0.0, "group_id": "actor0_1",
           "metadata": {"model_version": 7, "rollout_index": 1, "step_index": 0}, "finished": False}]
    writes = [
        ("actor", dict(instance=0, partition=0), [(g1, None), (g2, None)]),
        ("training_data", dict(instance=0, partition_range=(0, 2)),
         [({"i": i, "arr": (np.arange(3) + i)}, None) for i in range(5)] + [({"i": 99, "arr": np.zeros(2)}, 1)]),
        ("weight_update_request", dict(instance=0, partition=0),
         [(Success(version=3, timestamp=12.5), None), (WithTensor(name="t", values=torch.arange(4, dtype=torch.float32)), None)]),
    ]
    streams.reset_streams_backend()
    streams.set_streams_backend("files")
    try:
        mine = tmp_path / "mine"
        for topic, kw, items in writes:
            spec = streams.StreamRangeSpec(exp_path=mine, topic=topic, **kw) if "partition_range" in kw \
                else streams.SingleStreamSpec(exp_path=mine, topic=topic, **kw)
            with streams.write_to_streams(spec) as w:
                for payload, part in items:
                    w.write(payload, partition=part) if part is not None else w.write(payload)
        tree = {str(p.relative_to(mine)): p.read_text(encoding="utf-8") for p in sorted(mine.rglob("*.jsonl"))}
        assert sorted(tree) == sorted(rec["tree"])
        for rel, text in rec["tree"].items():
            want = [json.loads(line) for line in text.splitlines()]
            got = [json.loads(line) for line in tree[rel].splitlines()]
            assert got == want, rel
        # reference-written tree -> this package's reader
        theirs = tmp_path / "theirs"
        for rel, text in rec["tree"].items():
            f = theirs / rel
            f.parent.mkdir(parents=True, exist_ok=True)
            f.write_text(text, encoding="utf-8")
        with streams.read_stream(streams.SingleStreamSpec(exp_path=theirs, topic="actor")) as r:
            assert r.read_available() == rec["actor_read_back"]
    finally:
        streams.reset_streams_backend()


def test_trainer_messages_are_what_the_reference_parses():
    """tests/golden/trainer_messages.json: dumps of this package's trainer messages that the REFERENCE's TrainerMessage
    union (finetune_loop.py:138-171, as state.py:35-47 uses it) parsed into the class of the same name with the same
    fields.  The dumps must stay what was recorded; the topic name must stay the reference's."""
    import json
    from pipelinerl_b200 import weights as w
    from tests.helpers import GOLDEN
    rec = json.loads((GOLDEN / "trainer_messages.json").read_text())
    assert rec["topic"] == "weight_update_request"
    mk = {"WeightUpdateSuccess": lambda: w.WeightUpdateSuccess(version=12, timestamp=3.5),
          "SamplesProcessed": lambda: w.SamplesProcessed(samples_processed=640, timestamp=4.5),
          "TrainingDone": lambda: w.TrainingDone(timestamp=5.5),
          "WeightUpdateRequest": lambda: w.WeightUpdateRequest(
              version=13, timestamp=6.5, parameters_info=[w.ParameterInfo(name="w", shape=[2, 3], dtype="bfloat16")])}
    assert len(rec["messages"]) == 4
    for m in rec["messages"]:
        assert m["reference_class"] == m["ours"]
        assert mk[m["ours"]]().model_dump() == m["dump"] == m["reference_fields"]


def test_config1_plumbing_end_to_end_on_cpu(tmp_path):
    """BASELINE configs[0]-style plumbing run without a GPU: the guessing plugin on a scripted sampler -> scheduler with
    groups of 4 -> `actor` topic on the file stream backend -> preprocess (RL columns, leave-one-out advantages) ->
    packed micro-batches -> the ORACLE learner (oracle/learner_oracle + pg_oracle: the CPU stand-in for the CUDA hot
    path) -> AdamW oracle.  Everything between the sampler and the loss is this package's host code."""
    import asyncio
    import numpy as np
    import torch
    from oracle import adamw_oracle, learner_oracle, pg_oracle
    from pipelinerl_b200 import streams
    from pipelinerl_b200.actor import publish_groups_to_stream, schedule_rollouts
    from pipelinerl_b200.domains.guessing import generate_guessing_rollout, load_problems
    from pipelinerl_b200.finetune.rl import RLConfig
    from pipelinerl_b200.llm import TrainableLLM
    from pipelinerl_b200.preprocess import pack_micro_batches, preprocess_dataset
    from tests.helpers import ScriptedSampler, ScriptedTokenizer, tiny_cfg, tiny_weights

    class Tok(ScriptedTokenizer):
        eos_token_id = 2
    streams.reset_streams_backend()
    streams.set_streams_backend("files")
    sampler = ScriptedSampler("cfg1", [512, 256, 128, None])     # three wrong guesses, then a malformed answer
    try:
        llm = TrainableLLM(base_url=sampler.base_url, model_name="scripted", tokenizer_name="scripted",
                           parameters={"max_tokens": 8, "temperature": 1.0}, collect_logprobs=True)
        llm.tokenizer = Tok()
        writer, on_group = publish_groups_to_stream(tmp_path)
        problems = load_problems(["train"])[:2]
        stats = asyncio.new_event_loop().run_until_complete(
            schedule_rollouts({}, 4, problems, [llm], generate_guessing_rollout, on_group, get_model_version=lambda: 3,
                              scheduler_name="actor0"))
        writer.__exit__(None, None, None)
        assert stats["groups"] == 2 and stats["started"] == 8
        with streams.read_stream(streams.SingleStreamSpec(exp_path=tmp_path, topic="actor")) as r:
            groups = r.read_available()
        assert len(groups) == 2 and all(g[0]["metadata"]["model_version"] == 3 for g in groups)
        samples = [s for g in groups for s in g]
        rcfg = RLConfig(policy_loss="ppo", kl_coef=0.0, final_kl_coef=0.0, batch_size=len(samples))
        entries = preprocess_dataset(samples, Tok(), seq_length=4096, rl_config=rcfg)
        assert len(entries) == len(samples) and all(len(e["advantages"]) == len(e["input_ids"]) for e in entries)
        batches = pack_micro_batches(entries, Tok(), seq_length=1024, samples_per_step=len(samples))
        assert batches and all(b.is_packed for b in batches)
        # the oracle learner consumes the packed rows this package produced
        cfg = tiny_cfg("gqa2")
        w = {k: v.float() for k, v in tiny_weights(cfg).items()}
        ocfg = pg_oracle.OracleRLConfig.from_dict(rcfg.model_dump())
        total, grads = 0.0, None
        for b in batches:
            cols = {k: getattr(b, k)[0] for k in ("input_ids", "labels", "position_ids", "segment_ids", "rewards",
                                                  "advantages", "ref_logprobs", "old_logprobs", "group_tokens",
                                                  "num_labels", "overflow")}
            cols["input_ids"] = cols["input_ids"] % cfg.vocab_size
            cols["labels"] = torch.where(cols["labels"] >= 0, cols["input_ids"], cols["labels"])
            loss, st, lp, g = learner_oracle.learner_step(cfg, w, cols, ocfg, 0, 10)
            assert np.isfinite(loss) and all(torch.isfinite(v).all() for v in g.values())
            total += loss
            grads = g if grads is None else {k: grads[k] + v for k, v in g.items()}
        names = sorted(grads)
        p = [w[n].reshape(-1).numpy().copy() for n in names]
        before = [x.copy() for x in p]
        m, v = [np.zeros_like(x) for x in p], [np.zeros_like(x) for x in p]
        norm = adamw_oracle.adamw_step(p, [grads[n].reshape(-1).numpy() for n in names], m, v, names, 1, 1e-3, 0.01,
                                       max_grad_norm=0.3)
        assert np.isfinite(norm) and norm > 0
        assert any(np.abs(a - b).max() > 0 for a, b in zip(p, before))
    finally:
        sampler.close()
        streams.reset_streams_backend()


def test_prefetch_batches_bounds_lookahead_and_forwards_errors():
    """Loader thread of the trainer (finetune_loop.py:494-505): at most `maxsize` finished batches wait in the queue (plus
    the one the producer is blocked on), order is preserved, and a source exception reaches the consumer."""
    import threading
    import time
    from pipelinerl_b200.finetune_loop import prefetch_batches
    produced = []

    def source(n, fail_at=None):
        for i in range(n):
            if i == fail_at:
                raise RuntimeError("stream broke")
            produced.append(i)
            yield i
    got = []
    it = prefetch_batches(source(6), maxsize=1)
    first = next(it)
    time.sleep(0.2)                       # the loader may run ahead by the queue slot + the item it is blocked on
    assert first == 0 and len(produced) <= 3
    got = [first] + list(it)
    assert got == list(range(6))
    produced.clear()
    with pytest.raises(RuntimeError, match="stream broke"):
        list(prefetch_batches(source(5, fail_at=3)))
    assert threading.active_count() < 50


def test_trainer_state_listener_agrees_with_the_reference_listener(tmp_path):
    """tests/golden/trainer_state_case.json: a `weight_update_request` topic written by this package's trainer-side code,
    and the state the REFERENCE's TrainerState listener (pipelinerl/state.py:20-65) reached after tailing it
    (make_golden_trainer_state.py).  This package's listener must reach the same state from the same file."""
    import json
    from pipelinerl_b200 import streams
    from pipelinerl_b200.state import TrainerState
    from tests.helpers import GOLDEN
    rec = json.loads((GOLDEN / "trainer_state_case.json").read_text())
    assert rec["topic_file"] == "streams/weight_update_request/0/0/0.jsonl"
    f = tmp_path / rec["topic_file"]
    f.parent.mkdir(parents=True)
    f.write_text(rec["content"])
    streams.reset_streams_backend()
    streams.set_streams_backend("files")
    try:
        st = TrainerState(tmp_path)
        st.start_listening()
        assert st.wait_for_training_done(timeout=20)
        got = {"propagated_weight_version": st.propagated_weight_version, "samples_processed": st.samples_processed,
               "training_done": st.training_done}
        assert got == rec["reference_state"]
        st.stop()
    finally:
        streams.reset_streams_backend()
