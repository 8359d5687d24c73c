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
