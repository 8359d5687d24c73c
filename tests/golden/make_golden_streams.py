"""Golden for the stream API (§8b "Stream API"): the reference's file backend (pipelinerl/streams.py:238-423) EXECUTED:
`set_streams_backend("files")`, `write_to_streams(SingleStreamSpec | StreamRangeSpec)`, round-robin and explicit
partition writes, BaseModel / tensor payloads.

    python tests/golden/make_golden_streams.py      (authoring container only)

orjson and redis are not installed here.  redis is never touched by the file backend; orjson is replaced by a stand-in
built on `json` (compact separators, numpy arrays as lists), so what this fixture pins is the PROTOCOL — directory layout
`<exp>/streams/<topic>/<instance>/<partition>/0.jsonl`, one JSON document per line, partitioning, what a pydantic model or
a tensor turns into — not orjson's byte-level float formatting (JSON-equal lines are what readers rely on).
"""
import json
import sys
import tempfile
import types
from pathlib import Path

import numpy as np
import torch
from pydantic import BaseModel

OUT = Path(__file__).resolve().parent


def stub_modules():
    oj = types.ModuleType("orjson")
    oj.OPT_SERIALIZE_NUMPY = 1

    def dumps(obj, option=0):
        def default(o):
            if isinstance(o, np.ndarray):
                return o.tolist()
            if isinstance(o, (np.integer,)):
                return int(o)
            if isinstance(o, (np.floating,)):
                return float(o)
            raise TypeError(type(o))
        return json.dumps(obj, separators=(",", ":"), ensure_ascii=False, default=default).encode("utf-8")
    oj.dumps, oj.loads = dumps, json.loads
    sys.modules["orjson"] = oj
    rd = types.ModuleType("redis")
    rd.exceptions = types.ModuleType("redis.exceptions")
    sys.modules["redis"], sys.modules["redis.exceptions"] = rd, rd.exceptions


class Success(BaseModel):
    kind: str = "weight_update_success"
    version: int
    timestamp: float


class WithTensor(BaseModel):
    model_config = {"arbitrary_types_allowed": True}
    name: str
    values: torch.Tensor


def writes():
    """(topic, spec kwargs, [(payload factory, partition)]) — the same list drives the replay in the test."""
    g1 = [{"text": "a b", "input_ids": [1, 2, 3], "logprobs": [-0.5, -0.25], "reward": 1.0, "group_id": "actor0_0",
           "metadata": {"model_version": 7, "rollout_index": 0, "step_index": 0}, "finished": True}]
    g2 = [{"text": "é ü", "input_ids": [4], "logprobs": [-1.5], "reward": 0.0, "group_id": "actor0_1",
           "metadata": {"model_version": 7, "rollout_index": 1, "step_index": 0}, "finished": False}]
    return [
        ("actor", dict(instance=0, partition=0), [(g1, None), (g2, None)]),
        ("training_data", dict(instance=0, partition_range=(0, 2)),
         [({"i": i, "arr": (np.arange(3) + i)}, None) for i in range(5)] + [({"i": 99, "arr": np.zeros(2)}, 1)]),
        ("weight_update_request", dict(instance=0, partition=0),
         [(Success(version=3, timestamp=12.5), None), (WithTensor(name="t", values=torch.arange(4, dtype=torch.float32)), None)]),
    ]


def main():
    stub_modules()
    sys.path.insert(0, "/root/reference")
    import pipelinerl.streams as ref
    ref.set_streams_backend("files")
    with tempfile.TemporaryDirectory() as tmp:
        exp = Path(tmp)
        for topic, kw, items in writes():
            spec = ref.StreamRangeSpec(exp_path=exp, topic=topic, **kw) if "partition_range" in kw \
                else ref.SingleStreamSpec(exp_path=exp, topic=topic, **kw)
            with ref.write_to_streams(spec) as w:
                for payload, part in items:
                    w.write(payload, partition=part) if part is not None else w.write(payload)
        tree = {str(p.relative_to(exp)): p.read_text(encoding="utf-8") for p in sorted(exp.rglob("*.jsonl"))}
        # what the reference READER yields for the actor topic (first two records of the tailing generator)
        with ref.read_stream(ref.SingleStreamSpec(exp_path=exp, topic="actor")) as r:
            it = r.read()
            read_back = [next(it), next(it)]
    (OUT / "streams_tree.json").write_text(json.dumps({"tree": tree, "actor_read_back": read_back}, indent=1, ensure_ascii=False))
    for k, v in tree.items():
        print(k, len(v.splitlines()), "lines")


if __name__ == "__main__":
    main()
