"""Golden for the trainer -> world messages on topic `weight_update_request` (pipelinerl/finetune_loop.py:138-171): the
reference's pydantic classes and its `TrainerMessage` union (cut out of the reference file with ast and exec'd) PARSE the
dumps of this package's message classes (pipelinerl_b200/weights.py), as `TrainerState.start_listening` does
(pipelinerl/state.py:35-47).  Recorded: each dump, and the reference class + fields it was parsed into.

    python tests/golden/make_golden_messages.py      (authoring container only)
"""
import ast
import json
import sys
import time
from pathlib import Path
from typing import Literal

from pydantic import BaseModel, TypeAdapter

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
OUT = Path(__file__).resolve().parent
SRC = Path("/root/reference/pipelinerl/finetune_loop.py")
NAMES = {"ParameterInfo", "WeightUpdateRequest", "WeightUpdateSuccess", "SamplesProcessed", "TrainingDone"}


def reference_messages():
    tree = ast.parse(SRC.read_text())
    body = [n for n in tree.body if (isinstance(n, ast.ClassDef) and n.name in NAMES) or
            (isinstance(n, ast.Assign) and any(isinstance(t, ast.Name) and t.id in ("TrainerMessage", "TRAINER_TOPIC")
                                                for t in n.targets))]
    ns = dict(BaseModel=BaseModel, Literal=Literal, time=time)
    exec(compile(ast.Module(body=body, type_ignores=[]), str(SRC), "exec"), ns)
    return ns


def main():
    from pipelinerl_b200 import weights as mine
    ref = reference_messages()
    adapter = TypeAdapter(ref["TrainerMessage"])
    msgs = [mine.WeightUpdateSuccess(version=12, timestamp=3.5), mine.SamplesProcessed(samples_processed=640, timestamp=4.5),
            mine.TrainingDone(timestamp=5.5),
            mine.WeightUpdateRequest(version=13, timestamp=6.5,
                                     parameters_info=[mine.ParameterInfo(name="w", shape=[2, 3], dtype="bfloat16")])]
    out = {"topic": ref["TRAINER_TOPIC"], "messages": []}
    for m in msgs:
        dump = m.model_dump()
        parsed = adapter.validate_python(dump)
        out["messages"].append({"ours": type(m).__name__, "dump": dump, "reference_class": type(parsed).__name__,
                                "reference_fields": parsed.model_dump()})
        print(type(m).__name__, "->", type(parsed).__name__)
    (OUT / "trainer_messages.json").write_text(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
