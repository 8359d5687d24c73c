"""Control-plane interop recording: the REFERENCE's `TrainerState` listener (pipelinerl/state.py:20-65, class source cut out
with ast) tails a `weight_update_request` topic that THIS package's trainer-side code wrote (pipelinerl_b200 streams +
message classes, in the order `run_training` / `WeightUpdateManager` emit them), with the reference's own
`TrainerMessage` union (finetune_loop.py:138-171) and this package's stream reader standing in for `pipelinerl.streams`
(whose file format is pinned separately by make_golden_streams.py).

    python tests/golden/make_golden_trainer_state.py      (authoring container only)

Recorded: the topic file and the state the reference listener ended up in; tests replay the same file through
pipelinerl_b200.state.TrainerState.
"""
import ast
import json
import logging
import sys
import tempfile
import threading
import time
from pathlib import Path
from typing import Literal

from pydantic import BaseModel, TypeAdapter

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(Path(__file__).resolve().parent))
from make_golden_messages import reference_messages  # noqa: E402

OUT = Path(__file__).resolve().parent


def reference_trainer_state(ns_msgs, streams_mod):
    tree = ast.parse(Path("/root/reference/pipelinerl/state.py").read_text())
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "TrainerState")
    ns = dict(Path=Path, threading=threading, time=time, TypeAdapter=TypeAdapter, logger=logging.getLogger("ref_state"),
              SingleStreamSpec=streams_mod.SingleStreamSpec, read_stream=streams_mod.read_stream,
              TRAINER_TOPIC=ns_msgs["TRAINER_TOPIC"], TrainerMessage=ns_msgs["TrainerMessage"],
              WeightUpdateSuccess=ns_msgs["WeightUpdateSuccess"], SamplesProcessed=ns_msgs["SamplesProcessed"],
              TrainingDone=ns_msgs["TrainingDone"], BaseModel=BaseModel, Literal=Literal)
    exec(compile(ast.Module(body=[cls], type_ignores=[]), "state.py", "exec"), ns)
    return ns["TrainerState"]


def main():
    from pipelinerl_b200 import streams, weights
    streams.reset_streams_backend()
    streams.set_streams_backend("files")
    msgs = [weights.SamplesProcessed(samples_processed=8, timestamp=1.0), weights.WeightUpdateSuccess(version=8, timestamp=2.0),
            weights.SamplesProcessed(samples_processed=16, timestamp=3.0), weights.WeightUpdateSuccess(version=16, timestamp=4.0),
            weights.SamplesProcessed(samples_processed=24, timestamp=5.0), weights.TrainingDone(timestamp=6.0)]
    with tempfile.TemporaryDirectory() as tmp:
        exp = Path(tmp)
        with streams.write_to_streams(streams.SingleStreamSpec(exp_path=exp, topic=weights.TRAINER_TOPIC)) as w:
            for m in msgs:
                w.write(m)
        RefState = reference_trainer_state(reference_messages(), streams)
        st = RefState(exp)
        st.start_listening()
        assert st.wait_for_training_done(timeout=20)
        topic_file = next(exp.rglob("*.jsonl"))
        rec = {"topic_file": str(topic_file.relative_to(exp)), "content": topic_file.read_text(),
               "reference_state": {"propagated_weight_version": st.propagated_weight_version,
                                   "samples_processed": st.samples_processed, "training_done": st.training_done}}
    (OUT / "trainer_state_case.json").write_text(json.dumps(rec, indent=1))
    print(rec["reference_state"], rec["topic_file"])


if __name__ == "__main__":
    main()
