"""Golden for row a5's writer: the loop of the reference preprocessor that DEALS packed micro-batches to the trainer
ranks (pipelinerl/preprocess.py:594-656: round-robin over lead trainers, per-trainer sample quota per optimizer step,
sentinel batches, seq_parallel slices), EXECUTED on scripted entries.

The loop is inline in `run_preprocessing_loop` (a 400-line function that needs hydra / redis / wandb), so its `while`
statement is cut out of the reference file with ast and exec'd in a namespace holding the local variables the
reference sets up at :463-483, the reference's own `collate_packed`, `create_sentinel_batch` and
`write_micro_batch_slices`, and a recording `data_writer`.  The statements that run are the reference's.

    python tests/golden/make_golden_dealer.py      (authoring container only)

Recorded (tests/golden/dealer_cases.json): per scenario the entries fed (in arrival chunks) and every write in order:
target rank, sentinel flag, and the full micro-batch columns."""
import ast
import copy
import json
import logging
import sys
import time
import types
from collections import deque
from pathlib import Path

import numpy as np

OUT = Path(__file__).resolve().parent
sys.path.insert(0, str(OUT))
import make_golden as mg  # noqa: E402

SRC = Path("/root/reference/pipelinerl/preprocess.py")

SCENARIOS = {
    # name: (num_trainers, seq_parallel, samples_per_lead_per_step, seq_length, n_groups, attempts, arrival chunk sizes)
    "one_trainer": (1, 1, 4, 60, 3, 4, [5, 7]),
    "two_trainers": (2, 1, 3, 40, 5, 4, [6, 1, 13]),
    "three_trainers_sentinels": (3, 1, 2, 26, 5, 4, [20]),
    "four_trainers_sp2": (4, 2, 2, 48, 4, 4, [9, 7]),
}


def reference_loop_node():
    tree = ast.parse(SRC.read_text())
    hits = [n for n in ast.walk(tree) if isinstance(n, ast.While)
            and ast.unparse(n.test).startswith("len(processed_entries_queue) > 0 and (not batch_done)")]
    assert len(hits) == 1, len(hits)
    return hits[0]


def dump_batch(b):
    return {k: (v.tolist() if hasattr(v, "tolist") else v) for k, v in mg.batch_to_np(b).items()}


def run_scenario(ref_rl, ref_data, ref_utils, ref_write_slices, loop_code, spec, seed):
    num_trainers, sp, spl, seq_length, n_groups, attempts, chunks = spec
    rng = np.random.default_rng(seed)
    samples = mg.make_samples(rng, n_groups=n_groups, attempts=attempts, vocab=97)
    entries = mg.preprocess_like_reference(ref_rl, ref_data, samples, ref_rl.RLConfig())
    entries = [e for e in entries if len(e["input_ids"]) <= seq_length]
    writes = []

    class Writer:
        def write(self, batch, partition):
            writes.append({"rank": int(partition), "batch": dump_batch(batch)})
    num_lead = num_trainers // sp
    ns = dict(
        cfg=types.SimpleNamespace(finetune=types.SimpleNamespace(seq_packing=True, seq_parallel=sp, seq_length=seq_length,
                                                                 train_batch_size=1),
                                  preprocess=types.SimpleNamespace(dataset_buffer_size=0)),
        logger=logging.getLogger("ref_preprocess"), time=time, tokenizer=mg._Tok(),
        collate_packed=ref_data.collate_packed, create_sentinel_batch=ref_utils.create_sentinel_batch,
        write_micro_batch_slices=ref_write_slices, data_writer=Writer(),
        # the local state of run_preprocessing_loop, as initialised at preprocess.py:463-483 with published_samples = 0
        num_trainers=num_trainers, trainer_id=0, published_samples=0,
        samples_per_trainer={i: 0 for i in range(0, num_trainers, sp)}, samples_per_lead_per_step=spl,
        train_batch_size=spl * num_lead, batch_boundary=spl * num_lead, target_samples_per_lead=spl,
        time_to_write=False, current_batch=[], current_length=0, batch_done=False, max_model_version=0,
        processed_entries_queue=deque(),
    )
    fed, pos, done_flags = [], 0, []
    arrivals = list(chunks)
    while pos < len(entries):
        n = arrivals.pop(0) if arrivals else len(entries) - pos
        chunk = entries[pos:pos + n]
        pos += n
        fed.append(copy.deepcopy(chunk))
        ns["processed_entries_queue"].extend(copy.deepcopy(chunk))
        ns["max_model_version"] = max(e["model_version"] for e in ns["processed_entries_queue"])   # :586
        # the reference re-enters the loop while data is queued; `batch_done` is reset per outer iteration (:592)
        while ns["processed_entries_queue"]:
            before = (len(ns["processed_entries_queue"]), len(writes))
            ns["batch_done"] = False
            exec(loop_code, ns)
            done_flags.append(bool(ns["batch_done"]))
            if (len(ns["processed_entries_queue"]), len(writes)) == before:
                break
    return {"spec": {"num_trainers": num_trainers, "seq_parallel": sp, "samples_per_lead_per_step": spl,
                     "seq_length": seq_length}, "eos_token_id": mg._Tok.eos_token_id, "arrivals": fed, "writes": writes,
            "batch_done_flags": done_flags, "published_samples": ns["published_samples"],
            "samples_per_trainer": {str(k): v for k, v in ns["samples_per_trainer"].items()},
            "left_in_current_batch": len(ns["current_batch"])}


def main():
    ref_rl, ref_data, ref_utils = mg._import_reference()
    tree = ast.parse(SRC.read_text())
    fn = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "write_micro_batch_slices")
    helper_ns = {"StreamWriter": object, "PipelineBatchEncoding": object}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), str(SRC), "exec"), helper_ns)
    loop_code = compile(ast.Module(body=[reference_loop_node()], type_ignores=[]), str(SRC), "exec")
    out = {}
    for i, (name, spec) in enumerate(SCENARIOS.items()):
        out[name] = run_scenario(ref_rl, ref_data, ref_utils, helper_ns["write_micro_batch_slices"], loop_code, spec, 100 + i)
        w = out[name]["writes"]
        print(name, "writes:", [(x["rank"], "S" if x["batch"]["sentinel"] else len(x["batch"]["seq_boundaries"]) - 1) for x in w],
              "done:", out[name]["batch_done_flags"], "left:", out[name]["left_in_current_batch"])
    (OUT / "dealer_cases.json").write_text(json.dumps(out))


if __name__ == "__main__":
    main()
