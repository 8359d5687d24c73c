"""Golden for row a3's lag budget: the `max_lag` throttle of the reference's training actor loop
(pipelinerl/actor.py:509-534 arithmetic, :551-577 its use), EXECUTED on scripted (weight version, submit attempts) ticks.

The logic is inline in `ActorLoop._run` (a generator that needs hydra configs, queues and streams), so the statements are cut
out of the reference file with ast and exec'd / eval'd in a namespace holding a stub `self`:
  * `max_lag = ... if self.is_training else None` and the `if max_lag is not None:` block that computes
    `groups_per_update` and `can_submit_before_update`;
  * the `if self.trainer_state.propagated_weight_version > last_trainer_version:` block of the loop body;
  * the right-hand side of `blocked_by_lag = ...`.
One tick of the script = one iteration of the reference's outer loop: the version block once, then submit attempts until
blocked (the reference's inner `while True`).  The statements that run are the reference's.

    python tests/golden/make_golden_lag_budget.py      (authoring container only)

Recorded (tests/golden/lag_budget_cases.json): per case the configuration, the ticks, and after every tick the number of
groups submitted in it, the running total and `can_submit_before_update`."""
import ast
import json
import logging
import math
import types
from pathlib import Path

OUT = Path(__file__).resolve().parent
SRC = Path("/root/reference/pipelinerl/actor.py")

CASES = {
    # name: (max_lag, attempts, train_batch_size, gradient_accumulation_passes, weight_update_interval, ticks)
    # tick = (propagated_weight_version at the top of the iteration, problems available to submit in it)
    "base_like": (1024, 8, 1, 1024, 1, [(0, 500), (0, 10), (1, 500), (1, 5), (3, 1000), (3, 1), (4, 7), (9, 10 ** 4)]),
    "interval_rounds_up": (64, 8, 4, 8, 40, [(0, 100), (1, 100), (1, 100), (2, 3), (2, 100)]),
    "attempts_do_not_divide": (100, 7, 5, 3, 15, [(0, 50), (2, 50), (3, 1), (3, 50), (5, 50)]),
    "no_lag_allowance": (0, 4, 2, 2, 4, [(0, 9), (0, 9), (1, 9), (2, 0), (2, 9)]),
    "unthrottled": (None, 8, 1, 1024, 1, [(0, 300), (1, 300)]),
}


def find_statements():
    tree = ast.parse(SRC.read_text())
    run = next(n for n in ast.walk(tree) if isinstance(n, ast.FunctionDef) and n.name in ("run", "_run")
               and any(isinstance(m, ast.Assign) and ast.unparse(m.targets[0]) == "max_lag" for m in ast.walk(n)))
    max_lag_assign = next(m for m in ast.walk(run) if isinstance(m, ast.Assign) and ast.unparse(m.targets[0]) == "max_lag")
    arith_if = next(m for m in ast.walk(run) if isinstance(m, ast.If) and ast.unparse(m.test) == "max_lag is not None"
                    and any(isinstance(x, ast.Assign) and ast.unparse(x.targets[0]) == "can_submit_before_update" for x in ast.walk(m))
                    and m.orelse)
    version_if = next(m for m in ast.walk(run) if isinstance(m, ast.If)
                      and ast.unparse(m.test) == "self.trainer_state.propagated_weight_version > last_trainer_version")
    blocked = next(m for m in ast.walk(run) if isinstance(m, ast.Assign) and ast.unparse(m.targets[0]) == "blocked_by_lag")

    def code(nodes, mode="exec"):
        mod = ast.Module(body=list(nodes), type_ignores=[])
        ast.fix_missing_locations(mod)
        return compile(mod, str(SRC), mode)
    return (code([max_lag_assign, arith_if]), code([version_if]),
            compile(ast.fix_missing_locations(ast.Expression(body=blocked.value)), str(SRC), "eval"),
            [ast.get_source_segment(SRC.read_text(), n).splitlines()[0] for n in (max_lag_assign, arith_if, version_if, blocked)])


def run_case(setup_code, version_code, blocked_expr, spec):
    max_lag, attempts, tbs, gap, wui, ticks = spec
    fin = types.SimpleNamespace(max_lag=max_lag, train_batch_size=tbs, gradient_accumulation_passes=gap, weight_update_interval=wui)
    me = types.SimpleNamespace(cfg=types.SimpleNamespace(finetune=fin, attempts=attempts), is_training=True,
                               trainer_state=types.SimpleNamespace(propagated_weight_version=0))
    ns = {"self": me, "math": math, "logger": logging.getLogger("ref"), "last_trainer_version": 0, "submitted_groups": 0,
          "trainer_version_to_publish": None}
    exec(setup_code, ns)
    out = []
    for version, available in ticks:
        me.trainer_state.propagated_weight_version = version
        exec(version_code, ns)
        n = 0
        for _ in range(available):
            if eval(blocked_expr, ns):
                break
            ns["submitted_groups"] += 1
            n += 1
        cs = ns["can_submit_before_update"]
        out.append({"submitted_in_tick": n, "submitted_total": ns["submitted_groups"],
                    "can_submit_before_update": None if cs == math.inf else cs})
    return {"groups_per_update": ns["groups_per_update"], "after_tick": out}


def main():
    setup_code, version_code, blocked_expr, heads = find_statements()
    doc = {"reference_statements_executed": heads, "cases": {}}
    for name, spec in CASES.items():
        max_lag, attempts, tbs, gap, wui, ticks = spec
        doc["cases"][name] = {"config": {"max_lag": max_lag, "attempts": attempts, "train_batch_size": tbs,
                                         "gradient_accumulation_passes": gap, "weight_update_interval": wui},
                              "ticks": [list(t) for t in ticks], **run_case(setup_code, version_code, blocked_expr, spec)}
    (OUT / "lag_budget_cases.json").write_text(json.dumps(doc, indent=1))
    for k, v in doc["cases"].items():
        print(k, v["groups_per_update"], [(t["submitted_in_tick"], t["can_submit_before_update"]) for t in v["after_tick"]])


if __name__ == "__main__":
    main()
