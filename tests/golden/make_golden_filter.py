"""Golden for `filter_zero_advantage_groups` (pipelinerl/preprocess.py:316-353): the reference FUNCTION is executed
(its source is cut out of the reference file with ast and exec'd, because importing pipelinerl.preprocess pulls in
packages that are not installed here) on synthetic grouped entries.

    python tests/golden/make_golden_filter.py      (authoring container only)
"""
import ast
import json
from pathlib import Path

import numpy as np

OUT = Path(__file__).resolve().parent
SRC = Path("/root/reference/pipelinerl/preprocess.py")


def reference_function():
    tree = ast.parse(SRC.read_text())
    fn = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "filter_zero_advantage_groups")
    ns: dict = {}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), str(SRC), "exec"), ns)
    return ns["filter_zero_advantage_groups"]


def main():
    ref = reference_function()
    rng = np.random.default_rng(11)
    cases = []
    for c in range(4):
        entries = []
        order = rng.permutation(np.repeat(np.arange(5), 3))       # 5 groups x 3 samples, interleaved arrival order
        for i, g in enumerate(order):
            zero = g in (1, 3) if c != 3 else True                 # case 3: every group is degenerate
            n = int(rng.integers(2, 6))
            adv = [0.0] * n if zero else [0.0] * (n - 1) + [float(rng.normal())]
            if c == 2 and g == 1:
                adv = [5e-7] * n                                   # below epsilon: still counts as zero
            entries.append({"group_id": f"g{g}", "uid": i, "advantages": adv})
        kept, dropped = ref([dict(e) for e in entries])
        cases.append({"entries": entries, "kept_uids": [e["uid"] for e in kept], "dropped": dropped})
        print("case", c, "kept", len(kept), "dropped", dropped)
    (OUT / "filter_zero_advantage_cases.json").write_text(json.dumps(cases))


if __name__ == "__main__":
    main()
