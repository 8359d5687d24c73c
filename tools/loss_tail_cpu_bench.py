#!/usr/bin/env python
"""CPU timing of the finetune loss tail (SURVEY §8d "Reference CPU path timing", item 2): `rl_step` on a stub model
that returns fixed random fp32 logits [1, T, 152064] (T = 2048 to fit RAM; the tail is linear in T).

    python tools/loss_tail_cpu_bench.py [--tokens 2048] [--reference /root/reference]

Times (a) the REFERENCE's own rl_step (pipelinerl/finetune/rl/__init__.py:136-450) when the reference tree is
importable (authoring container), (b) the oracle port (oracle/pg_oracle.py) — forward+backward, all host threads.
One JSON line: microseconds per token for each, next to the GPU kernels' numbers quoted from bench.py components.
"""
import argparse
import json
import os
import sys
import time
import types
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def make_batch_arrays(T, V, n_seg, seed=0):
    rng = np.random.default_rng(seed)
    per = T // n_seg
    ids = rng.integers(8, V, size=(1, T))
    labels = ids.copy()
    pos = np.concatenate([np.arange(per)] * n_seg)[None]
    seg = np.repeat(np.arange(n_seg), per)[None]
    for s in range(n_seg):
        labels[0, s * per: s * per + per // 2] = -100
    f = lambda a: a.astype(np.float32)  # noqa: E731
    adv = np.repeat(rng.normal(size=n_seg), per)[None]
    return dict(input_ids=ids, attention_mask=np.ones((1, T), np.int64), labels=labels, position_ids=pos, segment_ids=seg,
                rewards=f(adv * 0 + 1), advantages=f(adv), ref_logprobs=f(-rng.random((1, T)) * 3),
                old_logprobs=f(-rng.random((1, T)) * 3), group_tokens=f(np.full((1, T), per)),
                num_labels=f(np.full((1, T), per // 2)), overflow=f(np.zeros((1, T))),
                seq_boundaries=np.arange(0, T + 1, per, dtype=np.int32))


def time_it(fn, reps):
    fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    return (time.perf_counter() - t0) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tokens", type=int, default=2048)
    ap.add_argument("--vocab", type=int, default=152064)
    ap.add_argument("--reference", default="/root/reference")
    ap.add_argument("--reps", type=int, default=2)
    a = ap.parse_args()
    T, V = a.tokens, a.vocab
    torch.set_num_threads(os.cpu_count())
    arrs = make_batch_arrays(T, V, 8)
    logits0 = torch.randn(1, T, V) * 1.5
    cfgd = dict(policy_loss="ppo", kl_coef=0.1, final_kl_coef=0.1, entropy_bonus=0.0, batch_size=8)
    out = {"bench": "loss_tail_cpu", "tokens": T, "vocab": V, "threads": os.cpu_count(),
           "config": cfgd, "what": "rl_step forward + backward on fixed fp32 logits (stub model)"}

    from oracle import pg_oracle
    ocfg = pg_oracle.OracleRLConfig.from_dict(cfgd)
    cols = {k: torch.from_numpy(v[0]) for k, v in arrs.items() if v.ndim == 2}

    def oracle_step():
        lg = logits0[0].clone().requires_grad_(True)
        loss, _, _, _ = pg_oracle.rl_step_oracle(lg, cols, ocfg, 0, 10)
        loss.backward()
    s = time_it(oracle_step, a.reps)
    out["oracle_port"] = {"s_per_step": round(s, 3), "us_per_token": round(s / T * 1e6, 1)}

    if Path(a.reference, "pipelinerl").is_dir():
        import transformers  # noqa: F401
        sys.path.insert(0, a.reference)
        om = types.ModuleType("omegaconf")
        om.DictConfig, om.ListConfig, om.OmegaConf = dict, list, object
        sys.modules.setdefault("omegaconf", om)
        acc = types.ModuleType("accelerate")
        acc.Accelerator = object
        sys.modules.setdefault("accelerate", acc)
        from pipelinerl.finetune import rl as ref_rl
        from pipelinerl.finetune.types import PipelineBatchEncoding as RefBatch
        kw = {k: torch.from_numpy(v) for k, v in arrs.items()}
        batch = RefBatch(**kw, model_version=0, sentinel=False, padding=0, is_packed=True)
        rcfg = ref_rl.RLConfig(**cfgd)

        class Stub(torch.nn.Module):
            def __init__(self):
                super().__init__()
                self.logits = torch.nn.Parameter(logits0.clone())

            def forward(self, **kw):
                return types.SimpleNamespace(logits=self.logits)
        model = Stub()

        def ref_step():
            model.logits.grad = None
            loss, _ = ref_rl.rl_step(model, batch, 0, 10, rcfg)
            loss.backward()
        s = time_it(ref_step, a.reps)
        out["reference"] = {"s_per_step": round(s, 3), "us_per_token": round(s / T * 1e6, 1),
                            "kind": "reference rl_step executed from " + a.reference}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
