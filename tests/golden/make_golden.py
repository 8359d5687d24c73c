"""Generate golden fixtures by EXECUTING the reference (ServiceNow/PipelineRL) on fixed seeds.

Run in the authoring container only (needs /root/reference, which does not exist
on the GPU box):

    python tests/golden/make_golden.py

The reference has no tests / golden vectors of its own, so
these fixtures are what pins oracle/ (and through it the CUDA kernels) to the
reference's behaviour.  Outputs: tests/golden/*.npz + *.json (small, committed).

Reference functions executed:
  pipelinerl.finetune.rl.rl_step / populate_rl_data / prepare_rl_fields
  pipelinerl.finetune.data.collate_packed / collate / preprocess_fn
  pipelinerl.finetune.utils.create_sentinel_batch
  torch.optim.AdamW with the groups of pipelinerl/finetune/optim.py:8-22
"""
from __future__ import annotations

import copy
import json
import zlib
import sys
import types
from pathlib import Path

import numpy as np
import torch

REF = "/root/reference"
OUT = Path(__file__).resolve().parent


def _import_reference():
    import transformers  # noqa: F401  (must be imported before the stubs)
    sys.path.insert(0, REF)
    om = types.ModuleType("omegaconf")
    om.DictConfig = dict
    om.ListConfig = list
    om.OmegaConf = object
    sys.modules.setdefault("omegaconf", om)
    acc = types.ModuleType("accelerate")
    acc.Accelerator = object
    sys.modules.setdefault("accelerate", acc)
    from pipelinerl.finetune import rl as ref_rl
    from pipelinerl.finetune import data as ref_data
    from pipelinerl.finetune import utils as ref_utils
    return ref_rl, ref_data, ref_utils


class _Tok:
    eos_token_id = 7
    padding_side = "right"


class _StubModel(torch.nn.Module):
    """rl_step only needs model(**inputs).logits (rl/__init__.py:190-207)."""

    def __init__(self, logits):
        super().__init__()
        self.logits = torch.nn.Parameter(logits)

    def forward(self, **kw):
        return types.SimpleNamespace(logits=self.logits)


def make_samples(rng: np.random.Generator, n_groups: int, attempts: int, vocab: int, max_prompt=9, max_gen=14,
                 finish_mix=True):
    """Synthetic TrainingText-like dicts as the actor publishes them (rollouts.py:13-57)."""
    samples = []
    for g in range(n_groups):
        prompt = rng.integers(8, vocab, size=int(rng.integers(3, max_prompt))).tolist()
        for a in range(attempts):
            n_gen = int(rng.integers(2, max_gen))
            gen = rng.integers(8, vocab, size=n_gen).tolist()
            fin = bool(rng.random() < 0.7) if finish_mix else True
            if fin:
                gen[-1] = _Tok.eos_token_id
            s = {
                "input_ids": prompt + gen,
                "labels": [-100] * len(prompt) + gen,
                "logprobs": (-rng.random(n_gen) * 3).astype(np.float64).tolist(),
                "ref_logprobs": (-rng.random(n_gen) * 3).astype(np.float64).tolist(),
                "reward": float(rng.random() < 0.5) + (0.25 if a % 3 == 0 else 0.0),
                "group_id": f"g{g}",
                "rollout_index": a,
                "step_index": 0,
                "finished": fin,
                "metadata": {"model_version": 3 + (a % 2)},
            }
            if a % 4 == 1:
                s["finish_reason"] = "length" if not fin else "stop"
            samples.append(s)
    return samples


def preprocess_like_reference(ref_rl, ref_data, samples, rl_config):
    """preprocess_dataset (preprocess.py:145-189) minus tokenizer/LLM I/O."""
    entries = []
    for s in samples:
        e = dict(s)
        enc = ref_data.preprocess_fn(e, _Tok(), seq_length=10_000, is_rl=True)
        for k in ("group_id", "rollout_index", "step_index", "finished"):
            enc[k] = s[k]
        if "finish_reason" in s:
            enc["finish_reason"] = s["finish_reason"]
        enc["model_version"] = s["metadata"]["model_version"]
        entries.append(enc)
    return ref_rl.populate_rl_data(entries, _Tok.eos_token_id, rl_config)


def batch_to_np(b):
    out = {}
    for k in ("input_ids", "attention_mask", "labels", "position_ids", "segment_ids", "rewards", "advantages",
              "ref_logprobs", "old_logprobs", "group_tokens", "num_labels", "overflow", "seq_boundaries"):
        v = getattr(b, k)
        if v is not None:
            out[k] = v.numpy()
    out["model_version"] = np.int64(b.model_version)
    out["sentinel"] = np.bool_(b.sentinel)
    out["padding"] = np.int64(b.padding)
    out["is_packed"] = np.bool_(b.is_packed)
    return out


def gen_rl_step_cases(ref_rl, ref_data, ref_utils):
    cases = {
        "ppo_default": dict(policy_loss="ppo", kl_coef=0.0, final_kl_coef=0.0, epsilon_low=0.02, epsilon_high=0.02,
                            batch_size=8, divide_advantage_by_std=False, clamp_log_ratio_ref_new_value=5),
        "ppo_kl_entropy": dict(policy_loss="ppo", kl_coef=0.1, final_kl_coef=0.02, entropy_bonus=0.01,
                               final_entropy_bonus=0.001, epsilon_low=0.2, epsilon_high=0.3, batch_size=16,
                               clamp_log_ratio_ref_new_value=1.5),
        "reinforce_default": dict(policy_loss="reinforce", kl_coef=0.0, final_kl_coef=0.0, epsilon_low=0.02,
                                  epsilon_high=0.02, batch_size=8, divide_advantage_by_std=False,
                                  clamp_log_ratio_ref_new_value=5),
        "reinforce_groupnorm_overlong": dict(policy_loss="reinforce", kl_coef=0.05, final_kl_coef=0.05,
                                             group_normalization=True, overlong_filtering=True, batch_size=8,
                                             epsilon_high=0.5, relu_log_p_weights=True),
        "ppo_rewards_not_adv": dict(policy_loss="ppo", use_advantages=False, kl_coef=0.0, final_kl_coef=0.0,
                                    batch_size=4),
        "gspo": dict(policy_loss="gspo", kl_coef=0.0, final_kl_coef=0.0, epsilon_low=0.05, epsilon_high=0.05,
                     batch_size=8),
        "gspo_groupnorm": dict(policy_loss="gspo", kl_coef=0.0, final_kl_coef=0.0, epsilon_low=0.003,
                               epsilon_high=0.004, group_normalization=True, batch_size=8),
    }
    vocab = 61
    for ci, (name, cfgd) in enumerate(cases.items()):
        rng = np.random.default_rng(100 + ci)
        torch.manual_seed(100 + ci)
        cfg = ref_rl.RLConfig(**cfgd)
        samples = make_samples(rng, n_groups=2, attempts=4, vocab=vocab)
        entries = preprocess_like_reference(ref_rl, ref_data, samples, cfg)
        batch = ref_data.collate_packed(entries, _Tok(), seq_parallel=1)
        T = batch.input_ids.shape[1]
        logits0 = torch.randn(1, T, vocab) * 1.5
        # make old/ref logprobs correlate with the new ones so that ratios land inside the clip range for
        # some tokens and outside for others (both branches of the clipped objective are exercised)
        with torch.no_grad():
            lp = torch.log_softmax(logits0[0, :-1] / cfg.temperature, dim=-1)
            new_lp = lp.gather(1, batch.input_ids[0, 1:, None])[:, 0]
            near = torch.rand(T - 1) < 0.6
            batch.old_logprobs[0, 1:] = torch.where(near, new_lp + 0.015 * torch.randn(T - 1), batch.old_logprobs[0, 1:])
            batch.ref_logprobs[0, 1:] = new_lp + 0.7 * torch.randn(T - 1)
        model = _StubModel(logits0.clone())
        cur, mx = 3, 10
        loss, stats = ref_rl.rl_step(model, batch, cur, mx, cfg)
        loss.backward()
        arrs = batch_to_np(batch)
        arrs["logits"] = logits0.numpy()
        arrs["grad_logits"] = model.logits.grad.numpy()
        arrs["loss"] = np.float64(loss.item())
        np.savez_compressed(OUT / f"rl_step_{name}.npz", **arrs)
        meta = {"config": cfg.model_dump(), "current_step": cur, "max_step": mx,
                "stats": {k: float(v) for k, v in stats.items()}}
        (OUT / f"rl_step_{name}.json").write_text(json.dumps(meta, indent=1, sort_keys=True))
        print(name, "loss", loss.item(), "T", T)

    # sentinel batch (finetune/utils.py:17-57): loss must be exactly zero, stats only input_size
    cfg = ref_rl.RLConfig(policy_loss="ppo", kl_coef=0.0, final_kl_coef=0.0, batch_size=8)
    sb = ref_utils.create_sentinel_batch("cpu", tokenizer=_Tok(), model_version=5)
    model = _StubModel(torch.randn(1, 8, vocab))
    loss, stats = ref_rl.rl_step(model, sb, 0, 10, cfg)
    arrs = batch_to_np(sb)
    arrs["logits"] = model.logits.detach().numpy()
    arrs["loss"] = np.float64(loss.item())
    np.savez_compressed(OUT / "rl_step_sentinel.npz", **arrs)
    (OUT / "rl_step_sentinel.json").write_text(json.dumps(
        {"config": cfg.model_dump(), "current_step": 0, "max_step": 10,
         "stats": {k: float(v) for k, v in stats.items()}}, indent=1, sort_keys=True))

    # unpacked batch through collate() (data.py:163-212)
    cfg = ref_rl.RLConfig(policy_loss="ppo", kl_coef=0.1, final_kl_coef=0.1, batch_size=3)
    rng = np.random.default_rng(77)
    torch.manual_seed(77)
    samples = make_samples(rng, n_groups=1, attempts=3, vocab=vocab)
    entries = preprocess_like_reference(ref_rl, ref_data, samples, cfg)
    keep = ["input_ids", "labels", "attention_mask", *ref_rl.RL_DATA_COLUMNS, "model_version"]
    entries2 = [{k: e[k] for k in keep} for e in entries]
    batch = ref_data.collate(entries2, _Tok())
    B, L = batch.input_ids.shape
    model = _StubModel(torch.randn(B, L, vocab))
    logits0 = model.logits.detach().clone()
    loss, stats = ref_rl.rl_step(model, batch, 0, 10, cfg)
    loss.backward()
    arrs = batch_to_np(batch)
    arrs["logits"] = logits0.numpy()
    arrs["grad_logits"] = model.logits.grad.numpy()
    arrs["loss"] = np.float64(loss.item())
    np.savez_compressed(OUT / "rl_step_unpacked.npz", **arrs)
    (OUT / "rl_step_unpacked.json").write_text(json.dumps(
        {"config": cfg.model_dump(), "current_step": 0, "max_step": 10,
         "stats": {k: float(v) for k, v in stats.items()}}, indent=1, sort_keys=True))


def gen_preprocess_cases(ref_rl, ref_data, ref_utils):
    """populate_rl_data + collate_packed on raw samples; the integer/bit-exact layout contract."""
    out = {}
    for name, cfgd, sp in [
        ("loo_std", dict(divide_advantage_by_std=True), 1),
        ("loo_nostd", dict(divide_advantage_by_std=False), 1),
        ("loo_std_sp4", dict(divide_advantage_by_std=True), 4),
    ]:
        rng = np.random.default_rng(zlib.crc32(name.encode()) % 1000)
        cfg = ref_rl.RLConfig(**cfgd)
        samples = make_samples(rng, n_groups=3, attempts=4, vocab=97)
        # one single-member group: loo_mean = own reward, std = NaN -> nan_to_num (rl/__init__.py:513-519)
        samples.append(dict(copy.deepcopy(samples[0]), group_id="solo", rollout_index=0))
        raw = copy.deepcopy(samples)
        entries = preprocess_like_reference(ref_rl, ref_data, samples, cfg)
        batch = ref_data.collate_packed(copy.deepcopy(entries), _Tok(), seq_parallel=sp)
        cols = ["input_ids", "labels", "attention_mask", "rewards", "advantages", "old_logprobs", "ref_logprobs",
                "overflow", "group_tokens", "num_labels", "model_version"]
        out[name] = {
            "config": cfg.model_dump(), "seq_parallel": sp, "eos_token_id": _Tok.eos_token_id, "raw_samples": raw,
            "entries": [{k: e[k] for k in cols} for e in entries],
            "batch": {k: (v.tolist() if hasattr(v, "tolist") else v) for k, v in batch_to_np(batch).items()},
        }
    (OUT / "preprocess_cases.json").write_text(json.dumps(out))
    print("preprocess cases:", list(out))


def gen_preprocess_large_cases(ref_rl, ref_data, ref_utils):
    """Big groups (64 attempts, BASELINE config 5), real-valued rewards, two steps per rollout (multi-turn: several
    samples share a rollout_index) and unfinished samples without finish_reason: exercises pandas' Kahan sum / Welford std
    beyond what 4-member groups can, and the "eos in input_ids" overflow rule."""
    out = {}
    for name, cfgd, sp, group_sizes in [
        ("group64_std", dict(divide_advantage_by_std=True), 1, [64, 33]),
        ("group64_nostd_sp4", dict(divide_advantage_by_std=False), 4, [64, 7, 1]),
    ]:
        rng = np.random.default_rng(zlib.crc32(name.encode()) % 100000)
        cfg = ref_rl.RLConfig(**cfgd)
        samples = []
        for g, size in enumerate(group_sizes):
            prompt = rng.integers(8, 97, size=int(rng.integers(3, 9))).tolist()
            for a in range(size):
                for step in range(2 if (a % 5 == 0) else 1):         # some rollouts have two turns
                    n_gen = int(rng.integers(2, 12))
                    gen = rng.integers(8, 97, size=n_gen).tolist()
                    fin = bool(rng.random() < 0.6)
                    if fin or rng.random() < 0.3:
                        gen[-1] = _Tok.eos_token_id
                    s = {"input_ids": prompt + gen, "labels": [-100] * len(prompt) + gen,
                         "logprobs": (-rng.random(n_gen) * 3).tolist(), "ref_logprobs": (-rng.random(n_gen) * 3).tolist(),
                         "reward": float(rng.random() * 3.7 - 1.1) * (1e3 if g == 1 else 1.0),
                         "group_id": f"g{g}", "rollout_index": a, "step_index": step, "finished": fin,
                         "metadata": {"model_version": 5 + (a % 3)}}
                    if a % 4 == 1:
                        s["finish_reason"] = "length" if not fin else "stop"
                    samples.append(s)
        raw = copy.deepcopy(samples)
        entries = preprocess_like_reference(ref_rl, ref_data, samples, cfg)
        batch = ref_data.collate_packed(copy.deepcopy(entries), _Tok(), seq_parallel=sp)
        cols = ["rewards", "advantages", "overflow", "group_tokens", "num_labels"]
        out[name] = {"config": cfg.model_dump(), "seq_parallel": sp, "eos_token_id": _Tok.eos_token_id, "raw_samples": raw,
                     "entry_scalars": [{k: e[k][0] for k in cols} for e in entries],
                     "batch": {k: (v.tolist() if hasattr(v, "tolist") else v) for k, v in batch_to_np(batch).items()}}
    import gzip
    with gzip.open(OUT / "preprocess_cases_large.json.gz", "wt", compresslevel=9) as f:
        f.write(json.dumps(out))
    print("large preprocess cases:", {k: len(v["raw_samples"]) for k, v in out.items()})


def gen_adamw_case():
    """torch.optim.AdamW with the decay groups of finetune/optim.py:8-22, 3 steps with clipping (finetune_loop.py:739)."""
    torch.manual_seed(5)
    shapes = {"embed.weight": (37, 16), "layers.0.q_proj.weight": (16, 16), "layers.0.q_proj.bias": (16,),
              "layers.0.input_layernorm.weight": (16,), "norm.weight": (16,), "lm_head.weight": (37, 16),
              "odd.bias": (3,), "odd.weight": (5, 7)}
    params = {n: torch.nn.Parameter(torch.randn(s) * 0.1) for n, s in shapes.items()}
    no_decay = ["bias", "LayerNorm.weight"]
    wd_p = [p for n, p in params.items() if not any(nd in n for nd in no_decay)]
    nd_p = [p for n, p in params.items() if any(nd in n for nd in no_decay)]
    opt = torch.optim.AdamW([{"params": wd_p, "weight_decay": 0.01}, {"params": nd_p, "weight_decay": 0.0}], lr=1e-3)
    arrs = {f"p0/{n}": p.detach().clone().numpy() for n, p in params.items()}
    norms = []
    for step in range(3):
        for n, p in params.items():
            p.grad = torch.randn_like(p) * (0.5 if step != 1 else 0.01)
            arrs[f"g{step}/{n}"] = p.grad.clone().numpy()
        norms.append(float(torch.nn.utils.clip_grad_norm_(list(params.values()), 0.3)))
        opt.step()
        for n, p in params.items():
            arrs[f"p{step + 1}/{n}"] = p.detach().clone().numpy()
    arrs["grad_norms"] = np.array(norms)
    np.savez_compressed(OUT / "adamw_case.npz", **arrs)
    (OUT / "adamw_case.json").write_text(json.dumps(
        {"names": list(shapes), "shapes": {k: list(v) for k, v in shapes.items()}, "lr": 1e-3, "weight_decay": 0.01,
         "betas": [0.9, 0.999], "eps": 1e-8, "max_grad_norm": 0.3, "no_decay": no_decay}, indent=1))
    print("adamw norms", norms)


if __name__ == "__main__":
    ref_rl, ref_data, ref_utils = _import_reference()
    gen_rl_step_cases(ref_rl, ref_data, ref_utils)
    gen_preprocess_cases(ref_rl, ref_data, ref_utils)
    gen_preprocess_large_cases(ref_rl, ref_data, ref_utils)
    gen_adamw_case()
