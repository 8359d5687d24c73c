"""Golden fixture for the WHOLE hot path 2: the reference's `rl_step` (pipelinerl/finetune/rl/__init__.py:136-450)
driving HF transformers' `Qwen2ForCausalLM` (the model class the reference trains, finetune/checkpoints.py:151-222),
fp32 on CPU, on one packed micro-batch; loss, the 32 statistics, the per-token new logprobs and the gradient of
EVERY parameter.

    python tests/golden/make_golden_learner.py      (authoring container: needs /root/reference + transformers)

Packed rows: on GPUs the reference reaches block-diagonal causal attention through flash-attn's varlen path, which HF
selects from `position_ids` that restart per sample (conf/finetune/base.yaml attn_implementation=flash_attention_2).
flash-attn does not run on CPU, so the generator wraps the HF model and hands it the equivalent 4-D additive mask
(same-sample AND causal); everything else is the unmodified reference / HF code path.

Weights are NOT stored (tests regenerate them with tests.helpers.tiny_weights).  Gradients are stored as a summary
per tensor: L2 norm + 257 elements at fixed strided positions (enough to catch layout / sign / scale errors, small
enough to commit).
"""
from __future__ import annotations

import json
import sys
import types
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(Path(__file__).resolve().parent))
from make_golden import _Tok, _import_reference, batch_to_np, make_samples, preprocess_like_reference  # noqa: E402
from pipelinerl_b200.model import ArenaLayout  # noqa: E402
from tests.helpers import tiny_cfg, tiny_weights  # noqa: E402

OUT = Path(__file__).resolve().parent
N_SAMPLE = 257


def sample_idx(numel: int) -> np.ndarray:
    return np.unique(np.linspace(0, numel - 1, num=min(N_SAMPLE, numel)).astype(np.int64))


class PackedHF(torch.nn.Module):
    """HF model + the block-diagonal causal mask flash-attn varlen implements for packed position_ids."""

    def __init__(self, hf):
        super().__init__()
        self.hf = hf

    def forward(self, input_ids=None, attention_mask=None, labels=None, position_ids=None, **kw):
        B, T = input_ids.shape
        if position_ids is None:
            position_ids = torch.arange(T)[None].expand(B, T)
        seg = (position_ids == 0).cumsum(-1)
        allowed = (seg[:, :, None] == seg[:, None, :]) & (torch.arange(T)[:, None] >= torch.arange(T)[None, :])[None]
        if attention_mask is not None:
            allowed = allowed & attention_mask[:, None, :].bool()
        mask4d = torch.zeros(B, 1, T, T).masked_fill(~allowed[:, None], torch.finfo(torch.float32).min)
        out = self.hf(input_ids=input_ids, attention_mask=mask4d, position_ids=position_ids)
        return types.SimpleNamespace(logits=out.logits)


def main():
    from transformers import Qwen2Config, Qwen2ForCausalLM
    ref_rl, ref_data, ref_utils = _import_reference()
    for kind, cfgd in (("gqa2", dict(policy_loss="ppo", kl_coef=0.1, final_kl_coef=0.02, entropy_bonus=0.01,
                                     final_entropy_bonus=0.001, epsilon_low=0.2, epsilon_high=0.3, batch_size=16,
                                     clamp_log_ratio_ref_new_value=1.5)),
                       ("gqa7", dict(policy_loss="gspo", kl_coef=0.0, final_kl_coef=0.0, epsilon_low=0.05,
                                     epsilon_high=0.05, batch_size=8))):
        cfg = tiny_cfg(kind)
        w = tiny_weights(cfg)
        hc = Qwen2Config(vocab_size=cfg.vocab_size, hidden_size=cfg.hidden_size, intermediate_size=cfg.intermediate_size,
                         num_hidden_layers=cfg.num_layers, num_attention_heads=cfg.num_q_heads,
                         num_key_value_heads=cfg.num_kv_heads, rope_theta=cfg.rope_theta, rms_norm_eps=cfg.rms_eps,
                         tie_word_embeddings=False, max_position_embeddings=4096, head_dim=cfg.head_dim,
                         attn_implementation="eager")
        hf = Qwen2ForCausalLM(hc).train().float()
        slices = ArenaLayout.build(cfg).hf_slices()
        sd = {hf_name: w[fused][r0:r0 + rn].clone() for hf_name, (fused, r0, rn) in slices.items()}
        missing, unexpected = hf.load_state_dict(sd, strict=False)
        assert not unexpected and all("rotary" in m or "inv_freq" in m for m in missing), (missing, unexpected)
        model = PackedHF(hf)

        rng = np.random.default_rng(500 + len(kind))
        torch.manual_seed(500)
        rcfg = ref_rl.RLConfig(**cfgd)
        samples = make_samples(rng, n_groups=2, attempts=4, vocab=cfg.vocab_size, max_prompt=14, max_gen=30)
        entries = preprocess_like_reference(ref_rl, ref_data, samples, rcfg)
        batch = ref_data.collate_packed(entries, _Tok(), seq_parallel=1)
        T = batch.input_ids.shape[1]
        with torch.no_grad():   # old / ref logprobs near the model's own, so both sides of the clip are exercised
            lg = model(input_ids=batch.input_ids, attention_mask=batch.attention_mask, position_ids=batch.position_ids).logits
            lp = torch.log_softmax(lg[0, :-1] / rcfg.temperature, -1).gather(1, batch.input_ids[0, 1:, None])[:, 0]
            batch.old_logprobs[0, 1:] = lp + 0.05 * torch.randn(T - 1)
            batch.ref_logprobs[0, 1:] = lp + 0.3 * torch.randn(T - 1)
        cur, mx = 3, 10
        loss, stats = ref_rl.rl_step(model, batch, cur, mx, rcfg)
        loss.backward()
        arrs = batch_to_np(batch)
        arrs["loss"] = np.float64(loss.item())
        arrs["new_logprobs"] = lp.numpy()
        grads = {}
        for hf_name, (fused, r0, rn) in slices.items():
            grads.setdefault(fused, torch.zeros_like(w[fused]))
            g = dict(hf.named_parameters())[hf_name].grad
            grads[fused][r0:r0 + rn] = g
        for fused, g in grads.items():
            flat = g.reshape(-1).double()
            key = fused.replace(".", "__")
            arrs["gnorm__" + key] = np.float64(flat.norm().item())
            arrs["gsamp__" + key] = flat[torch.from_numpy(sample_idx(flat.numel()))].numpy()
        np.savez_compressed(OUT / f"learner_step_{kind}.npz", **arrs)
        meta = {"config": rcfg.model_dump(), "current_step": cur, "max_step": mx,
                "stats": {k: float(v) for k, v in stats.items()}, "model": kind, "T": int(T)}
        (OUT / f"learner_step_{kind}.json").write_text(json.dumps(meta, indent=1, sort_keys=True))
        tot = float(torch.sqrt(sum((g.double() ** 2).sum() for g in grads.values())))
        print(kind, "T", T, "loss", loss.item(), "total grad norm", tot)


if __name__ == "__main__":
    main()
