"""Golden fixtures for the token step: HF transformers' Qwen2ForCausalLM (what the reference's trainer
runs, finetune/checkpoints.py:151-222) in fp32 on bf16-valued weights.

    python tests/golden/make_golden_decode.py      (authoring container; needs transformers only)

Weights are NOT stored: tests regenerate them with tests.helpers.tiny_weights (CPU torch RNG, seed 42).
Stored: teacher-forced logprobs of a fixed 150-token sequence and the full logits of its last 4 positions.
"""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
from pipelinerl_b200.model import ArenaLayout  # noqa: E402
from tests.helpers import tiny_cfg, tiny_weights  # noqa: E402


def main():
    from transformers import Qwen2Config, Qwen2ForCausalLM
    for kind in ("gqa2", "gqa7"):
        cfg = tiny_cfg(kind)
        w = tiny_weights(cfg)
        hc = Qwen2Config(vocab_size=cfg.vocab_size, hidden_size=cfg.hidden_size, intermediate_size=cfg.intermediate_size,
                         num_hidden_layers=cfg.num_layers, num_attention_heads=cfg.num_q_heads,
                         num_key_value_heads=cfg.num_kv_heads, rope_theta=cfg.rope_theta, rms_norm_eps=cfg.rms_eps,
                         tie_word_embeddings=False, max_position_embeddings=4096, head_dim=cfg.head_dim,
                         attn_implementation="eager")
        model = Qwen2ForCausalLM(hc).eval().float()
        sd = {}
        for hf_name, (fused, r0, rn) in ArenaLayout.build(cfg).hf_slices().items():
            sd[hf_name] = w[fused][r0:r0 + rn].clone()
        missing, unexpected = model.load_state_dict(sd, strict=False)
        assert not unexpected and all("rotary" in m or "inv_freq" in m for m in missing), (missing, unexpected)
        g = torch.Generator().manual_seed(7)
        tokens = torch.randint(0, cfg.vocab_size, (150,), generator=g)
        with torch.no_grad():
            logits = model(input_ids=tokens[None]).logits[0].float()
        for temp in (1.0, 0.7):
            lp = torch.log_softmax(logits[:-1] / temp, -1).gather(1, tokens[1:, None])[:, 0]
            np.savez_compressed(Path(__file__).parent / f"qwen2_tiny_{kind}_T{temp}.npz", tokens=tokens.numpy(),
                                logprobs=lp.numpy(), last_logits=logits[-4:].numpy(), temperature=np.float32(temp))
        print(kind, "logprob mean", float(lp.mean()), "logit std", float(logits.std()))


if __name__ == "__main__":
    main()
