"""Learner-side Qwen2 module whose parameters ARE views of one arena in the fused layout (model.py).

The transformer BODY here is plain torch (F.linear / SDPA): the sm_100a trainer-body kernels
(varlen attention fwd/bwd, fused blocks) are the next row of SURVEY §8(f) and are not built yet, so
this module is plumbing that lets the full actor -> preprocess -> rl_step -> FusedAdamW -> weight
push loop run end to end with the hot-path kernels that do exist (logprob tail, PG loss, AdamW, push).
Because parameter order and alignment equal `fused_shapes`, FusedAdamW's bf16 shadow arena has exactly
the sampler's arena layout and can be pushed as raw bytes.
"""
from __future__ import annotations

import math
import types

import torch
import torch.nn.functional as F

from .model import ArenaLayout, ModelConfig, fused_shapes


class TorchQwen2(torch.nn.Module):
    def __init__(self, cfg: ModelConfig, device, dtype=torch.float32, init: dict[str, torch.Tensor] | None = None,
                 seed: int = 42):
        super().__init__()
        self.cfg = cfg
        g = torch.Generator().manual_seed(seed)
        self.names = []
        for name, shape in fused_shapes(cfg):
            if init is not None:
                t = init[name].to(dtype)
            elif name.endswith("layernorm.weight") or name == "norm.weight":
                t = torch.ones(shape, dtype=dtype)
            elif name.endswith(".bias") or name.endswith("_lo"):
                t = torch.zeros(shape, dtype=dtype)
            else:
                t = (torch.randn(shape, generator=g) * 0.02).to(dtype)
            key = name.replace(".", "__")
            self.register_parameter(key, torch.nn.Parameter(t.to(device)))
            self.names.append(name)
        d = cfg.head_dim
        self.register_buffer("inv_freq", (1.0 / (cfg.rope_theta ** (torch.arange(0, d, 2, dtype=torch.int64).float()
                                                                     / d))).to(device), persistent=False)
        self.layout = ArenaLayout.build(cfg)

    def p(self, name: str) -> torch.Tensor:
        return getattr(self, name.replace(".", "__"))

    def named_parameters(self, *a, **k):  # fused names, arena order
        for name in self.names:
            yield name, self.p(name)

    def _rope(self, x, pos):
        ang = pos.float()[:, None] * self.inv_freq[None, :]
        cs, sn = torch.cos(ang)[:, None, :].to(x.dtype), torch.sin(ang)[:, None, :].to(x.dtype)
        x1, x2 = x[..., :64], x[..., 64:]
        return torch.cat([x1 * cs - x2 * sn, x2 * cs + x1 * sn], dim=-1)

    def forward_logprobs(self, batch, temperature: float):
        """rl_step's fast path: (new_logprobs [B, L-1], entropy [B, L-1]) through the fused tcgen05 head — the
        [L, V] logits are never materialised (finetune/fused_head.py)."""
        from .finetune.fused_head import fused_head_logprobs
        hidden = self.hidden_states(batch.input_ids, batch.position_ids if batch.is_packed else None)
        lps, ents = [], []
        for b in range(hidden.shape[0]):
            lp, ent = fused_head_logprobs(hidden[b, :-1], self.p("lm_head.weight"), batch.input_ids[b, 1:], temperature)
            lps.append(lp)
            ents.append(ent)
        return torch.stack(lps), torch.stack(ents)

    def forward(self, input_ids, attention_mask=None, labels=None, position_ids=None, **kw):
        """Packed rows [1, T] with position_ids restarting per sample (block-diagonal causal attention), or
        padded [B, L] batches."""
        x = self.hidden_states(input_ids, position_ids)
        return types.SimpleNamespace(logits=F.linear(x.float(), self.p("lm_head.weight").float()))

    def hidden_states(self, input_ids, position_ids=None):
        """Final-norm hidden states [B, T, H]."""
        c = self.cfg
        B, T = input_ids.shape
        if position_ids is None:
            position_ids = torch.arange(T, device=input_ids.device)[None].expand(B, T)
        outs = []
        for b in range(B):
            pos = position_ids[b]
            starts = (pos == 0).cumsum(0)
            allowed = (starts[:, None] == starts[None, :]) & (torch.arange(T, device=pos.device)[:, None] >=
                                                             torch.arange(T, device=pos.device)[None, :])
            h = self.p("embed_tokens.weight")[input_ids[b]]
            for l in range(c.num_layers):
                q_ = f"layers.{l}."
                x = self._norm(h, self.p(q_ + "input_layernorm.weight"))
                qkv = F.linear(x, self.p(q_ + "qkv_proj.weight"), self.p(q_ + "qkv_proj.bias") if c.qkv_bias else None)
                q = qkv[:, :c.q_size].view(T, c.num_q_heads, c.head_dim)
                k = qkv[:, c.q_size:c.q_size + c.kv_size].view(T, c.num_kv_heads, c.head_dim)
                v = qkv[:, c.q_size + c.kv_size:].view(T, c.num_kv_heads, c.head_dim)
                q, k = self._rope(q, pos), self._rope(k, pos)
                R = c.num_q_heads // c.num_kv_heads
                k, v = k.repeat_interleave(R, dim=1), v.repeat_interleave(R, dim=1)
                o = F.scaled_dot_product_attention(q.transpose(0, 1), k.transpose(0, 1), v.transpose(0, 1),
                                                   attn_mask=allowed[None], scale=1.0 / math.sqrt(c.head_dim))
                h = h + F.linear(o.transpose(0, 1).reshape(T, c.q_size), self.p(q_ + "o_proj.weight"))
                x = self._norm(h, self.p(q_ + "post_attention_layernorm.weight"))
                gu = F.linear(x, self.p(q_ + "gate_up_proj.weight"))
                h = h + F.linear(F.silu(gu[:, :c.intermediate_size]) * gu[:, c.intermediate_size:],
                                 self.p(q_ + "down_proj.weight"))
            outs.append(self._norm(h, self.p("norm.weight")))
        return torch.stack(outs)

    def _norm(self, h, g):
        hf = h.float()
        return (hf * torch.rsqrt((hf * hf).mean(-1, keepdim=True) + self.cfg.rms_eps)).to(h.dtype) * g
