"""Qwen2-family model description and the flat parameter arena shared by learner and sampler.

One contiguous bf16 buffer holds every parameter in the FUSED layout the token-step kernels read
(qkv_proj = [q; k; v] rows, gate_up_proj = [gate; up] rows).  HF parameter names
(`model.layers.N.self_attn.q_proj.weight`, ...) map to row-slices of those fused tensors, which is
the name mapping vLLM's `load_weights` performs for the reference at pipelinerl/vllm1.py:122
(q/k/v_proj -> qkv_proj, gate/up_proj -> gate_up_proj).  Because learner and samplers use the SAME
arena layout, the in-flight weight update (hot path 3) is a plain byte copy of the arena — no
per-tensor loop, no name mapping at push time.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import torch


@dataclass(frozen=True)
class ModelConfig:
    vocab_size: int
    hidden_size: int
    intermediate_size: int
    num_layers: int
    num_q_heads: int
    num_kv_heads: int
    head_dim: int = 128
    rope_theta: float = 1_000_000.0
    rms_eps: float = 1e-6
    qkv_bias: bool = True
    fp32_head: bool = False  # keep a bf16 residual of the head (W = hi + lo): fp32-equivalent lm_head
    lm_head_rows: int | None = None  # vocabulary rows of THIS shard's lm_head (vocab-parallel head under TP)

    @property
    def q_size(self) -> int:
        return self.num_q_heads * self.head_dim

    @property
    def kv_size(self) -> int:
        return self.num_kv_heads * self.head_dim

    @property
    def qkv_size(self) -> int:
        return self.q_size + 2 * self.kv_size

    @property
    def head_rows(self) -> int:
        return self.lm_head_rows if self.lm_head_rows is not None else self.vocab_size

    def shard(self, tp: int) -> "ModelConfig":
        """Per-rank configuration under tensor parallelism: heads, MLP width and lm_head rows divided by tp
        (column-parallel qkv / gate_up / head, row-parallel o_proj / down_proj); embeddings and norms replicated."""
        from dataclasses import replace
        if tp == 1:
            return self
        for what, v in (("q heads", self.num_q_heads), ("kv heads", self.num_kv_heads),
                        ("intermediate", self.intermediate_size), ("vocab", self.vocab_size)):
            if v % tp:
                raise ValueError(f"{what} ({v}) not divisible by tp={tp}")
        return replace(self, num_q_heads=self.num_q_heads // tp, num_kv_heads=self.num_kv_heads // tp,
                       intermediate_size=self.intermediate_size // tp, lm_head_rows=self.vocab_size // tp)

    @staticmethod
    def qwen2_5_7b(**kw) -> "ModelConfig":
        return ModelConfig(vocab_size=152064, hidden_size=3584, intermediate_size=18944, num_layers=28,
                           num_q_heads=28, num_kv_heads=4, **kw)

    @staticmethod
    def qwen2_5_32b(**kw) -> "ModelConfig":
        return ModelConfig(vocab_size=152064, hidden_size=5120, intermediate_size=27648, num_layers=64,
                           num_q_heads=40, num_kv_heads=8, **kw)

    @staticmethod
    def tiny(**kw) -> "ModelConfig":
        """Plumbing / parity-test model (config 1 of BASELINE.json is not defined by the reference)."""
        base = dict(vocab_size=512, hidden_size=256, intermediate_size=640, num_layers=2, num_q_heads=4,
                    num_kv_heads=2)
        base.update(kw)
        return ModelConfig(**base)

    def num_params(self) -> int:
        return sum(n for _, n in ((name, _numel(shape)) for name, shape in fused_shapes(self)))


def _numel(shape) -> int:
    n = 1
    for s in shape:
        n *= s
    return n


def fused_shapes(cfg: ModelConfig) -> list[tuple[str, tuple[int, ...]]]:
    """Arena order.  Names are the fused (kernel-side) tensor names."""
    H, I = cfg.hidden_size, cfg.intermediate_size
    out: list[tuple[str, tuple[int, ...]]] = []
    for l in range(cfg.num_layers):
        p = f"layers.{l}."
        out.append((p + "input_layernorm.weight", (H,)))
        out.append((p + "qkv_proj.weight", (cfg.qkv_size, H)))
        if cfg.qkv_bias:
            out.append((p + "qkv_proj.bias", (cfg.qkv_size,)))
        out.append((p + "o_proj.weight", (H, cfg.q_size)))
        out.append((p + "post_attention_layernorm.weight", (H,)))
        out.append((p + "gate_up_proj.weight", (2 * I, H)))
        out.append((p + "down_proj.weight", (H, I)))
    out.append(("embed_tokens.weight", (cfg.vocab_size, H)))
    out.append(("norm.weight", (H,)))
    out.append(("lm_head.weight", (cfg.head_rows, H)))
    if cfg.fp32_head:
        out.append(("lm_head.weight_lo", (cfg.head_rows, H)))
    return out


@dataclass
class ArenaLayout:
    cfg: ModelConfig
    offsets: dict[str, int] = field(default_factory=dict)
    shapes: dict[str, tuple[int, ...]] = field(default_factory=dict)
    total: int = 0

    @staticmethod
    def build(cfg: ModelConfig, align: int = 64) -> "ArenaLayout":
        lay = ArenaLayout(cfg)
        at = 0
        for name, shape in fused_shapes(cfg):
            lay.offsets[name] = at
            lay.shapes[name] = shape
            at = (at + _numel(shape) + align - 1) // align * align  # 128-byte aligned tensors (TMA needs 16)
        lay.total = at
        return lay

    # HF name -> (fused name, row start, row count)
    def hf_slices(self) -> dict[str, tuple[str, int, int]]:
        c = self.cfg
        m: dict[str, tuple[str, int, int]] = {}
        for l in range(c.num_layers):
            hp, fp = f"model.layers.{l}.", f"layers.{l}."
            m[hp + "input_layernorm.weight"] = (fp + "input_layernorm.weight", 0, c.hidden_size)
            m[hp + "post_attention_layernorm.weight"] = (fp + "post_attention_layernorm.weight", 0, c.hidden_size)
            for kind in ("weight", "bias") if c.qkv_bias else ("weight",):
                m[hp + f"self_attn.q_proj.{kind}"] = (fp + f"qkv_proj.{kind}", 0, c.q_size)
                m[hp + f"self_attn.k_proj.{kind}"] = (fp + f"qkv_proj.{kind}", c.q_size, c.kv_size)
                m[hp + f"self_attn.v_proj.{kind}"] = (fp + f"qkv_proj.{kind}", c.q_size + c.kv_size, c.kv_size)
            m[hp + "self_attn.o_proj.weight"] = (fp + "o_proj.weight", 0, c.hidden_size)
            m[hp + "mlp.gate_proj.weight"] = (fp + "gate_up_proj.weight", 0, c.intermediate_size)
            m[hp + "mlp.up_proj.weight"] = (fp + "gate_up_proj.weight", c.intermediate_size, c.intermediate_size)
            m[hp + "mlp.down_proj.weight"] = (fp + "down_proj.weight", 0, c.hidden_size)
        m["model.embed_tokens.weight"] = ("embed_tokens.weight", 0, c.vocab_size)
        m["model.norm.weight"] = ("norm.weight", 0, c.hidden_size)
        m["lm_head.weight"] = ("lm_head.weight", 0, c.vocab_size)
        return m


class ParamArena:
    """Flat bf16 parameter buffer + named views."""

    def __init__(self, cfg: ModelConfig, device, dtype=torch.bfloat16, data: torch.Tensor | None = None):
        self.cfg = cfg
        self.layout = ArenaLayout.build(cfg)
        if data is None:
            data = torch.zeros(self.layout.total, dtype=dtype, device=device)
        assert data.numel() == self.layout.total and data.dtype == dtype
        self.data = data
        self.version = 0

    def view(self, name: str) -> torch.Tensor:
        off, shape = self.layout.offsets[name], self.layout.shapes[name]
        return self.data[off:off + _numel(shape)].view(shape)

    def ptr(self, name: str) -> int:
        return self.data.data_ptr() + self.layout.offsets[name] * self.data.element_size()

    def names(self) -> list[str]:
        return list(self.layout.offsets)

    def nbytes(self) -> int:
        return self.data.numel() * self.data.element_size()

    def init_random(self, seed: int = 42, std: float = 0.02) -> "ParamArena":
        """normal(0, 0.02) weights, unit norm gains, zero biases — HF's Qwen2 initialisation; seed = conf/base.yaml:7."""
        g = torch.Generator(device=self.data.device).manual_seed(seed)
        for name in self.names():
            v = self.view(name)
            if name.endswith("layernorm.weight") or name == "norm.weight":
                v.fill_(1.0)
            elif name.endswith(".bias") or name.endswith("_lo"):
                v.zero_()
            else:
                # chunked to bound the fp32 temporary for [152064, 3584] tensors
                flat = v.view(-1)
                step = 1 << 26
                for s in range(0, flat.numel(), step):
                    n = min(step, flat.numel() - s)
                    flat[s:s + n] = (torch.randn(n, generator=g, device=self.data.device, dtype=torch.float32)
                                     * std).to(self.data.dtype)
        return self

    def load_hf_state_dict(self, sd: dict[str, torch.Tensor]) -> None:
        slices = self.layout.hf_slices()
        seen = set()
        for hf_name, t in sd.items():
            if hf_name not in slices:
                raise KeyError(f"unexpected parameter {hf_name}")
            fused, r0, rn = slices[hf_name]
            dst = self.view(fused)[r0:r0 + rn]
            if self.cfg.fp32_head and hf_name == "lm_head.weight":
                hi = t.to(torch.bfloat16)
                dst.copy_(hi)
                self.view("lm_head.weight_lo").copy_((t.float() - hi.float()).to(torch.bfloat16))
            else:
                dst.copy_(t.to(self.data.dtype))
            seen.add(hf_name)
        missing = set(slices) - seen
        if missing == {"lm_head.weight"}:  # tied embeddings
            self.view("lm_head.weight").copy_(self.view("embed_tokens.weight"))
        elif missing:
            raise KeyError(f"missing parameters: {sorted(missing)[:4]} ...")

    def hf_state_dict(self) -> dict[str, torch.Tensor]:
        return {hf: self.view(fused)[r0:r0 + rn] for hf, (fused, r0, rn) in self.layout.hf_slices().items()}


def shard_fused_weights(cfg: ModelConfig, full: dict[str, torch.Tensor], rank: int, tp: int) -> dict[str, torch.Tensor]:
    """Slice full fused tensors (names of fused_shapes(cfg)) into rank `rank`'s tensor-parallel shard."""
    loc = cfg.shard(tp)
    d, out = cfg.head_dim, {}
    ql, kl, I, Il = loc.q_size, loc.kv_size, cfg.intermediate_size, loc.intermediate_size
    for name, t in full.items():
        if name.endswith("qkv_proj.weight") or name.endswith("qkv_proj.bias"):
            q, k, v = t[:cfg.q_size], t[cfg.q_size:cfg.q_size + cfg.kv_size], t[cfg.q_size + cfg.kv_size:]
            out[name] = torch.cat([q[rank * ql:(rank + 1) * ql], k[rank * kl:(rank + 1) * kl], v[rank * kl:(rank + 1) * kl]])
        elif name.endswith("o_proj.weight"):
            out[name] = t[:, rank * ql:(rank + 1) * ql].contiguous()
        elif name.endswith("gate_up_proj.weight"):
            out[name] = torch.cat([t[rank * Il:(rank + 1) * Il], t[I + rank * Il:I + (rank + 1) * Il]])
        elif name.endswith("down_proj.weight"):
            out[name] = t[:, rank * Il:(rank + 1) * Il].contiguous()
        elif name.startswith("lm_head.weight"):
            out[name] = t[rank * loc.head_rows:(rank + 1) * loc.head_rows]
        else:
            out[name] = t
    return out
