"""Trainer step on random-init Qwen2.5-7B (hot path 2 end to end, one B200): rl_step (native body forward ->
fused tcgen05 head -> PG loss) -> backward (native body, fp32 gradient accumulation) for `--micro` packed
micro-batches of `--tokens` tokens, then the fused AdamW step and the refresh of the transposed weight copies.

Prints one JSON line: tokens/s, seconds per optimizer step at the stated batch, model FLOPs utilisation against
MEASURED_PEAKS.json (bf16 dense), and the CUDA-event breakdown (forward / backward / optimizer)."""
import argparse
import json
import sys
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from pipelinerl_b200 import _lib  # noqa: E402
from pipelinerl_b200.finetune.optim import FusedAdamW  # noqa: E402
from pipelinerl_b200.finetune.rl import RLConfig, rl_step  # noqa: E402
from pipelinerl_b200.finetune.types import PipelineBatchEncoding  # noqa: E402
from pipelinerl_b200.learner_model import NativeQwen2  # noqa: E402
from pipelinerl_b200.model import ModelConfig  # noqa: E402


def synthetic_batch(cfg, T, n_samples, dev, seed):
    """`n_samples` equal samples packed in one row of T tokens: first half of each is prompt (label -100)."""
    g = torch.Generator().manual_seed(seed)
    per = T // n_samples
    ids = torch.randint(0, 151643, (1, T), generator=g)
    pos = torch.cat([torch.arange(per)] * n_samples)[None]
    seg = torch.arange(n_samples).repeat_interleave(per)[None]
    labels = ids.clone()
    for s in range(n_samples):
        labels[0, s * per: s * per + per // 2] = -100
        labels[0, s * per] = -100
    n_lab = (labels[0].view(n_samples, per) >= 0).sum(1).float().repeat_interleave(per)[None]
    adv = (torch.randint(0, 2, (n_samples,), generator=g).float() * 2 - 1).repeat_interleave(per)[None]
    old = -(torch.rand(1, T, generator=g) * 0.2 + 11.8)   # random-init model: logprob ~ -log(V) = -11.93
    b = PipelineBatchEncoding(
        input_ids=ids, attention_mask=torch.ones(1, T, dtype=torch.long), labels=labels, position_ids=pos,
        segment_ids=seg, rewards=(adv + 1) / 2, advantages=adv, ref_logprobs=old.clone(), old_logprobs=old,
        group_tokens=torch.full((1, T), float(per)), num_labels=n_lab, overflow=torch.zeros(1, T),
        seq_boundaries=torch.arange(0, T + 1, per, dtype=torch.int32), model_version=0, is_packed=True)
    return b.to_device(dev)


def measure(model_name="7b", tokens=16384, samples_per_row=1, micro=2, steps=2, warmup=1, layers=0, keep_attn=-1,
            profile=False, dev=None, log=True):
    """Run the trainer step and return the result dict (also used by bench.py's `components.trainer_step`)."""
    dev = dev or torch.device("cuda:0")
    torch.cuda.set_device(dev)
    try:  # ~170 GB of the 192 GB are live at the peak: growable segments keep the caching allocator from fragmenting
        torch.cuda.memory._set_allocator_settings("expandable_segments:True")
    except Exception:  # noqa: BLE001
        pass
    cfg = ModelConfig.qwen2_5_7b() if model_name == "7b" else ModelConfig.tiny()
    if layers:
        from dataclasses import replace
        cfg = replace(cfg, num_layers=layers)

    def say(msg):
        if log:
            print(f"[train_bench] {msg}", file=sys.stderr, flush=True)
    t0 = time.time()
    model = NativeQwen2(cfg, dev)
    opt = FusedAdamW(model.named_parameters(), lr=1e-6, weight_decay=0.01, max_grad_norm=0.3, grad_dtype=torch.float32)
    model.bind(opt)
    if keep_attn >= 0:
        model.body.keep_attention_layers = keep_attn
    torch.cuda.synchronize()
    say(f"model + optimizer state resident: {torch.cuda.memory_allocated() / 1e9:.1f} GB ({time.time() - t0:.1f} s)")
    n_samples_step = micro * samples_per_row
    rcfg = RLConfig(batch_size=n_samples_step)   # reference defaults: ppo, kl_coef 0.1, temperature 1.0
    batches = [synthetic_batch(cfg, tokens, samples_per_row, dev, 100 + i) for i in range(micro)]
    ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731
    launches0 = None
    rec = []
    torch.cuda.reset_peak_memory_stats()
    for step in range(warmup + steps):
        if step == warmup:
            launches0 = _lib.launch_count()
        e = [ev() for _ in range(2 * micro + 3)]
        opt.zero_grad()
        e[0].record()
        losses = []
        for i, b in enumerate(batches):
            loss, stats = rl_step(model, b, step, 1000, rcfg)
            e[2 * i + 1].record()
            loss.backward()
            e[2 * i + 2].record()
            losses.append(loss.detach())
        gn = opt.step()
        model.after_optimizer_step()
        e[-1].record()
        torch.cuda.synchronize()
        fwd = sum(e[2 * i].elapsed_time(e[2 * i + 1]) for i in range(micro))
        bwd = sum(e[2 * i + 1].elapsed_time(e[2 * i + 2]) for i in range(micro))
        optm = e[2 * micro].elapsed_time(e[-1])
        total = e[0].elapsed_time(e[-1])
        rec.append((total, fwd, bwd, optm, float(sum(losses)), float(gn)))
        say(f"step {step}: {total:.1f} ms (fwd {fwd:.1f} bwd {bwd:.1f} opt {optm:.1f}) loss {rec[-1][4]:.5f} "
            f"grad_norm {rec[-1][5]:.4f} peak {torch.cuda.max_memory_allocated() / 1e9:.1f} GB")
    if profile:
        from torch.profiler import ProfilerActivity, profile as tprofile
        with tprofile(activities=[ProfilerActivity.CUDA]) as prof:
            opt.zero_grad()
            for b in batches:
                loss, _ = rl_step(model, b, 0, 1000, rcfg)
                loss.backward()
            opt.step()
            model.after_optimizer_step()
            torch.cuda.synchronize()
        print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=30, max_name_column_width=70),
              file=sys.stderr, flush=True)
    timed = rec[warmup:]
    ms = sum(r[0] for r in timed) / len(timed)
    total_tokens = micro * tokens
    c = cfg
    body_params = c.num_layers * (c.qkv_size * c.hidden_size + c.hidden_size * c.q_size + 3 * c.hidden_size * c.intermediate_size)
    head_params = c.vocab_size * c.hidden_size
    per_seg = tokens // samples_per_row
    attn_fwd = 4.0 * c.num_layers * c.num_q_heads * c.head_dim * (per_seg * (per_seg + 1) / 2) * samples_per_row
    # model FLOPs: 6 N per token + causal attention (forward 1x + backward 2x); recompute is NOT counted
    model_flops = micro * (6.0 * (body_params + head_params) * tokens + 3.0 * attn_fwd)
    peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text()) if (ROOT / "MEASURED_PEAKS.json").exists() else {}
    peak = peaks.get("bf16_tflops_sustained", 1459.7)
    out = {"bench": "trainer_step", "model": "Qwen2.5-7B" if model_name == "7b" else "tiny", "layers": c.num_layers,
           "tokens_per_micro_batch": tokens, "samples_per_micro_batch": samples_per_row, "micro_batches_per_step": micro,
           "samples_per_optimizer_step": n_samples_step, "ms_per_optimizer_step": round(ms, 2),
           "optimizer_steps_per_s": round(1000.0 / ms, 5), "trainer_tokens_per_s": round(total_tokens / ms * 1000.0, 1),
           "fwd_ms": round(sum(r[1] for r in timed) / len(timed), 2), "bwd_ms": round(sum(r[2] for r in timed) / len(timed), 2),
           "opt_ms": round(sum(r[3] for r in timed) / len(timed), 2),
           "model_TFLOPs": round(model_flops / ms / 1e9, 1),
           "mfu_vs_measured_sustained_peak": round(model_flops / ms / 1e9 / peak, 4), "peak_TFLOPs": peak,
           "libprl_launches_per_step": (_lib.launch_count() - launches0) // max(1, steps),
           "peak_memory_GB": round(torch.cuda.max_memory_allocated() / 1e9, 1),
           "loss": rec[-1][4], "grad_norm": rec[-1][5], "grad_accumulation": "fp32 in the optimizer arena",
           "attention": "torch SDPA (library)", "keep_attention_layers": model.body.keep_attention_layers,
           "gemm": "prl_gemm_ex (tcgen05 cta_group::2, MN-major dgrad/wgrad operands)"}
    del model, opt, batches
    import gc
    gc.collect()
    torch.cuda.empty_cache()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="7b")
    ap.add_argument("--tokens", type=int, default=16384)
    ap.add_argument("--samples-per-row", type=int, default=1)
    ap.add_argument("--micro", type=int, default=2, help="micro-batches per optimizer step")
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--layers", type=int, default=0, help="override the layer count (debug)")
    ap.add_argument("--keep-attn", type=int, default=-1, help="layers whose attention half is kept for backward (-1 = all)")
    ap.add_argument("--profile", action="store_true", help="after the timed steps, print the per-kernel CUDA time of one step")
    a = ap.parse_args()
    print(json.dumps(measure(a.model, a.tokens, a.samples_per_row, a.micro, a.steps, a.warmup, a.layers, a.keep_attn,
                             a.profile)))


if __name__ == "__main__":
    main()
