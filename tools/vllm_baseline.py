#!/usr/bin/env python
"""Performance baseline for hot path 1 on the same box: vLLM (the engine the reference serves rollouts with;
the image ships 0.22, the reference pins 0.18.1) on random-init Qwen2.5-7B, 64 sequences x 8192-token synthetic
prompts, temperature 1, logprobs on, prefix caching off, 1024-token chunked prefill (conf/base.yaml:59-73).
Decode tokens/s = 64 * (N2 - N1) / (T(N2 new tokens) - T(N1 new tokens)), which cancels prefill.
NOT part of the product path: a library baseline for SURVEY §8d / BASELINE.md B4.  One JSON line."""
import json
import os
import sys
import tempfile
import time

os.environ.setdefault("VLLM_ENABLE_V1_MULTIPROCESSING", "0")
os.environ.setdefault("HF_HUB_OFFLINE", "1")
os.environ.setdefault("TRANSFORMERS_OFFLINE", "1")


def main():
    import numpy as np
    batch = int(os.environ.get("BATCH", 64))
    ctx = int(os.environ.get("CTX", 8192))
    n1, n2 = 8, 72
    d = tempfile.mkdtemp()
    cfg = {"architectures": ["Qwen2ForCausalLM"], "model_type": "qwen2", "vocab_size": 152064, "hidden_size": 3584,
           "intermediate_size": 18944, "num_hidden_layers": 28, "num_attention_heads": 28, "num_key_value_heads": 4,
           "hidden_act": "silu", "max_position_embeddings": 32768, "rms_norm_eps": 1e-6, "rope_theta": 1000000.0,
           "tie_word_embeddings": False, "torch_dtype": "bfloat16", "use_sliding_window": False, "bos_token_id": 151643,
           "eos_token_id": 151643, "attention_dropout": 0.0, "initializer_range": 0.02}
    with open(os.path.join(d, "config.json"), "w") as f:
        json.dump(cfg, f)
    from vllm import LLM, SamplingParams
    import vllm
    llm = LLM(model=d, load_format="dummy", skip_tokenizer_init=True, dtype="bfloat16", max_model_len=ctx + 256,
              max_num_seqs=batch, max_num_batched_tokens=1024, enable_prefix_caching=False, enable_chunked_prefill=True,
              gpu_memory_utilization=0.9, seed=42)
    rng = np.random.default_rng(0)
    prompts = [{"prompt_token_ids": rng.integers(8, 151643, size=ctx).tolist()} for _ in range(batch)]

    def run(n_new):
        sp = SamplingParams(max_tokens=n_new, temperature=1.0, ignore_eos=True, logprobs=0, detokenize=False)
        t0 = time.perf_counter()
        outs = llm.generate(prompts, sp, use_tqdm=False)
        dt = time.perf_counter() - t0
        assert all(len(o.outputs[0].token_ids) == n_new for o in outs)
        return dt
    run(2)  # warm-up (graph capture, allocator)
    t1 = run(n1)
    t2 = run(n2)
    dec = batch * (n2 - n1) / (t2 - t1)
    print(json.dumps({"bench": "vllm_baseline", "vllm": vllm.__version__, "model": "Qwen2.5-7B dummy weights",
                      "batch": batch, "context": ctx, "decode_tokens_per_s": round(dec, 1),
                      "ms_per_step": round((t2 - t1) / (n2 - n1) * 1e3, 3), "t_n1_s": round(t1, 3), "t_n2_s": round(t2, 3),
                      "prefill_plus_%d_tokens_s" % n1: round(t1, 3), "flags": "max_num_seqs=64, max_num_batched_tokens=1024, "
                      "chunked prefill, prefix caching off, logprobs on, bf16 (no fp32 lm_head plugin)"}), flush=True)


if __name__ == "__main__":
    main()
