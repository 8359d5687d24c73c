"""Golden for row a1 against the reference's ACTUAL sampler engine family: vLLM (the reference pins 0.18.1,
pyproject.toml:22; the image ships 0.22.0 with FlashInfer sm_100) loaded with the SAME tiny bf16 weights the parity
tests use (tests.helpers.tiny_weights), on a GPU:

    gpurun -- python tests/golden/make_golden_vllm.py gpurun_out/vllm_golden       (GPU box only: vLLM needs libcuda)

Recorded per model kind (gqa2: 4q/2kv heads, gqa7: Qwen2.5-7B's 7:1 grouping) into vllm_tiny_<kind>.json:
  * greedy continuations (temperature 0) of fixed prompts with the logprob vLLM reports for every sampled token
    (`logprobs=0`: the field the reference's client reads, async_llm.py:173-207; at temperature 1 / top_p 1 / top_k -1,
    the reference's training parameters, processed and raw logprobs coincide);
  * teacher-forced `prompt_logprobs` of a 150-token sequence (the reference-logprob path, llm.py:606-648).
The committed JSON is then compared with this package's engine in tests/test_gpu_decode.py (bf16 engines differ by their
summation order: the test bounds the difference at 1.5x what was measured when the golden was made)."""
import json
import os
import sys
import tempfile
from pathlib import Path

os.environ.setdefault("VLLM_ENABLE_V1_MULTIPROCESSING", "0")
os.environ.setdefault("HF_HUB_OFFLINE", "1")
os.environ.setdefault("TRANSFORMERS_OFFLINE", "1")

import torch  # noqa: E402

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
from pipelinerl_b200.model import ArenaLayout  # noqa: E402
from tests.helpers import tiny_cfg, tiny_weights  # noqa: E402


def write_checkpoint(kind: str, d: str):
    from safetensors.torch import save_file
    cfg = tiny_cfg(kind)
    w = tiny_weights(cfg)
    sd = {}
    for hf_name, (fused, r0, rn) in ArenaLayout.build(cfg).hf_slices().items():
        sd[hf_name] = w[fused][r0:r0 + rn].to(torch.bfloat16).contiguous().clone()
    save_file(sd, os.path.join(d, "model.safetensors"), metadata={"format": "pt"})
    hc = {"architectures": ["Qwen2ForCausalLM"], "model_type": "qwen2", "vocab_size": cfg.vocab_size,
          "hidden_size": cfg.hidden_size, "intermediate_size": cfg.intermediate_size, "num_hidden_layers": cfg.num_layers,
          "num_attention_heads": cfg.num_q_heads, "num_key_value_heads": cfg.num_kv_heads, "head_dim": cfg.head_dim,
          "hidden_act": "silu", "max_position_embeddings": 4096, "rms_norm_eps": cfg.rms_eps, "rope_theta": cfg.rope_theta,
          "tie_word_embeddings": False, "torch_dtype": "bfloat16", "use_sliding_window": False, "bos_token_id": 1,
          "eos_token_id": cfg.vocab_size - 1, "attention_dropout": 0.0}
    with open(os.path.join(d, "config.json"), "w") as f:
        json.dump(hc, f)
    return cfg


def main():
    out_dir = Path(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/vllm_golden")
    out_dir.mkdir(parents=True, exist_ok=True)
    import vllm
    from vllm import LLM, SamplingParams
    for kind in ("gqa2", "gqa7"):
        d = tempfile.mkdtemp()
        cfg = write_checkpoint(kind, d)
        llm = LLM(model=d, skip_tokenizer_init=True, dtype="bfloat16", max_model_len=1024, max_num_seqs=8,
                  enable_prefix_caching=False, gpu_memory_utilization=0.25, seed=42, enforce_eager=True)
        g = torch.Generator().manual_seed(11)
        prompts = [torch.randint(0, cfg.vocab_size - 1, (n,), generator=g).tolist() for n in (5, 33, 64, 130, 200)]
        sp = SamplingParams(max_tokens=24, temperature=0.0, ignore_eos=True, logprobs=0, detokenize=False)
        outs = llm.generate([{"prompt_token_ids": p} for p in prompts], sp, use_tqdm=False)
        gens = []
        for o in outs:
            c = o.outputs[0]
            ids = list(c.token_ids)
            lps = [float(step[t].logprob) for step, t in zip(c.logprobs, ids)]
            gens.append({"ids": ids, "logprobs": lps})
        g2 = torch.Generator().manual_seed(7)
        seq = torch.randint(0, cfg.vocab_size, (150,), generator=g2).tolist()      # same sequence as make_golden_decode.py
        sp2 = SamplingParams(max_tokens=1, temperature=0.0, prompt_logprobs=0, detokenize=False)
        o2 = llm.generate([{"prompt_token_ids": seq}], sp2, use_tqdm=False)[0]
        plp = [float(d_[t].logprob) for d_, t in zip(o2.prompt_logprobs[1:], seq[1:])]
        rec = {"engine": f"vllm {vllm.__version__}", "kind": kind, "dtype": "bfloat16", "prompts": prompts, "greedy": gens,
               "teacher_forced": {"tokens": seq, "logprobs": plp}}
        (out_dir / f"vllm_tiny_{kind}.json").write_text(json.dumps(rec))
        print(kind, "greedy logprob mean", sum(sum(x["logprobs"]) for x in gens) / sum(len(x["ids"]) for x in gens),
              "teacher-forced mean", sum(plp) / len(plp), flush=True)
        del llm
        import gc
        gc.collect()
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
