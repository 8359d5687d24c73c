"""Cross-process, cross-GPU weight update (CUDA IPC + NVLink P2P).  Needs >= 2 GPUs: skipped on the
single-GPU test box, run with `gpurun --gpus 2`."""
import json
import os
import subprocess
import sys
from pathlib import Path

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_gpu_push_is_byte_exact_and_sampler_never_pauses():
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29583")
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29583", str(ROOT / "tools" / "push_bench.py"),
                          "--learners", "1", "--model", "tiny", "--updates", "3", "--context", "128", "--batch", "8"],
                         capture_output=True, text=True, env=env, timeout=600)
    assert res.returncode == 0, res.stderr[-3000:]
    out = json.loads([l for l in res.stdout.splitlines() if l.startswith("{")][-1])
    assert out["bytes_identical_on_all_ranks"]
    assert out["ours"]["stall_ms_max"] is not None and out["ours"]["stall_ms_max"] < 50


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_gpu_sharded_adamw_exchange_matches_oracle():
    """P2P reduce-scatter + AdamW shard + P2P all-gather == AdamW on the summed gradients (oracle), and every rank
    ends with bit-identical bf16 parameters."""
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29584")
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29584", str(ROOT / "tools" / "dp_adamw_bench.py"),
                          "--check"], capture_output=True, text=True, env=env, timeout=600)
    assert res.returncode == 0, res.stderr[-3000:]
    out = json.loads([l for l in res.stdout.splitlines() if l.startswith("{")][-1])
    assert out["ok"] and out["world"] == 2


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_gpu_tensor_parallel_engine_matches_oracle():
    """TP=2 engine (GEMM epilogue stores partials into the peer's reduction buffer, counters in peer memory, sampler
    partial exchange) == single-model oracle; both ranks sample identical tokens."""
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29585")
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29585", str(ROOT / "tools" / "tp_bench.py"),
                          "--check"], capture_output=True, text=True, env=env, timeout=900)
    assert res.returncode == 0, res.stderr[-3000:]
    out = json.loads([l for l in res.stdout.splitlines() if l.startswith("{")][-1])
    assert out["ok"]


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_gpu_native_learners_match_single_learner():
    """Two data-parallel native learners (fp32 micro-batch accumulation, bf16 P2P exchange fused with AdamW) end with
    bit-identical parameters on both ranks that equal the single-learner result on all micro-batches (to bf16
    exchange rounding)."""
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29586")
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29586", str(ROOT / "tools" / "train_bench.py"),
                          "--check"], capture_output=True, text=True, env=env, timeout=900)
    assert res.returncode == 0, res.stderr[-3000:]
    outs = [json.loads(l) for l in res.stdout.splitlines() if l.startswith("{")]
    assert outs and all(o["ok"] for o in outs), outs


@pytest.mark.parametrize("policy_loss", ["ppo", "gspo"])
@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_gpu_sequence_parallel_learner_matches_single_learner(policy_loss):
    """Two ranks share every packed row (rl_step(seq_parallel_group=...): slices of the same micro-batch, K / V
    all-gathered per layer, dK / dV reduce-scattered): summed loss, gradient norm and updated parameters equal one learner
    running the whole rows (reference: finetune_loop.py:507-517 ring attention over make_slices); with GSPO the per-segment
    sums are all-reduced inside the loss tail (rl/utils.py:194-206)."""
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29587")
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29587", str(ROOT / "tools" / "train_bench.py"),
                          "--check-sp", "--policy-loss", policy_loss], capture_output=True, text=True, env=env, timeout=900)
    assert res.returncode == 0, res.stderr[-3000:]
    outs = [json.loads(l) for l in res.stdout.splitlines() if l.startswith("{")]
    assert outs and all(o["ok"] for o in outs), outs
