"""Binary micro-batch records (pipelinerl_b200/records.py) and the GPU-resident preprocess (csrc/preprocess_pack.cu).

CPU: the record round-trips, travels through the file streams as bytes, and the host `populate_rl_data` equals the
reference (pandas) on 64-member groups with real-valued rewards (tests/golden/preprocess_cases_large.json.gz, produced by
executing the reference).  GPU: `GpuPreprocessor.pack(record)` equals the reference's populate_rl_data + collate_packed
BIT FOR BIT on every column of every golden case."""
import copy
import gzip
import json
from pathlib import Path

import numpy as np
import pytest
import torch

from pipelinerl_b200.finetune.data import collate_packed, preprocess_fn
from pipelinerl_b200.finetune.rl import RLConfig, populate_rl_data
from pipelinerl_b200.records import GpuPreprocessor, RecordView, encode_micro_batch_record

GOLDEN = Path(__file__).parent / "golden"


class Tok:
    eos_token_id = 7
    padding_side = "right"


def _cases():
    small = json.loads((GOLDEN / "preprocess_cases.json").read_text())
    with gzip.open(GOLDEN / "preprocess_cases_large.json.gz", "rt") as f:
        large = json.loads(f.read())
    return {**small, **large}


CASES = _cases()


def _chunk(case):
    chunk = []
    for s in copy.deepcopy(case["raw_samples"]):
        s["model_version"] = s["metadata"]["model_version"]
        chunk.append(s)
    return chunk


@pytest.mark.parametrize("name", ["group64_std", "group64_nostd_sp4"])
def test_host_populate_rl_data_matches_pandas_on_large_groups(name):
    case = CASES[name]
    cfg = RLConfig(**case["config"])
    entries = []
    for s in _chunk(case):
        enc = preprocess_fn(s, Tok(), seq_length=10_000, is_rl=True)
        for k in ("group_id", "rollout_index", "step_index", "finished", "model_version"):
            enc[k] = s[k]
        if "finish_reason" in s:
            enc["finish_reason"] = s["finish_reason"]
        entries.append(enc)
    entries = populate_rl_data(entries, Tok.eos_token_id, cfg)
    for e, want in zip(entries, case["entry_scalars"]):
        for k, w in want.items():
            assert e[k][0] == w, (k, e[k][0], w)          # float64, exact
    batch = collate_packed(entries, Tok(), seq_parallel=case["seq_parallel"])
    for k, w in case["batch"].items():
        g = getattr(batch, k)
        if isinstance(g, torch.Tensor):
            assert torch.equal(g, torch.tensor(w, dtype=g.dtype)), k
        else:
            assert g == w, k


def test_record_layout_and_errors():
    case = CASES["loo_std"]
    chunk = _chunk(case)
    blob = encode_micro_batch_record(chunk, [2, 0, 5], seq_parallel=4)
    v = RecordView(blob)
    assert (v.n_chunk, v.n_pack) == (len(chunk), 3) and len(blob) % 16 == 0
    assert v.total_tok == sum(len(chunk[i]["input_ids"]) for i in (2, 0, 5)) and (v.total_tok + v.padding) % 4 == 0
    assert v.array("input_ids", np.int32).tolist() == sum((chunk[i]["input_ids"] for i in (2, 0, 5)), [])
    assert np.array_equal(v.array("reward", np.float64), np.array([s["reward"] for s in chunk]))
    assert v.model_version == min(chunk[i]["model_version"] for i in (2, 0, 5))
    bytes_per_token = len(blob) / v.total_tok
    assert bytes_per_token < 40          # tiny samples here; ~12-16 B/token at 16 K tokens
    with pytest.raises(ValueError):
        RecordView(blob[:-16])
    with pytest.raises(ValueError):
        RecordView(b"\0" * 256)
    bad = copy.deepcopy(chunk)
    bad[0]["logprobs"] = bad[0]["logprobs"][:-1]
    with pytest.raises(ValueError, match="Target tokens"):
        encode_micro_batch_record(bad, [0])


def test_records_travel_through_the_stream_topic_as_bytes(tmp_path):
    from pipelinerl_b200 import streams
    streams.reset_streams_backend()
    streams.set_streams_backend("files")
    try:
        chunk = _chunk(CASES["loo_nostd"])
        blobs = [encode_micro_batch_record(chunk, [0, 1]), encode_micro_batch_record(chunk, [3])]
        spec = streams.StreamRangeSpec(exp_path=tmp_path, topic="training_data", partition_range=(0, 2))
        with streams.write_to_streams(spec) as w:
            w.write(blobs[0], 1)
            w.write({"kind": "json still works"}, 1)
            w.write(blobs[1], 1)
        d = tmp_path / "streams" / "training_data" / "0" / "1"
        assert (d / "0.jsonl").exists() and (d / "0.bin").exists()
        assert len((d / "0.jsonl").read_text().splitlines()) == 3          # one JSON document per record, as ever
        with streams.read_stream(streams.SingleStreamSpec(exp_path=tmp_path, topic="training_data", partition=1)) as r:
            got = r.read_available()
        assert got[0] == blobs[0] and got[1] == {"kind": "json still works"} and got[2] == blobs[1]
    finally:
        streams.reset_streams_backend()


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CASES))
def test_gpu_preprocess_is_bit_exact_against_the_reference(cuda_device, name):
    case = CASES[name]
    cfg = RLConfig(**case["config"])
    chunk = _chunk(case)
    blob = encode_micro_batch_record(chunk, list(range(len(chunk))), seq_parallel=case["seq_parallel"])
    pre = GpuPreprocessor(cuda_device, case["eos_token_id"], divide_advantage_by_std=cfg.divide_advantage_by_std)
    batch = pre.pack(blob)
    torch.cuda.synchronize()
    for k, w in case["batch"].items():
        g = getattr(batch, k)
        if isinstance(g, torch.Tensor):
            want = torch.tensor(w, dtype=g.dtype)
            assert g.device.type == "cuda" and g.shape == want.shape, (k, g.shape, want.shape)
            assert torch.equal(g.cpu(), want), f"{name}: column {k} differs"
        else:
            assert g == w, (k, g, w)


@pytest.mark.gpu
def test_gpu_preprocess_subset_pack_equals_host_path(cuda_device):
    """a micro-batch holding only SOME samples of the chunk: statistics still over the whole groups"""
    case = CASES["group64_std"]
    cfg = RLConfig(**case["config"])
    chunk = _chunk(case)
    entries = []
    for s in copy.deepcopy(chunk):
        enc = preprocess_fn(s, Tok(), seq_length=10_000, is_rl=True)
        for k in ("group_id", "rollout_index", "step_index", "finished", "model_version"):
            enc[k] = s[k]
        if "finish_reason" in s:
            enc["finish_reason"] = s["finish_reason"]
        entries.append(enc)
    entries = populate_rl_data(entries, Tok.eos_token_id, cfg)
    pick = [5, 70, 3, 116, 64]
    want = collate_packed([entries[i] for i in pick], Tok(), seq_parallel=2)
    pre = GpuPreprocessor(cuda_device, Tok.eos_token_id, divide_advantage_by_std=True)
    got = pre.pack(encode_micro_batch_record(chunk, pick, seq_parallel=2))
    for k in ("input_ids", "labels", "attention_mask", "position_ids", "segment_ids", "rewards", "advantages", "ref_logprobs",
              "old_logprobs", "group_tokens", "num_labels", "overflow", "seq_boundaries"):
        assert torch.equal(getattr(got, k).cpu(), getattr(want, k)), k
    assert got.model_version == want.model_version and got.padding == want.padding


@pytest.mark.gpu
def test_record_dealer_on_gpu_equals_tensor_dealer_on_host(cuda_device):
    """The same dealing rules with two transports: MicroBatchDealer (host tensors: populate_rl_data + collate_packed) and
    RecordDealer (binary records -> GPU pack).  Every micro-batch, to every rank, must be identical bit for bit."""
    from collections import deque

    from pipelinerl_b200.preprocess import MicroBatchDealer, RecordDealer, preprocess_dataset, record_entries
    rng = np.random.default_rng(7)
    samples = []
    for g in range(6):
        prompt = rng.integers(8, 97, size=int(rng.integers(3, 9))).tolist()
        for a in range(5):
            n_gen = int(rng.integers(2, 14))
            gen = rng.integers(8, 97, size=n_gen).tolist()
            fin = bool(rng.random() < 0.5)
            if fin:
                gen[-1] = Tok.eos_token_id
            samples.append({"input_ids": prompt + gen, "labels": [-100] * len(prompt) + gen,
                            "logprobs": (-rng.random(n_gen) * 3).tolist(), "ref_logprobs": (-rng.random(n_gen) * 3).tolist(),
                            "reward": float(rng.random() * 2 - 0.5), "group_id": f"g{g}", "finished": fin,
                            "metadata": {"model_version": 3 + a % 2, "rollout_index": a, "step_index": 0}})
    chunks = [samples[:15], samples[15:]]          # two chunks of whole groups (chunk_n_groups = 3)
    rl = RLConfig(divide_advantage_by_std=True)
    host_writes, gpu_writes = [], []
    host = MicroBatchDealer(Tok(), 40, 2, 4, write=lambda r, b: host_writes.append((r, b)))
    gpu = RecordDealer(Tok(), 40, 2, 4, write=lambda r, b: gpu_writes.append((r, b)))
    hq, gq = deque(), deque()
    for ch in chunks:
        hq.extend(preprocess_dataset(copy.deepcopy(ch), Tok(), seq_length=40, rl_config=rl))
        gq.extend(record_entries(copy.deepcopy(ch), seq_length=40))
        for dealer, q in ((host, hq), (gpu, gq)):
            while q:
                before = (len(q), dealer.published_samples, dealer.trainer_id)
                dealer.deal(q)
                if (len(q), dealer.published_samples, dealer.trainer_id) == before:
                    break
    assert len(host_writes) == len(gpu_writes) and len(host_writes) >= 4
    pre = GpuPreprocessor(cuda_device, Tok.eos_token_id, divide_advantage_by_std=True)
    n_records = 0
    for (hr, hb), (gr, gb) in zip(host_writes, gpu_writes):
        assert hr == gr
        if hb.sentinel:
            assert gb.sentinel
            continue
        assert isinstance(gb, bytes)
        n_records += 1
        got = pre.pack(gb)
        for k in ("input_ids", "labels", "attention_mask", "position_ids", "segment_ids", "rewards", "advantages",
                  "ref_logprobs", "old_logprobs", "group_tokens", "num_labels", "overflow", "seq_boundaries"):
            assert torch.equal(getattr(got, k).cpu(), getattr(hb, k)), k
        assert got.model_version == hb.model_version
    assert n_records >= 4
