"""Binary micro-batch records and the GPU-resident preprocess.

The reference moves training data as JSON: the preprocessor serialises twelve `[1, T]` tensors per micro-batch to a JSONL
line (pipelinerl/streams.py:269-277, ~170 bytes per token) and the trainer parses and validates it back
(finetune_loop.py:109); before that, `populate_rl_data` runs a pandas pipeline and `collate_packed` builds the tensors from
Python lists (rl/__init__.py:453-570, data.py:215-283).  Here the preprocessor only does the integer bookkeeping that needs
no token data (the dealing of samples to trainer ranks by LENGTH, preprocess.MicroBatchDealer) and ships a compact record:

    header | chunk scalar table (reward, group / step / rollout ids, lengths of EVERY sample of the chunk of whole groups)
           | packed-sample table | input_ids, labels (int32) | sampler logprobs (+ reference logprobs) (float32)

12-16 bytes per token.  The learner uploads it with ONE host->device copy and `prl_preprocess_pack`
(csrc/preprocess_pack.cu) computes the leave-one-out advantages with pandas' own recurrences and writes all
PipelineBatchEncoding columns on the GPU -- the micro-batch never exists as host tensors.  The topic API is unchanged:
records travel as `bytes` through `write_to_streams` / `read_stream` (streams.py keeps a JSONL line per record that points
into a sibling binary file).
"""
from __future__ import annotations

import ctypes as C
from typing import Any, Sequence

import numpy as np

MAGIC = 0x4D4C5250          # "PRLM"
VERSION = 1
_HEADER_WORDS = 16          # int64 each
IGNORE_INDEX = -100

_REASON_CODE = {"length": 1, "stop": 2, "content_filter": 2}


def _pad16(b: bytes) -> bytes:
    r = (-len(b)) % 16
    return b + b"\0" * r if r else b


def sample_flags(entry: dict[str, Any]) -> int:
    """bit 0 finished; bits 1-2 finish_reason (rl/__init__.py:541-554): 1 length, 2 stop | content_filter, 0 otherwise"""
    reason = entry.get("finish_reason")
    code = _REASON_CODE.get(reason.strip().lower(), 0) if isinstance(reason, str) else 0
    return (1 if entry.get("finished") else 0) | (code << 1)


def encode_micro_batch_record(chunk: Sequence[dict[str, Any]], pack: Sequence[int], seq_parallel: int = 1,
                              use_ref_logprobs: bool = True) -> bytes:
    """chunk: the samples of whole rollout groups the statistics are taken over (what `populate_rl_data` receives:
    input_ids, labels, reward | rewards, logprobs | old_logprobs, ref_logprobs, group_id, rollout_index, step_index,
    finished, finish_reason?, model_version).  pack: indices into `chunk` of the samples of THIS micro-batch, row order."""
    n_chunk, n_pack = len(chunk), len(pack)
    if n_chunk < 1 or n_pack < 1:
        raise ValueError("empty micro-batch record")
    reward = np.empty(n_chunk, dtype=np.float64)
    stat_slot = np.empty(n_chunk, dtype=np.int32)
    rollout_slot = np.empty(n_chunk, dtype=np.int32)
    group_slot = np.empty(n_chunk, dtype=np.int32)
    n_tok_all = np.empty(n_chunk, dtype=np.int32)
    stats: dict[tuple, int] = {}
    rollouts: dict[tuple, int] = {}
    groups: dict[Any, int] = {}
    for i, e in enumerate(chunk):
        reward[i] = e["reward"] if "reward" in e else e["rewards"][0]
        g = e["group_id"]
        group_slot[i] = groups.setdefault(g, len(groups))
        stat_slot[i] = stats.setdefault((g, e.get("step_index", 0)), len(stats))
        rollout_slot[i] = rollouts.setdefault((g, e.get("rollout_index", 0)), len(rollouts))
        n_tok_all[i] = e["n_tok"] if "n_tok" in e else len(e["input_ids"])
    ids_parts, lab_parts, lp_parts, ref_parts = [], [], [], []
    tok_off = np.zeros(n_pack + 1, dtype=np.int32)
    lp_off = np.zeros(n_pack + 1, dtype=np.int32)
    flags = np.empty(n_pack, dtype=np.int32)
    has_ref = False
    for p, i in enumerate(pack):
        e = chunk[i]
        ids = np.asarray(e["input_ids"], dtype=np.int32)
        lab = np.asarray(e["labels"], dtype=np.int32)
        lps = e["logprobs"] if "logprobs" in e else _strip_left(e["old_logprobs"], lab)
        lp = np.asarray(lps, dtype=np.float32)
        if ids.shape != lab.shape:
            raise ValueError("input_ids and labels differ in length")
        n_target = int(np.count_nonzero(lab != IGNORE_INDEX))
        if n_target != lp.size:                                   # the reference's assert (rl/__init__.py:583)
            raise ValueError(f"Target tokens: {n_target}, old logprobs: {lp.size}")
        ref = e.get("ref_logprobs")
        if use_ref_logprobs and ref is not None and len(ref):
            ref = ref if "logprobs" in e else _strip_left(ref, lab)
            ref = np.asarray(ref, dtype=np.float32)
            if ref.size != lp.size:
                raise ValueError("ref_logprobs and logprobs differ in length")
            ref_parts.append(ref)
            has_ref = True
        elif has_ref:
            raise ValueError("either every packed sample carries ref_logprobs or none does")
        ids_parts.append(ids)
        lab_parts.append(lab)
        lp_parts.append(lp)
        tok_off[p + 1] = tok_off[p] + ids.size
        lp_off[p + 1] = lp_off[p] + lp.size
        flags[p] = sample_flags(e)
    if has_ref and len(ref_parts) != n_pack:
        raise ValueError("either every packed sample carries ref_logprobs or none does")
    total_tok, total_lp = int(tok_off[-1]), int(lp_off[-1])
    padding = (-total_tok) % seq_parallel if seq_parallel > 1 else 0
    versions = [int(chunk[i].get("model_version", 0)) for i in pack]
    header = np.zeros(_HEADER_WORDS, dtype=np.int64)
    header[:13] = [MAGIC, VERSION, n_chunk, n_pack, padding, total_tok, total_lp, len(stats), len(rollouts), len(groups),
                   int(has_ref), min(versions), max(versions)]
    sections = [header.tobytes(), reward.tobytes(), stat_slot.tobytes(), rollout_slot.tobytes(), group_slot.tobytes(),
                n_tok_all.tobytes(), np.asarray(pack, dtype=np.int32).tobytes(), flags.tobytes(), tok_off.tobytes(),
                lp_off.tobytes(), np.concatenate(ids_parts).tobytes(), np.concatenate(lab_parts).tobytes(),
                np.concatenate(lp_parts).tobytes()]
    if has_ref:
        sections.append(np.concatenate(ref_parts).tobytes())
    return b"".join(_pad16(s) for s in sections)


def _strip_left(values, labels: np.ndarray):
    """token-aligned column ([0] * prompt + per-label values, prepare_rl_fields) -> the per-label values"""
    n_target = int(np.count_nonzero(labels != IGNORE_INDEX))
    return list(values)[len(values) - n_target:]


class RecordView:
    """Section offsets of an encoded record (host side: parses the 128-byte header only)."""

    def __init__(self, blob):
        mv = memoryview(blob)
        h = np.frombuffer(mv[:_HEADER_WORDS * 8], dtype=np.int64)
        if int(h[0]) != MAGIC or int(h[1]) != VERSION:
            raise ValueError("not a PRLM micro-batch record (bad magic / version)")
        (self.n_chunk, self.n_pack, self.padding, self.total_tok, self.total_lp, self.n_stat_slots, self.n_rollout_slots,
         self.n_groups, has_ref, self.model_version, self.max_model_version) = (int(x) for x in h[2:13])
        self.has_ref = bool(has_ref)
        self.nbytes = len(mv)
        sizes = [("reward", 8 * self.n_chunk), ("stat_slot", 4 * self.n_chunk), ("rollout_slot", 4 * self.n_chunk),
                 ("group_slot", 4 * self.n_chunk), ("n_tok_all", 4 * self.n_chunk), ("pack_idx", 4 * self.n_pack),
                 ("pack_flags", 4 * self.n_pack), ("tok_off", 4 * (self.n_pack + 1)), ("lp_off", 4 * (self.n_pack + 1)),
                 ("input_ids", 4 * self.total_tok), ("labels", 4 * self.total_tok), ("logprobs", 4 * self.total_lp)]
        if self.has_ref:
            sizes.append(("ref_logprobs", 4 * self.total_lp))
        self.offsets, at = {}, _HEADER_WORDS * 8
        for name, nb in sizes:
            self.offsets[name] = (at, nb)
            at += (nb + 15) // 16 * 16
        if at != self.nbytes:
            raise ValueError(f"truncated or oversized record: {self.nbytes} bytes, sections need {at}")
        self._mv = mv

    def array(self, name: str, dtype) -> np.ndarray:
        off, nb = self.offsets[name]
        return np.frombuffer(self._mv[off:off + nb], dtype=dtype)

    @property
    def tokens(self) -> int:
        return self.total_tok + self.padding


class GpuPreprocessor:
    """Learner side: record bytes -> device-resident PipelineBatchEncoding (one H2D copy + three kernel launches)."""

    def __init__(self, device, eos_token_id: int, divide_advantage_by_std: bool = True):
        import torch
        from . import _lib
        self.dev = torch.device(device)
        if self.dev.type != "cuda":
            raise RuntimeError("GpuPreprocessor needs a CUDA device: pipelinerl_b200 has no CPU fallback")
        self.lib = _lib.load()
        self.eos, self.divide = int(eos_token_id), bool(divide_advantage_by_std)
        self._staging = None

    def pack(self, blob) -> "PipelineBatchEncoding":  # noqa: F821
        import torch
        from . import _lib
        from .finetune.types import PipelineBatchEncoding
        v = RecordView(blob)
        n = v.nbytes
        if self._staging is None or self._staging.numel() < n:
            self._staging = torch.empty(max(n, 1 << 20), dtype=torch.uint8).pin_memory()
        self._staging[:n].copy_(torch.frombuffer(bytearray(blob) if isinstance(blob, bytes) else blob, dtype=torch.uint8))
        dblob = torch.empty(n, dtype=torch.uint8, device=self.dev)
        dblob.copy_(self._staging[:n], non_blocking=True)
        base = dblob.data_ptr()
        rec = _lib.MbRecord()
        rec.n_chunk, rec.n_pack, rec.padding, rec.total_tok, rec.total_lp = v.n_chunk, v.n_pack, v.padding, v.total_tok, v.total_lp
        rec.n_stat_slots, rec.n_rollout_slots, rec.n_groups = v.n_stat_slots, v.n_rollout_slots, v.n_groups
        for name in ("reward", "stat_slot", "rollout_slot", "group_slot", "n_tok_all", "pack_idx", "pack_flags", "tok_off",
                     "lp_off", "input_ids", "labels", "logprobs"):
            setattr(rec, name, base + v.offsets[name][0])
        rec.ref_logprobs = base + v.offsets["ref_logprobs"][0] if v.has_ref else None
        T = v.tokens
        i64 = torch.empty(5, T, dtype=torch.int64, device=self.dev)
        f32 = torch.empty(7, T, dtype=torch.float32, device=self.dev)
        bounds = torch.empty(v.n_pack + 1 + (1 if v.padding else 0), dtype=torch.int32, device=self.dev)
        cols = _lib.MbColumns()
        int_names = ("input_ids", "labels", "attention_mask", "position_ids", "segment_ids")
        f_names = ("rewards", "advantages", "ref_logprobs", "old_logprobs", "group_tokens", "num_labels", "overflow")
        for k, name in enumerate(int_names):
            setattr(cols, name, i64[k].data_ptr())
        for k, name in enumerate(f_names):
            setattr(cols, name, f32[k].data_ptr())
        cols.seq_boundaries = bounds.data_ptr()
        ws = torch.empty(int(self.lib.prl_preprocess_workspace_bytes(v.n_pack, v.n_stat_slots, v.n_rollout_slots, v.n_groups)),
                         dtype=torch.uint8, device=self.dev)
        _lib.check(self.lib.prl_preprocess_pack(C.byref(rec), int(self.divide), self.eos, C.byref(cols), ws.data_ptr(),
                                                ws.numel(), _lib.stream_ptr()))
        fields = {name: i64[k][None] for k, name in enumerate(int_names)}
        fields.update({name: f32[k][None] for k, name in enumerate(f_names)})
        return PipelineBatchEncoding(**fields, seq_boundaries=bounds, model_version=v.model_version, is_packed=True,
                                     padding=v.padding)
