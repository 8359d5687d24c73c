"""Preprocess stage: rollout groups -> RL columns -> packed micro-batches.

Host code mirroring pipelinerl/preprocess.py: `preprocess_dataset` (:145-189: prepare_rl_fields per sample,
populate_rl_data per chunk of whole groups, ref logprobs copied from old logprobs when kl_coef == 0, :160-161)
and the writer's packing rule (:598-637: greedily add samples to the current micro-batch while the token
count stays <= seq_length; `collate_packed`; round-robin over trainer ranks; sentinel batches equalise the
micro-batch counts per optimizer step).
"""
from __future__ import annotations

from collections import deque
from typing import Any, Callable

from .finetune.data import collate_packed, preprocess_fn
from .finetune.rl import RLConfig, populate_rl_data
from .finetune.types import PipelineBatchEncoding
from .finetune.utils import create_sentinel_batch


def preprocess_dataset(samples: list[dict[str, Any]], tokenizer, seq_length: int, rl_config: RLConfig) -> list[dict]:
    """samples: TrainingText.model_dump() dicts of WHOLE groups."""
    entries = []
    for s in samples:
        s = dict(s)
        if not s.get("ref_logprobs"):
            s["ref_logprobs"] = list(s["logprobs"])
        enc = preprocess_fn(s, tokenizer, seq_length, is_rl=True)
        meta = s.get("metadata", {})
        enc["group_id"] = s["group_id"]
        enc["rollout_index"] = meta.get("rollout_index", 0)
        enc["step_index"] = meta.get("step_index", 0)
        enc["finished"] = s.get("finished", False)
        enc["model_version"] = meta.get("model_version", 0)
        entries.append(enc)
    entries = populate_rl_data(entries, tokenizer.eos_token_id, rl_config)
    entries = [e for e in entries if len(e["input_ids"]) <= seq_length]
    if rl_config.filter_zero_advantage_groups:
        entries, _ = filter_zero_advantage_groups(entries)
    return entries


def filter_zero_advantage_groups(entries: list[dict], epsilon: float = 1e-6) -> tuple[list[dict], int]:
    """Drop every group none of whose samples carries an advantage with |a| > epsilon (all attempts got the same
    reward: no learning signal).  Applied by the preprocessor when `RLConfig.filter_zero_advantage_groups` is set
    (pipelinerl/preprocess.py:316-353, call site :548-552).  Groups keep their first-seen order, samples their order
    inside the group; returns (kept entries, number of samples dropped)."""
    by_group: dict[Any, list[dict]] = {}
    for e in entries:
        by_group.setdefault(e["group_id"], []).append(e)
    kept: list[dict] = []
    dropped = 0
    for members in by_group.values():
        if any(abs(a) > epsilon for m in members for a in m["advantages"]):
            kept.extend(members)
        else:
            dropped += len(members)
    return kept, dropped


def pack_micro_batches(entries: list[dict], tokenizer, seq_length: int, seq_parallel: int = 1,
                       pin_memory: bool = False, samples_per_step: int | None = None) -> list[PipelineBatchEncoding]:
    """Offline packing of a finished list of entries for ONE trainer.  With `samples_per_step` the micro-batches are cut
    exactly at the optimizer-step sample boundary (the writer's rule, preprocess.py:620-622: the trainer steps on
    `total_samples == target_samples`, finetune_loop.py:711), so a packed row never straddles two steps; entries left
    after the last complete step are flushed as a final (incomplete-step) micro-batch."""
    if samples_per_step is not None:
        out: list[PipelineBatchEncoding] = []
        dealer = MicroBatchDealer(tokenizer, seq_length, 1, samples_per_step, write=lambda _rank, b: out.append(b),
                                  seq_parallel=1, pin_memory=pin_memory)
        queue = deque(entries)
        while queue:
            dealer.deal(queue)
        if dealer.current_batch:
            out.append(collate_packed(dealer.current_batch, tokenizer, seq_parallel, pin_memory=pin_memory))
        return out
    batches, cur, cur_len = [], [], 0
    for e in entries:
        n = len(e["input_ids"])
        if cur and cur_len + n > seq_length:
            batches.append(collate_packed(cur, tokenizer, seq_parallel, pin_memory=pin_memory))
            cur, cur_len = [], 0
        cur.append(e)
        cur_len += n
    if cur:
        batches.append(collate_packed(cur, tokenizer, seq_parallel, pin_memory=pin_memory))
    return batches


class MicroBatchDealer:
    """The preprocessor's writer: deals packed micro-batches to the learner ranks (pipelinerl/preprocess.py:598-653,
    state set up at :463-483; trainer-side counterpart finetune_loop.py:674-713).

    * micro-batches go to the LEAD trainers round-robin (`trainer_id += seq_parallel`, modulo the number of trainers);
      with `seq_parallel > 1` a micro-batch is cut into slices for ranks `lead .. lead + seq_parallel - 1` (:356-367);
    * every lead trainer receives EXACTLY `samples_per_lead_per_step` samples per optimizer step: a micro-batch is closed
      when the next sample would overflow `seq_length` (:611-613) or when it completes the trainer's quota (:620-622);
    * a trainer whose quota is already complete gets a sentinel batch instead (:600-607) so that all ranks run the same
      number of micro-batches (their collectives stay aligned) until the step's last real batch is out;
    * the step ends when `published_samples == batch_boundary` and the round-robin pointer is back at rank 0 (:650).
    A partially filled micro-batch survives between `deal()` calls, exactly as the reference's loop variables do.

    `write(rank, batch)` is the stream write (`data_writer.write(batch, rank)`: topic `training_data`, partition = rank).
    """

    def __init__(self, tokenizer, seq_length: int, num_trainers: int, samples_per_lead_per_step: int,
                 write: Callable[[int, PipelineBatchEncoding], None], seq_parallel: int = 1, published_samples: int = 0,
                 pin_memory: bool = False):
        assert num_trainers % seq_parallel == 0
        self.tokenizer, self.seq_length, self.write = tokenizer, seq_length, write
        self.num_trainers, self.seq_parallel, self.pin_memory = num_trainers, seq_parallel, pin_memory
        self.num_lead_trainers = num_trainers // seq_parallel
        self.samples_per_lead_per_step = samples_per_lead_per_step
        self.train_batch_size = samples_per_lead_per_step * self.num_lead_trainers          # :468
        assert published_samples % self.num_lead_trainers == 0                               # :472
        self.published_samples = published_samples
        self.samples_per_trainer = {i: published_samples // num_trainers for i in range(0, num_trainers, seq_parallel)}
        self.trainer_id = 0
        self.current_batch: list[dict] = []
        self.current_length = 0
        self.batch_boundary = published_samples + self.train_batch_size
        self.target_samples_per_lead = self.samples_per_trainer[0] + samples_per_lead_per_step
        self.max_model_version = 0
        self.steps_dealt = 0

    def _collate(self, entries: list[dict]):
        return collate_packed(entries, self.tokenizer, self.seq_parallel, pin_memory=self.pin_memory)

    def _emit(self, batch) -> None:
        if self.seq_parallel > 1:
            for index, piece in enumerate(batch.make_slices(self.seq_parallel)):
                self.write(self.trainer_id + index, piece)
        else:
            self.write(self.trainer_id, batch)
        self.trainer_id = (self.trainer_id + self.seq_parallel) % self.num_trainers

    def deal(self, queue: "deque[dict]", keep_going: bool = False) -> bool:
        """Consume entries from `queue` (left to right) until it is empty or one optimizer step's worth of micro-batches
        has been written; returns True when a step boundary was reached (reference `batch_done`).  `keep_going` is the
        reference's `dataset_buffer_size` mode (:594,:654-656): keep dealing steps while the queue holds data, and do not
        stop on an empty queue before the step is complete."""
        if queue:
            self.max_model_version = max(e["model_version"] for e in queue)                  # :586
        batch_done = False
        while (queue and not batch_done) or (keep_going and not batch_done):
            tid = self.trainer_id
            if self.samples_per_trainer[tid] == self.target_samples_per_lead:
                self._emit(create_sentinel_batch(device=None, tokenizer=self.tokenizer, model_version=self.max_model_version))
            else:
                time_to_write = False
                while queue:
                    n = len(queue[0]["input_ids"])
                    if self.current_length + n > self.seq_length:
                        time_to_write = True
                        break
                    self.current_batch.append(queue.popleft())
                    self.current_length += n
                    if len(self.current_batch) + self.samples_per_trainer[tid] == self.target_samples_per_lead:
                        time_to_write = True
                        break
                if time_to_write:
                    assert self.current_batch, "a sample longer than seq_length reached the writer"
                    batch = self._collate(self.current_batch)
                    n_samples = len(self.current_batch)
                    self.current_batch, self.current_length = [], 0
                    self._emit(batch)
                    self.published_samples += n_samples
                    self.samples_per_trainer[tid] += n_samples
                elif keep_going and not queue:
                    break      # nothing more to pack right now (the reference spins here until data arrives)
            batch_done = self.published_samples == self.batch_boundary and self.trainer_id == 0
            if batch_done:
                self.batch_boundary += self.train_batch_size
                self.target_samples_per_lead += self.samples_per_lead_per_step
                self.steps_dealt += 1
                if keep_going and queue:
                    batch_done = False
        return batch_done


def record_entries(samples: list[dict[str, Any]], seq_length: int) -> list[dict[str, Any]]:
    """Entries for `RecordDealer` from the TrainingText dicts of WHOLE groups (one `actor`-topic chunk): no per-token host
    work -- each entry keeps the sample's own arrays plus a reference to the chunk's scalar table (reward, group / rollout /
    step ids and length of EVERY sample of the chunk), which is all the GPU needs for the leave-one-out statistics.
    Samples longer than `seq_length` are dropped after the table is built, as the reference drops them after
    populate_rl_data (preprocess.py:178-186)."""
    table = []
    for s in samples:
        meta = s.get("metadata", {})
        table.append({"reward": s["reward"], "group_id": s["group_id"], "rollout_index": meta.get("rollout_index", 0),
                      "step_index": meta.get("step_index", 0), "n_tok": len(s["input_ids"])})
    out = []
    for pos, s in enumerate(samples):
        if len(s["input_ids"]) > seq_length:
            continue
        e = {"input_ids": s["input_ids"], "labels": s["labels"], "logprobs": s["logprobs"],
             "finished": s.get("finished", False), "model_version": s.get("metadata", {}).get("model_version", 0),
             "_chunk": table, "_chunk_pos": pos}
        if s.get("ref_logprobs"):
            e["ref_logprobs"] = s["ref_logprobs"]
        if "finish_reason" in s:
            e["finish_reason"] = s["finish_reason"]
        out.append(e)
    return out


class RecordDealer(MicroBatchDealer):
    """MicroBatchDealer whose micro-batches are binary records (records.py) instead of host tensors: the same dealing
    rules (they only look at sample LENGTHS), but what is written to `training_data` is 12-16 bytes per token and the RL
    columns are produced on the learner's GPU (records.GpuPreprocessor / csrc/preprocess_pack.cu).  Sentinel batches stay
    the small JSON objects they are.  A micro-batch may mix samples of several chunks: its scalar table is the union of
    theirs (group ids are unique across chunks, so the statistics keys stay disjoint)."""

    def __init__(self, *a, use_ref_logprobs: bool = True, **kw):
        super().__init__(*a, **kw)
        if self.seq_parallel != 1:
            raise NotImplementedError("binary records are not sliced across sequence-parallel ranks")
        self.use_ref_logprobs = use_ref_logprobs

    def _collate(self, entries: list[dict]):
        from .records import encode_micro_batch_record
        chunk: list[dict] = []
        base: dict[int, int] = {}
        pack = []
        for e in entries:
            table = e["_chunk"]
            if id(table) not in base:
                base[id(table)] = len(chunk)
                chunk.extend(dict(row) for row in table)
            i = base[id(table)] + e["_chunk_pos"]
            chunk[i].update({k: v for k, v in e.items() if not k.startswith("_")})
            pack.append(i)
        return encode_micro_batch_record(chunk, pack, seq_parallel=1, use_ref_logprobs=self.use_ref_logprobs)
