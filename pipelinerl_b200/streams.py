"""Topic streams between the pipeline stages (file backend).

Public API of the reference kept as is (pipelinerl/streams.py:33-69, 390-423):
    set_streams_backend("files"); SingleStreamSpec(exp_path, topic, instance=0, partition=0);
    StreamRangeSpec(..., partition_range=(lo, hi));
    with write_to_streams(spec, mode="a") as w: w.write(obj, partition=None)
    with read_stream(spec) as r: for x in r.read(): ...      (blocking tail from the beginning)
File layout <exp>/streams/<topic>/<instance>/<partition>/0.jsonl (:238-243), one JSON object per
line, flushed per write; pydantic models, tensors and numpy arrays are serialised as plain lists.
Redis is a deployment option of the reference, outside the hot path.
"""
from __future__ import annotations

import json
import os
import time
from pathlib import Path
from typing import Any, Iterator, Literal

import numpy as np
from pydantic import BaseModel

_REREAD_DELAY = 0.02
_RECHECK_DELAY = 0.5
_backend: str | None = None


def set_streams_backend(backend: str, **kwargs) -> None:
    global _backend
    if _backend is not None and _backend != backend:
        raise ValueError("Backend already set. Cannot change it.")
    if backend != "files":
        raise ValueError(f"Invalid backend: {backend}. This build ships the 'files' backend.")
    _backend = backend


def reset_streams_backend() -> None:  # test helper
    global _backend
    _backend = None


def raise_if_backend_not_set() -> None:
    if _backend is None:
        raise ValueError("Backend not set. Please call set_streams_backend() first.")


class SingleStreamSpec(BaseModel):
    exp_path: Path
    topic: str
    instance: int = 0
    partition: int = 0

    def __str__(self) -> str:
        return f"{self.topic}/{self.instance}/{self.partition}"


class StreamRangeSpec(BaseModel):
    exp_path: Path
    topic: str
    instance: int = 0
    partition_range: tuple[int, int]

    def __str__(self) -> str:
        return f"{self.topic}/{self.instance}/{self.partition_range[0]}-{self.partition_range[1]}"


StreamSpec = SingleStreamSpec | StreamRangeSpec


def stream_dir(exp_path: Path, topic: str, instance: int, partition: int) -> Path:
    return Path(exp_path) / "streams" / topic / str(instance) / str(partition)


def stream_file(directory: Path, shard_id: int) -> Path:
    return directory / f"{shard_id}.jsonl"


def binary_file(directory: Path, shard_id: int) -> Path:
    return directory / f"{shard_id}.bin"


BIN_KEY = "__bin__"


def _plain(obj: Any) -> Any:
    if isinstance(obj, BaseModel):
        return _plain(obj.model_dump())
    if hasattr(obj, "model_dump") and callable(obj.model_dump):
        return _plain(obj.model_dump())
    if isinstance(obj, dict):
        return {k: _plain(v) for k, v in obj.items()}
    if isinstance(obj, (list, tuple)):
        return [_plain(v) for v in obj]
    if isinstance(obj, np.ndarray):
        return obj.tolist()
    if isinstance(obj, (np.integer,)):
        return int(obj)
    if isinstance(obj, (np.floating,)):
        return float(obj)
    if isinstance(obj, Path):
        return str(obj)
    if hasattr(obj, "detach") and hasattr(obj, "cpu"):  # torch.Tensor without importing torch here
        return obj.detach().cpu().tolist()
    return obj


class FileStreamWriter:
    def __init__(self, stream: SingleStreamSpec, mode: Literal["w", "a"] = "a"):
        self.stream, self.mode = stream, mode

    def __enter__(self):
        d = stream_dir(self.stream.exp_path, self.stream.topic, self.stream.instance, self.stream.partition)
        os.makedirs(d, exist_ok=True)
        self._file = open(stream_file(d, 0), self.mode)
        return self

    def __exit__(self, *exc):
        self._file.close()
        if getattr(self, "_bin", None) is not None:
            self._bin.close()

    def write(self, data: Any, partition: int | None = None) -> None:
        if partition is not None:
            raise ValueError("a single-partition writer takes no partition argument")
        if isinstance(data, (bytes, bytearray, memoryview)):
            data = self._write_binary(data)
        self._file.write(json.dumps(_plain(data), separators=(",", ":")))
        self._file.write("\n")
        self._file.flush()

    def _write_binary(self, payload) -> dict:
        """Binary records (records.py: packed micro-batches at 12-16 B/token instead of ~170 B/token of JSON) keep the
        topic protocol -- one JSON document per line, same files -- by travelling in a sibling `0.bin`: the line is a
        pointer {"__bin__": [offset, nbytes]} that the reader resolves back to `bytes`."""
        if getattr(self, "_bin", None) is None:
            d = stream_dir(self.stream.exp_path, self.stream.topic, self.stream.instance, self.stream.partition)
            self._bin = open(binary_file(d, 0), "ab" if self.mode == "a" else "wb")
        self._bin.seek(0, os.SEEK_END)
        off = self._bin.tell()
        self._bin.write(payload)
        self._bin.flush()          # payload is on disk before the pointer line becomes visible
        return {BIN_KEY: [off, len(payload)]}


class RoundRobinFileStreamWriter:
    def __init__(self, streams: StreamRangeSpec, mode: Literal["w", "a"] = "a"):
        self.streams, self.mode = streams, mode

    def __enter__(self):
        lo, hi = self.streams.partition_range
        self._writers = [FileStreamWriter(SingleStreamSpec(exp_path=self.streams.exp_path, topic=self.streams.topic,
                                                           instance=self.streams.instance, partition=p), self.mode)
                         for p in range(lo, hi)]
        for w in self._writers:
            w.__enter__()
        self._next = 0
        return self

    def __exit__(self, *exc):
        for w in self._writers:
            w.__exit__(*exc)

    def write(self, data: Any, partition: int | None = None) -> None:
        if partition is not None:
            self._writers[partition - self.streams.partition_range[0]].write(data)
            return
        self._writers[self._next].write(data)
        self._next = (self._next + 1) % len(self._writers)


class FileStreamReader:
    def __init__(self, stream: SingleStreamSpec, poll: float = _REREAD_DELAY):
        self.stream, self.poll = stream, poll
        self._stop = False

    def __enter__(self):
        d = stream_dir(self.stream.exp_path, self.stream.topic, self.stream.instance, self.stream.partition)
        path = stream_file(d, 0)
        while not path.exists():
            if self._stop:
                raise FileNotFoundError(path)
            time.sleep(_RECHECK_DELAY)
        self._path = path
        self._file = open(path, "r")
        return self

    def __exit__(self, *exc):
        self._file.close()
        if getattr(self, "_bin", None) is not None:
            self._bin.close()

    def close(self) -> None:
        self._stop = True

    def _resolve(self, doc: Any) -> Any:
        if isinstance(doc, dict) and len(doc) == 1 and BIN_KEY in doc:
            off, n = doc[BIN_KEY]
            if getattr(self, "_bin", None) is None:
                self._bin = open(binary_file(self._path.parent, 0), "rb")
            self._bin.seek(off)
            payload = self._bin.read(n)
            if len(payload) != n:
                raise IOError(f"binary record truncated: wanted {n} bytes at {off}, got {len(payload)}")
            return payload
        return doc

    def read(self) -> Iterator[Any]:
        """Blocking tail; a partially written last line is re-read after a short delay."""
        pos = self._file.tell()
        while not self._stop:
            line = self._file.readline()
            if line.endswith("\n"):
                yield self._resolve(json.loads(line))
                pos = self._file.tell()
            else:
                self._file.seek(pos)
                time.sleep(self.poll)

    def read_available(self) -> list[Any]:
        """Non-blocking: everything complete that is in the file now (used by polling stages and tests)."""
        out = []
        pos = self._file.tell()
        while True:
            line = self._file.readline()
            if not line.endswith("\n"):
                self._file.seek(pos)
                return out
            out.append(self._resolve(json.loads(line)))
            pos = self._file.tell()


def read_stream(stream: SingleStreamSpec) -> FileStreamReader:
    raise_if_backend_not_set()
    if not isinstance(stream, SingleStreamSpec):
        raise ValueError(f"Invalid stream spec: {stream}")
    return FileStreamReader(stream)


def write_to_streams(streams: StreamSpec, mode: Literal["w", "a"] = "a"):
    raise_if_backend_not_set()
    if isinstance(streams, SingleStreamSpec):
        return FileStreamWriter(streams, mode)
    if isinstance(streams, StreamRangeSpec):
        return RoundRobinFileStreamWriter(streams, mode)
    raise ValueError(f"Invalid stream spec: {streams}")
