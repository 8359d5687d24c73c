"""Trainer state as seen by the actor / preprocessor (reference: pipelinerl/state.py:20-65): a listener
thread tails the `weight_update_request` topic and exposes the last propagated weight version, the number of
samples the trainer has processed and whether training is done."""
from __future__ import annotations

import threading
import time
from pathlib import Path

from .streams import SingleStreamSpec, read_stream
from .weights import TRAINER_TOPIC


class TrainerState:
    def __init__(self, exp_path: Path):
        self.exp_path = Path(exp_path)
        self.propagated_weight_version: int | None = None
        self.samples_processed: int | None = None
        self.training_done = False
        self._done = threading.Event()
        self._reader = None

    def debug_mode_init(self) -> None:
        self.propagated_weight_version = 0
        self.samples_processed = 0
        self.training_done = True
        self._done.set()

    def _on_message(self, msg: dict) -> None:
        kind = msg.get("kind")
        if kind == "weight_update_success":
            self.propagated_weight_version = int(msg["version"])
        elif kind == "samples_processed":
            self.samples_processed = int(msg["samples_processed"])
        elif kind == "training_done":
            self.training_done = True
            self._done.set()

    def start_listening(self) -> None:
        spec = SingleStreamSpec(exp_path=self.exp_path, topic=TRAINER_TOPIC)

        def listen():
            with read_stream(spec) as reader:
                self._reader = reader
                for msg in reader.read():
                    self._on_message(msg)
        threading.Thread(target=listen, daemon=True, name="trainer-state").start()

    def stop(self) -> None:
        if self._reader is not None:
            self._reader.close()

    def wait_for_training_done(self, timeout: float | None = None) -> bool:
        return self._done.wait(timeout=timeout)

    def wait_for_processed_samples(self, poll: float = 0.05) -> int:
        while self.samples_processed is None:
            time.sleep(poll)
        return self.samples_processed

    def wait_for_model_version(self, poll: float = 0.05) -> int:
        while self.propagated_weight_version is None:
            time.sleep(poll)
        return self.propagated_weight_version
