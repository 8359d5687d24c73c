"""LLM handle used inside rollout plugins (reference: pipelinerl/llm.py).

`TrainableLLM` keeps the constructor fields plugins and the actor use (base_url, model_name,
tokenizer_name, parameters, collect_logprobs, chat_template_kwargs; actor.py:820-830) and the
attributes plugins read (`llm.parameters["max_tokens"]`, `llm.tokenizer`).  `base_url` addresses an
in-process sampler engine ("inproc://<name>", registered by pipelinerl_b200.serving.EngineServer)
instead of an HTTP vLLM server; the request still carries token ids in and token ids + logprobs out.
"""
from __future__ import annotations

import datetime
from typing import Any
from uuid import uuid4

from pydantic import BaseModel, Field


class Prompt(BaseModel):
    id: str = Field(default_factory=lambda: str(uuid4()))
    tools: list[dict] | None = None
    messages: list[dict] = Field(default_factory=list)
    token_ids: list[int] = Field(default_factory=list)

    @staticmethod
    def from_user_message(content: str) -> "Prompt":
        return Prompt(messages=[{"role": "user", "content": content}])

    def __bool__(self) -> bool:
        return bool(self.messages)


class LLMOutput(BaseModel):
    role: str = "assistant"
    content: str = ""
    tool_calls: list[Any] | None = None


class TokenLogprob(BaseModel):
    logprob: float
    token_id: int
    generated: int = 1


class LLMCall(BaseModel):
    timestamp: str = Field(default_factory=lambda: datetime.datetime.now().isoformat())
    prompt: Prompt
    output: LLMOutput
    prompt_length_tokens: int = -1
    output_length_tokens: int = -1
    cached: bool = False
    llm_info: dict = Field(default_factory=dict)
    cost: float = 0
    logprobs: list[TokenLogprob] = Field(default_factory=list, exclude=True)


class SyntheticTokenizer:
    """Offline stand-in for a HF tokenizer (no tokenizer files exist in this environment): whitespace
    words hashed into a fixed vocabulary, with the small interface the plugin surface touches
    (apply_chat_template, decode, eos/bos tokens).  Any HF tokenizer object can be used instead."""

    def __init__(self, vocab_size: int = 1024, eos_token_id: int = 2):
        self.vocab_size, self.eos_token_id, self.bos_token_id = vocab_size, eos_token_id, 1
        self.eos_token, self.bos_token = "<eos>", ""
        self.padding_side = "right"
        self._ROLE = {"system": 3, "user": 4, "assistant": 5}

    def _word(self, w: str) -> int:
        if w.isdigit() and int(w) < 400:  # numbers are stable tokens: guessing-game answers round-trip
            return 16 + int(w)
        h = 2166136261
        for ch in w.encode():
            h = ((h ^ ch) * 16777619) & 0xFFFFFFFF
        return 416 + h % (self.vocab_size - 416)

    def encode(self, text: str) -> list[int]:
        return [self._word(w) for w in text.split()]

    def decode(self, ids) -> str:
        out = []
        for i in ids:
            if i == self.eos_token_id:
                out.append(self.eos_token)
            elif 16 <= i < 416:
                out.append(str(i - 16))
            else:
                out.append(f"<{i}>")
        return " ".join(out)

    def apply_chat_template(self, conversation, tokenize: bool = True, add_generation_prompt: bool = False, **kw):
        ids, text = [], []
        for m in conversation:
            ids.append(self._ROLE.get(m["role"], 6))
            ids.extend(self.encode(str(m.get("content") or "")))
            ids.append(7)  # end of turn
            text.append(f"<|{m['role']}|> {m.get('content') or ''} <|end|>")
        if add_generation_prompt:
            ids.append(self._ROLE["assistant"])
            text.append("<|assistant|>")
        return ids if tokenize else " ".join(text) + " "


class TrainableLLM:
    def __init__(self, base_url: str, model_name: str = "", tokenizer_name: str = "", parameters: dict | None = None,
                 collect_logprobs: bool = True, chat_template_kwargs: dict | None = None, tokenizer=None,
                 api_token: str = ""):
        self.base_url = base_url
        self.model_name = model_name
        self.tokenizer_name = tokenizer_name or model_name
        self.parameters = dict(parameters or {})
        self.collect_logprobs = collect_logprobs
        self.chat_template_kwargs = chat_template_kwargs
        self.api_token = api_token
        self.tokenizer = tokenizer
        self._calls = 0

    def load_tokenizer(self):
        """The model's tokenizer from the local HF cache (no network on the boxes this runs on).  The synthetic tokenizer
        is used ONLY when asked for by name (`tokenizer_name="synthetic"`, benches / plumbing tests) or injected through
        `tokenizer=`: a missing or misspelt real tokenizer raises instead of silently training on hashed word ids."""
        if self.tokenizer is None:
            if self.tokenizer_name in ("synthetic", "synthetic-tokenizer"):
                self.tokenizer = SyntheticTokenizer()
            else:
                import transformers
                self.tokenizer = transformers.AutoTokenizer.from_pretrained(self.tokenizer_name, local_files_only=True)
        return self.tokenizer

    def log_output(self, prompt: Prompt, output: LLMOutput, cached: bool = False, count_tokens: bool = True) -> LLMCall:
        self._calls += 1
        return LLMCall(prompt=prompt, output=output, cached=cached, llm_info={"model_name": self.model_name})

    def get_stats(self) -> dict:
        return {"calls": self._calls}
