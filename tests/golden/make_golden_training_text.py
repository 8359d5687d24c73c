"""Golden fixture for row a2: the reference's `make_training_text` (pipelinerl/async_llm.py:215-346) executed on a
locally built chat tokenizer (tests.helpers.tiny_chat_tokenizer: no files, no network).

    python tests/golden/make_golden_training_text.py      (authoring container: needs /root/reference)

litellm and jsonref (third-party, not installed here) are only used by the reference for type aliases / schema
helpers on this path and are stubbed; the function body that runs is the reference's.
"""
from __future__ import annotations

import json
import sys
import types
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
from tests.helpers import tiny_chat_tokenizer  # noqa: E402

OUT = Path(__file__).resolve().parent

CASES = [
    dict(name="stop", messages=[{"role": "user", "content": "what is the answer to life"}], content="it is forty two",
         finish_reason="stop", gen=[11, 4, 12, 13, 30], prompt_len=9, out_len=5),
    dict(name="length", messages=[{"role": "system", "content": "guess a number"}, {"role": "user", "content": "hello"}],
         content="my guess is", finish_reason="length", gen=[20, 14, 4], prompt_len=12, out_len=3),
    dict(name="eos_in_content", messages=[{"role": "user", "content": "hello world"}], content="too high<|im_end|>",
         finish_reason=None, gen=[21, 22, 31], prompt_len=6, out_len=3),
    dict(name="no_eos_no_reason", messages=[{"role": "user", "content": "hello world"}], content="too low",
         finish_reason=None, gen=[21, 23], prompt_len=6, out_len=2),
    dict(name="chat_kwargs", messages=[{"role": "user", "content": "hello"}], content="correct", finish_reason="stop",
         gen=[24], prompt_len=8, out_len=1, chat_template_kwargs={"style": "terse"}),
    dict(name="tools", messages=[{"role": "user", "content": "what is the answer"}], content="", finish_reason="stop",
         gen=[27, 28], prompt_len=20, out_len=2,
         tools=[{"type": "function", "function": {"name": "get_weather", "parameters": {"type": "object"}}}],
         tool_calls=[{"id": "call_1", "name": "get_weather", "arguments": "{\"city\": \"Paris\"}"}]),
]


def _import_reference():
    import transformers  # noqa: F401
    from pydantic import BaseModel

    def stub(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules.setdefault(name, m)
        return sys.modules[name]

    class _OC:
        @staticmethod
        def to_container(x, resolve=True):
            return x
    stub("omegaconf", DictConfig=dict, ListConfig=list, OmegaConf=_OC)
    stub("jsonref")
    stub("accelerate", Accelerator=object)

    class Function(BaseModel):
        name: str
        arguments: str

    class ChatCompletionMessageToolCall(BaseModel):
        id: str
        type: str = "function"
        function: Function

    class Message(BaseModel):           # the attributes litellm.utils.Message exposes on this path
        role: str = "assistant"
        content: str | None = None
        tool_calls: list[ChatCompletionMessageToolCall] | None = None
    lit = stub("litellm", ChatCompletionMessageToolCall=ChatCompletionMessageToolCall)
    lit.utils = stub("litellm.utils", Message=Message)
    sys.path.insert(0, "/root/reference")
    import pipelinerl.async_llm as ref_async
    import pipelinerl.llm as ref_llm
    return ref_async, ref_llm, ChatCompletionMessageToolCall, Function


def main():
    ref_async, ref_llm, ToolCall, Function = _import_reference()
    tok = tiny_chat_tokenizer()
    # the reference pins transformers 4.57, where apply_chat_template(tokenize=True) returns a plain list of ids; the
    # transformers installed here (5.x) returns a BatchEncoding unless told otherwise -> restore the pinned behaviour
    _orig = tok.apply_chat_template

    def _apply(*a, **k):
        if k.get("tokenize", True):
            k.setdefault("return_dict", False)
        return _orig(*a, **k)
    tok.apply_chat_template = _apply
    out = []
    for c in CASES:
        llm = types.SimpleNamespace(tokenizer=tok, chat_template_kwargs=c.get("chat_template_kwargs"), model_name="tiny")
        tcs = [ToolCall(id=t["id"], function=Function(name=t["name"], arguments=t["arguments"]))
               for t in c.get("tool_calls", [])] or None
        call = ref_llm.LLMCall(prompt=ref_llm.Prompt(messages=c["messages"], tools=c.get("tools")),
                               output=ref_llm.LLMOutput(content=c["content"], tool_calls=tcs),
                               prompt_length_tokens=c["prompt_len"], output_length_tokens=c["out_len"], cached=False,
                               llm_info={"finish_reason": c["finish_reason"]} if c["finish_reason"] else {},
                               logprobs=[ref_llm.TokenLogprob(logprob=-0.25 * (i + 1), token_id=t)
                                         for i, t in enumerate(c["gen"])])
        tt = ref_async.make_training_text(llm, call)
        out.append({"case": c, "expected": {k: getattr(tt, k) for k in
                                            ("text", "n_predicted", "input_ids", "labels", "logprobs", "finished",
                                             "prompt_tokens", "output_tokens")}})
        print(c["name"], "prompt ids", len(tt.input_ids) - len(c["gen"]), "finished", tt.finished, repr(tt.text[:60]))
    (OUT / "training_text_cases.json").write_text(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
