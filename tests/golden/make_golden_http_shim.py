"""Interop recording for the HTTP shim: the REFERENCE's client (pipelinerl/async_llm.py:86-212, executed from
/root/reference) talks to pipelinerl_b200.http_shim.HttpShim (fake engine of tests/test_http_shim.py) over loopback.

    python tests/golden/make_golden_http_shim.py      (authoring container only)

Stored in tests/golden/http_shim_interop.json: the payload the reference sent, the shim's response, and the LLMCall
fields the reference parsed from it.  litellm / jsonref / accelerate / omegaconf are stubbed as in
make_golden_training_text.py (type aliases only on this path)."""
from __future__ import annotations

import asyncio
import json
import sys
import types
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(Path(__file__).resolve().parent))
from make_golden_training_text import _import_reference  # noqa: E402
from tests.helpers import tiny_chat_tokenizer  # noqa: E402
from tests.test_http_shim import FakeServer  # noqa: E402

OUT = Path(__file__).resolve().parent

CASES = [
    dict(name="plain", messages=[{"role": "user", "content": "what is the answer to life"}],
         parameters={"max_tokens": 8, "temperature": 1.0}),
    dict(name="length_and_kwargs", messages=[{"role": "system", "content": "guess a number"},
                                             {"role": "user", "content": "hello world"}],
         parameters={"max_tokens": 3, "temperature": 0.7}, chat_template_kwargs={"style": "terse"}),
    dict(name="tools_and_override", messages=[{"role": "user", "content": "hello"}],
         parameters={"max_tokens": 16, "temperature": 1.0}, max_tokens_override=2,
         tools=[{"type": "function", "function": {"name": "get_weather", "parameters": {"type": "object"}}}]),
]


async def main_async():
    import aiohttp
    from aiohttp import web
    from pipelinerl_b200.http_shim import HttpShim
    ref_async, ref_llm, _, _ = _import_reference()
    tok = tiny_chat_tokenizer()
    server = FakeServer().start()
    shim = HttpShim(server, tok, "tiny")
    captured = {}

    @web.middleware
    async def capture(request, handler):
        if request.method == "POST":
            captured["request"] = await request.json()
        resp = await handler(request)
        if request.method == "POST":
            captured["response"] = json.loads(resp.body)
        return resp
    shim.app.middlewares.append(capture)
    url = await shim.start()
    out = {"chat": []}
    try:
        async with aiohttp.ClientSession() as session:
            for c in CASES:
                llm = types.SimpleNamespace(base_url=url, model_name="tiny", api_token=None, collect_logprobs=True,
                                            parameters=c["parameters"], chat_template_kwargs=c.get("chat_template_kwargs"),
                                            load_tokenizer=lambda: None, tokenizer=tok)
                llm.log_output = lambda prompt, output, count_tokens=False: ref_llm.LLMCall(
                    prompt=prompt, output=output, cached=False, llm_info={})
                call = await ref_async.llm_async_generate(llm, ref_llm.Prompt(messages=c["messages"], tools=c.get("tools")),
                                                          session, max_tokens_override=c.get("max_tokens_override"))
                out["chat"].append({"name": c["name"], "request": captured["request"], "response": captured["response"],
                                    "parsed": {"content": call.output.content,
                                               "token_ids": [lp.token_id for lp in call.logprobs],
                                               "logprobs": [lp.logprob for lp in call.logprobs],
                                               "prompt_length_tokens": call.prompt_length_tokens,
                                               "output_length_tokens": call.output_length_tokens,
                                               "finish_reason": call.llm_info.get("finish_reason")}})
                print(c["name"], "->", call.output.content[:40].replace("\n", " "), call.llm_info.get("finish_reason"),
                      len(call.logprobs), "logprobs")
        # the reference-logprob pass: TrainableLLM.get_batch_logprobs_token_ids (llm.py:606-648) is synchronous
        # (requests.post) -> run it in a worker thread while this loop serves the shim
        prompts, completions = [[5, 6, 7], [9, 10, 11, 12]], [[8, 3], [13]]
        self_ns = types.SimpleNamespace(tokenizer=tok, load_tokenizer=lambda: None, api_token=None, model_name="tiny",
                                        base_url=url)
        scored = await asyncio.get_running_loop().run_in_executor(
            None, lambda: ref_llm.TrainableLLM.get_batch_logprobs_token_ids(self_ns, prompts, completions))
        out["score"] = {"prompt_token_ids": prompts, "completion_token_ids": completions,
                        "request": captured["request"], "response": captured["response"], "parsed": scored}
        print("score ->", [[(d["token_id"], round(d["logprob"], 3)) for d in s["content"]] for s in scored])
    finally:
        await shim.stop()
        server.stop()
    (OUT / "http_shim_interop.json").write_text(json.dumps(out, indent=1))


if __name__ == "__main__":
    asyncio.new_event_loop().run_until_complete(main_async())
