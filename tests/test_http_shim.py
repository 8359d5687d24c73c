"""HTTP shim (pipelinerl_b200/http_shim.py) over a FAKE engine: wire format of the reference's vLLM server.

tests/golden/http_shim_interop.json was recorded in the authoring container by pointing the REFERENCE's own client
(`pipelinerl.async_llm.llm_async_generate`, async_llm.py:86-212) at this shim (make_golden_http_shim.py): it holds the
request payload the reference sent, the response the shim gave and the LLMCall fields the reference parsed out of it.
Here the same payload is replayed: the response must be the recorded one (so what the reference accepted stays what the
shim produces), plus the scoring endpoint and the error paths."""
import asyncio
import json
import threading
import time

import pytest

from tests.helpers import GOLDEN, tiny_chat_tokenizer


class FakeRequest:
    def __init__(self, prompt_ids, params):
        n = min(params.max_tokens, 5)
        self.output_ids = [(sum(prompt_ids) + 7 * i) % 29 + 1 for i in range(n)]
        self.output_logprobs = [-0.5 - 0.125 * i for i in range(n)]
        self.finish_reason = "length" if n == params.max_tokens else "stop"
        self.model_version = 3


class FakeEngine:
    class arena:  # noqa: N801
        version = 3

    def score(self, seqs, temperature):
        return [[-(0.25 + 0.01 * t) for t in s[1:]] for s in seqs]


class FakeServer:
    """Stands in for serving.EngineServer: async generate + an engine thread that calls on_step_boundary."""

    def __init__(self):
        self.engine, self.on_step_boundary, self.error = FakeEngine(), None, None
        self._stop = threading.Event()
        self._thread = threading.Thread(target=self._run, daemon=True)

    def start(self):
        self._thread.start()
        return self

    def stop(self):
        self._stop.set()
        self._thread.join(timeout=5)

    def _run(self):
        while not self._stop.is_set():
            if self.on_step_boundary:
                self.on_step_boundary(self.engine)
            time.sleep(0.002)

    async def generate(self, prompt_ids, params):
        await asyncio.sleep(0)
        return FakeRequest(prompt_ids, params)


def _strip(resp: dict) -> dict:
    return {k: v for k, v in resp.items() if k not in ("id", "created")}


def _run(coro):
    return asyncio.new_event_loop().run_until_complete(coro)


def test_chat_completions_replays_what_the_reference_client_accepted():
    import aiohttp
    from pipelinerl_b200.http_shim import HttpShim
    rec = json.loads((GOLDEN / "http_shim_interop.json").read_text())

    async def go():
        server = FakeServer().start()
        shim = HttpShim(server, tiny_chat_tokenizer(), "tiny")
        url = await shim.start()
        try:
            async with aiohttp.ClientSession() as s:
                for case in rec["chat"]:
                    async with s.post(url + "/v1/chat/completions", json=case["request"]) as r:
                        assert r.status == 200
                        got = await r.json()
                    assert _strip(got) == _strip(case["response"]), case["name"]
                    # the fields the reference client reads (async_llm.py:173-207)
                    ch = got["choices"][0]
                    assert [int(x["token"].split(":")[-1]) for x in ch["logprobs"]["content"]] == case["parsed"]["token_ids"]
                    assert got["usage"]["prompt_tokens"] == case["parsed"]["prompt_length_tokens"]
                    assert got["usage"]["completion_tokens"] == case["parsed"]["output_length_tokens"]
                    assert ch["finish_reason"] == case["parsed"]["finish_reason"]
                # reference-logprob pass as TrainableLLM.get_batch_logprobs_token_ids sends and reads it (llm.py:606-648)
                sc = rec["score"]
                async with s.post(url + "/v1/completions", json=sc["request"]) as r:
                    assert r.status == 200
                    got = await r.json()
                assert _strip(got) == _strip(sc["response"])
                for i, comp in enumerate(sc["completion_token_ids"]):
                    tail = got["choices"][i]["prompt_logprobs"][-len(comp):]
                    parsed = [{**v, "generated": 0, "token_id": k} for lp in tail for k, v in lp.items()]
                    assert parsed == sc["parsed"][i]["content"]
                    assert [int(p["token_id"]) for p in parsed] == comp
                async with s.get(url + "/health") as r:
                    assert r.status == 200 and (await r.text()) == "OK"
                async with s.post(url + "/receive_weight_update", json={"version": 9}) as r:
                    assert (await r.json())["serving_version"] == 3
        finally:
            await shim.stop()
            server.stop()
    _run(go())


def test_scoring_endpoint_and_rejections():
    import aiohttp
    from pipelinerl_b200.http_shim import HttpShim

    async def go():
        server = FakeServer().start()
        shim = HttpShim(server, tiny_chat_tokenizer(), "tiny")
        url = await shim.start()
        try:
            async with aiohttp.ClientSession() as s:
                body = {"model": "tiny", "prompt": [[5, 6, 7, 8], [9, 10]], "temperature": 0.0, "max_tokens": 0,
                        "logprobs": 0, "echo": True, "n": 1, "stream": False}
                async with s.post(url + "/v1/completions", json=body) as r:
                    assert r.status == 200
                    out = await r.json()
                # what the reference reads (llm.py:639-648): choices[i].prompt_logprobs[-n:] = [{"<id>": {"logprob": ..}}]
                plp = out["choices"][0]["prompt_logprobs"]
                assert plp[0] is None and list(plp[1]) == ["6"] and abs(plp[1]["6"]["logprob"] + 0.31) < 1e-9
                assert [list(d)[0] for d in plp[1:]] == ["6", "7", "8"]
                assert list(out["choices"][1]["prompt_logprobs"][1]) == ["10"]
                for bad in ({"model": "tiny", "messages": [{"role": "user", "content": "hello"}], "top_p": 0.9},
                            {"model": "tiny", "messages": [{"role": "user", "content": "hello"}], "top_k": 20},
                            {"model": "tiny", "messages": []}):
                    async with s.post(url + "/v1/chat/completions", json=bad) as r:
                        assert r.status == 400 and "error" in await r.json()
                async with s.post(url + "/v1/completions", json={"prompt": [1, 2], "max_tokens": 4, "echo": True}) as r:
                    assert r.status == 400
        finally:
            await shim.stop()
            server.stop()
    _run(go())
