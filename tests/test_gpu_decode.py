"""Token step (hot path 1) through the engine / C ABI vs oracle/decode_oracle.py, HF fp32 goldens and vLLM bf16 goldens.

Tolerances.  Per-kernel parity (GEMM, attention, sampler) is tested at <= 1e-3 relative on identical inputs, and one
transformer layer at Qwen2.5-7B's real widths meets 1e-3 relative in the mean (max 1.4e-3).  END-TO-END logprobs of a
multi-layer bf16 model carry a floor no implementation pair escapes, because a 1-ulp difference in an fp32 sum flips the bf16
rounding of an activation: the oracle run incrementally vs in one pass differs by 0.9e-2..1.2e-2 max / < 2e-3 mean
(tests/test_oracle_golden.py::test_decode_oracle_vs_hf), and THE REFERENCE'S OWN ENGINE FAMILY (vLLM 0.22 bf16, same weights,
recorded on a B200: tests/golden/vllm_tiny_*.json) differs from this engine by 1.35e-2 / 2.35e-2 max, 4.8e-3 / 5.9e-3 mean
-- the same size as this engine's difference to the fp32 oracle (1.29e-2 / 1.79e-2 max).  Every end-to-end bound below is
1.5 x the value measured on a B200 (the tests print what they measure); greedy token ids must equal the oracle's wherever its
top-2 logit margin exceeds 5e-2."""
import numpy as np
import pytest
import torch

from oracle.decode_oracle import OracleQwen2
from tests.helpers import GOLDEN, tiny_cfg, tiny_weights

pytestmark = pytest.mark.gpu


# end-to-end bounds = 1.5 x the differences measured on a B200 (printed by the tests; profiles/r2_summary.md):
# measured max / mean |dlogprob| vs the oracle: gqa2 0.0129 / 0.0029, gqa7 0.0179 / 0.0046 (vs HF fp32: 0.0111 / 0.0030, 0.0153 / 0.0047)
E2E_BOUNDS = {"gqa2": (1.95e-2, 4.5e-3), "gqa7": (2.7e-2, 7.1e-3)}


def make_engine(cfg, weights, dev, **kw):
    from pipelinerl_b200.engine import DecodeEngine
    from pipelinerl_b200.model import ParamArena
    arena = ParamArena(cfg, dev)
    for name in arena.names():
        arena.view(name).copy_(weights[name].to(torch.bfloat16))
    return DecodeEngine(cfg, arena, device=dev, **kw)


@pytest.mark.parametrize("kind", ["gqa2", "gqa7"])
def test_teacher_forced_logprobs_match_oracle_and_hf(cuda_device, kind):
    """Feed a 150-token sequence as the prompt (prefill-by-decode), read the logits of every step."""
    cfg = tiny_cfg(kind)
    w = tiny_weights(cfg)
    eng = make_engine(cfg, w, cuda_device, max_batch=4, max_seq_len=256, max_new_tokens=8, use_cuda_graph=False,
                      prefill_chunk=0, fused_head=False)  # prefill-by-decode; materialised logits are inspected below
    gold = np.load(GOLDEN / f"qwen2_tiny_{kind}_T0.7.npz")
    tokens = gold["tokens"].tolist()
    from pipelinerl_b200.engine import SamplingParams
    eng.temperature, eng.greedy = 0.7, True
    eng.add_request(tokens, SamplingParams(max_tokens=2, temperature=0.7, greedy=True), model_version=0)
    # a second, shorter sequence in another slot exercises per-slot positions / block tables
    eng.add_request(tokens[:37], SamplingParams(max_tokens=2, temperature=0.7, greedy=True))
    got = []
    for t in range(len(tokens) - 1):
        eng.step()
        got.append(torch.log_softmax(eng.logits[0] / 0.7, -1)[tokens[t + 1]].item())
    got = np.array(got)
    orc = OracleQwen2(cfg, w)
    want = orc.score(tokens, 0.7).numpy()
    err = np.abs(got - want)
    err_hf = np.abs(got - gold["logprobs"])
    print(f"[decode e2e {kind}] vs oracle max {err.max():.4f} mean {err.mean():.5f} | vs HF fp32 max {err_hf.max():.4f} "
          f"mean {err_hf.mean():.5f}  (|logprob| ~ {np.abs(want).mean():.2f})")
    bmax, bmean = E2E_BOUNDS[kind]
    assert err.max() <= bmax and err.mean() <= bmean, (err.max(), err.mean(), int(err.argmax()))
    assert err_hf.max() <= bmax and err_hf.mean() <= bmean, (err_hf.max(), err_hf.mean())


@pytest.mark.parametrize("kind,use_graph,fused", [("gqa2", True, True), ("gqa7", False, True), ("gqa2", True, False)])
def test_greedy_generation_matches_oracle(cuda_device, kind, use_graph, fused):
    cfg = tiny_cfg(kind)
    w = tiny_weights(cfg)
    eng = make_engine(cfg, w, cuda_device, max_batch=8, max_seq_len=320, max_new_tokens=40, use_cuda_graph=use_graph,
                      fused_head=fused)
    from pipelinerl_b200.engine import SamplingParams
    g = torch.Generator().manual_seed(11)
    prompts = [torch.randint(0, cfg.vocab_size, (n,), generator=g).tolist() for n in (5, 64, 65, 130, 1, 17, 200, 33, 90)]
    outs = eng.generate(prompts, SamplingParams(max_tokens=24, temperature=1.0, greedy=True))
    orc = OracleQwen2(cfg, w)
    for pr, r in zip(prompts, outs):
        assert r.finish_reason == "length" and len(r.output_ids) == 24
        # replay the engine's own continuation through the oracle: logprobs must agree, and the oracle's
        # argmax must equal the engine's token wherever the oracle's margin is not a near-tie
        orc.reset()
        logits = orc.forward(torch.tensor(pr))[-1]
        for tok, lp in zip(r.output_ids, r.output_logprobs):
            ref_lp = torch.log_softmax(logits, -1)
            top2 = torch.topk(logits, 2).values
            if float(top2[0] - top2[1]) > 5e-2:
                assert int(torch.argmax(logits)) == tok
            assert abs(lp - float(ref_lp[tok])) <= E2E_BOUNDS[kind][0]
            logits = orc.forward(torch.tensor([tok]))[-1]


def test_sampling_distribution_and_logprob_capture(cuda_device):
    """Gumbel-max sampling draws from softmax(logits/T); the captured logprob is log_softmax at the drawn id."""
    from pipelinerl_b200 import _lib
    lib = _lib.load()
    B, V = 64, 1000
    g = torch.Generator().manual_seed(0)
    logits = (torch.randn(1, V, generator=g) * 2).repeat(B, 1).to(cuda_device)
    ids = torch.zeros(B, dtype=torch.int32, device=cuda_device)
    lps = torch.zeros(B, device=cuda_device)
    ws = torch.zeros(int(lib.prl_sample_workspace_bytes(B)), dtype=torch.uint8, device=cuda_device)
    counts = torch.zeros(V)
    T = 0.8
    ref = torch.log_softmax(logits[0].cpu() / T, -1)
    for step in range(400):
        _lib.check(lib.prl_sample_logprob(logits.data_ptr(), B, V, T, 0, 1234, step, ids.data_ptr(), lps.data_ptr(), ws.data_ptr(), ws.numel(), None))
        i = ids.cpu().long()
        assert torch.allclose(lps.cpu(), ref[i], atol=1e-4)
        counts += torch.bincount(i, minlength=V).float()
    n = counts.sum()
    p = ref.exp()
    top = torch.topk(p, 20).indices
    # 25 600 draws: the 20 most likely ids are within 5 sigma of their expectation
    sigma = torch.sqrt(n * p[top] * (1 - p[top]))
    assert ((counts[top] - n * p[top]).abs() < 5 * sigma + 1).all()
    # greedy == argmax, ties to the lowest index
    _lib.check(lib.prl_sample_logprob(logits.data_ptr(), B, V, 1.0, 1, 0, 0, ids.data_ptr(), lps.data_ptr(), ws.data_ptr(), ws.numel(), None))
    assert (ids.cpu() == int(torch.argmax(logits[0].cpu()))).all()


def test_paged_attention_long_context_vs_fp32(cuda_device):
    """Decode attention alone at BASELINE-like context (8192+ tokens, 7:1 GQA, ragged lengths, scattered pages)."""
    from pipelinerl_b200 import _lib
    lib = _lib.load()
    dev = cuda_device
    B, n_q, n_kv, D, P = 6, 28, 4, 128, 64
    lens = [8192, 8191, 1, 65, 12000, 640]
    max_blocks = 192
    n_pages = 1 + sum((l + P - 1) // P for l in lens) + 5
    g = torch.Generator().manual_seed(5)
    L, layer = 2, 1
    kv = (torch.randn(L * 2 * n_pages * n_kv * P * D, generator=g) * 0.5).to(torch.bfloat16).to(dev)
    kv5 = kv.view(L, 2, n_pages, n_kv, P, D)
    perm = torch.randperm(n_pages - 1, generator=g) + 1
    bt = torch.zeros(B, max_blocks, dtype=torch.int32)
    at = 0
    for b, l in enumerate(lens):
        k = (l + P - 1) // P
        bt[b, :k] = perm[at:at + k].int()
        at += k
    q = (torch.randn(B, n_q, D, generator=g)).to(torch.bfloat16).to(dev)
    bt_d, sl_d = bt.to(dev), torch.tensor(lens, dtype=torch.int32, device=dev)
    out = torch.zeros(B, n_q * D, dtype=torch.bfloat16, device=dev)
    for splits in (1, 4, int(lib.prl_paged_attn_splits(B, n_kv, max(lens)))):
        ws = torch.zeros(int(lib.prl_paged_attn_workspace_bytes(B, n_q, splits)), dtype=torch.uint8, device=dev)
        _lib.check(lib.prl_paged_attn_decode(q.data_ptr(), kv.data_ptr(), n_pages, L, layer, bt_d.data_ptr(), max_blocks,
                                             sl_d.data_ptr(), B, n_q, n_kv, D, P, splits, 1.0 / D ** 0.5, out.data_ptr(),
                                             ws.data_ptr(), ws.numel(), None))
        torch.cuda.synchronize()
        for b, l in enumerate(lens):
            k = (l + P - 1) // P
            pages = bt[b, :k].long().to(dev)
            K = kv5[layer, 0, pages].permute(1, 0, 2, 3).reshape(n_kv, k * P, D)[:, :l].float()
            V = kv5[layer, 1, pages].permute(1, 0, 2, 3).reshape(n_kv, k * P, D)[:, :l].float()
            qb = q[b].float().view(n_kv, n_q // n_kv, D)
            s = torch.einsum("grd,gsd->grs", qb, K) / D ** 0.5
            ref = torch.einsum("grs,gsd->grd", torch.softmax(s, -1), V).reshape(-1)
            err = (out[b].float() - ref).abs().max().item()
            assert err <= 4e-3 * max(1.0, ref.abs().max().item()), (splits, b, err)


def test_prefix_sharing_and_chunked_prefill(cuda_device):
    """8 attempts of one 150-token prompt (a GRPO group): one prefill, 7 prefix hits; results identical to an engine
    without sharing and to the oracle.  prefill_chunk=64 forces multi-chunk prefill with a ragged tail."""
    from pipelinerl_b200.engine import SamplingParams
    cfg = tiny_cfg("gqa7")
    w = tiny_weights(cfg)
    g = torch.Generator().manual_seed(21)
    prompt = torch.randint(0, cfg.vocab_size, (150,), generator=g).tolist()
    other = torch.randint(0, cfg.vocab_size, (70,), generator=g).tolist()
    prompts = [prompt] * 8 + [other]
    outs = {}
    for share in (True, False):
        eng = make_engine(cfg, w, cuda_device, max_batch=12, max_seq_len=256, max_new_tokens=16, prefill_chunk=64,
                          prefix_sharing=share)
        res = eng.generate(prompts, SamplingParams(max_tokens=12, greedy=True))
        outs[share] = [(r.output_ids, r.output_logprobs) for r in res]
        if share:
            # page-granular sharing: the 2 full pages (128 tokens) of the 149-token prefix are reused, the 21-token
            # tail of each attempt is re-prefilled
            assert eng.stats["prefix_hits"] == 7 and eng.stats["prefix_hit_tokens"] == 7 * 128
            assert eng.stats["prefill_tokens"] == 149 + 7 * 21 + 69
        else:
            assert eng.stats["prefix_hits"] == 0 and eng.stats["prefill_tokens"] == 8 * 149 + 69
        # all pages come back (prefix cache entries are evictable)
        eng._evict_prefixes(10 ** 9)
        assert len(eng.free_pages) == eng.n_pages - 1 and not any(eng.page_ref[1:])
    for (ids_a, lp_a), (ids_b, lp_b) in zip(outs[True], outs[False]):
        assert ids_a == ids_b and np.allclose(lp_a, lp_b, atol=1e-5)
    assert all(o[0] == outs[True][0][0] for o in outs[True][:8])
    orc = OracleQwen2(cfg, w)
    logits = orc.forward(torch.tensor(prompt))[-1]
    for tok, lp in zip(*outs[True][0]):
        ref = torch.log_softmax(logits, -1)
        assert abs(lp - float(ref[tok])) <= 3e-2
        logits = orc.forward(torch.tensor([tok]))[-1]


@pytest.mark.parametrize("kernel", ["tc", "tc2", "mma"])
@pytest.mark.parametrize("n_q,n_kv,seqs", [
    (28, 4, [(7000, 1000), (0, 37)]),              # Qwen2.5-7B grouping (R = 7 -> 18 tokens x 7 heads per UMMA tile)
    (4, 2, [(0, 300), (129, 70), (64, 1)]),        # R = 2; chunk starting mid-page; single-row chunk
    (8, 8, [(500, 129), (0, 128)]),                # R = 1 (MHA): 128 tokens per tile
    (8, 1, [(1000, 260)]),                         # R = 8
])
def test_prefill_attention_long_context_vs_fp32(cuda_device, kernel, n_q, n_kv, seqs):
    """Prefill attention alone (both kernels: tcgen05 `tc`, mma.sync `mma`): chunks of queries at arbitrary positions
    of longer sequences (causal), several sequences packed in one launch."""
    from pipelinerl_b200 import _lib
    lib = _lib.load()
    dev = cuda_device
    D, P = 128, 64
    max_blocks = 128
    n_pages = 1 + sum((p0 + ql + P - 1) // P for p0, ql in seqs) + 3
    g = torch.Generator().manual_seed(9)
    L, layer = 1, 0
    kv = (torch.randn(L * 2 * n_pages * n_kv * P * D, generator=g) * 0.5).to(torch.bfloat16).to(dev)
    kv5 = kv.view(L, 2, n_pages, n_kv, P, D)
    perm = torch.randperm(n_pages - 1, generator=g) + 1
    bt = torch.zeros(len(seqs), max_blocks, dtype=torch.int32)
    at = 0
    for z, (p0, ql) in enumerate(seqs):
        k = (p0 + ql + P - 1) // P
        bt[z, :k] = perm[at:at + k].int()
        at += k
    rows = sum(ql for _, ql in seqs)
    q = torch.randn(rows, n_q, D, generator=g).to(torch.bfloat16).to(dev)
    out = torch.zeros(rows, n_q * D, dtype=torch.bfloat16, device=dev)
    starts = [sum(s[1] for s in seqs[:z]) for z in range(len(seqs))]
    i32 = lambda v: torch.tensor(v, dtype=torch.int32, device=dev)
    qs, ql_t, p0_t, sl = i32(starts), i32([s[1] for s in seqs]), i32([s[0] for s in seqs]), i32(list(range(len(seqs))))
    bt_d = bt.to(dev)
    if kernel in ("tc", "tc2"):      # tc2: generation 2 (ping-pong softmax groups, P and O in TMEM) on the paged path
        _lib.check(lib.prl_attn_set_prefill_generation(2 if kernel == "tc2" else 1))
        _lib.check(lib.prl_paged_attn_prefill_tc(q.data_ptr(), rows, kv.data_ptr(), n_pages, L, layer, bt_d.data_ptr(),
                                                 max_blocks, qs.data_ptr(), ql_t.data_ptr(), p0_t.data_ptr(), sl.data_ptr(),
                                                 len(seqs), max(s[1] for s in seqs), n_q, n_kv, D, P, 1.0 / D ** 0.5,
                                                 out.data_ptr(), None))
    else:
        _lib.check(lib.prl_paged_attn_prefill(q.data_ptr(), kv.data_ptr(), n_pages, L, layer, bt_d.data_ptr(), max_blocks,
                                              qs.data_ptr(), ql_t.data_ptr(), p0_t.data_ptr(), sl.data_ptr(), len(seqs),
                                              max(s[1] for s in seqs), n_q, n_kv, D, P, 1.0 / D ** 0.5, out.data_ptr(),
                                              None))
    torch.cuda.synchronize()
    _lib.check(lib.prl_attn_set_prefill_generation(2))      # back to the default
    for z, (p0, ql) in enumerate(seqs):
        S = p0 + ql
        k = (S + P - 1) // P
        pages = bt[z, :k].long().to(dev)
        K = kv5[layer, 0, pages].permute(1, 0, 2, 3).reshape(n_kv, k * P, D)[:, :S].float()
        V = kv5[layer, 1, pages].permute(1, 0, 2, 3).reshape(n_kv, k * P, D)[:, :S].float()
        qz = q[starts[z]:starts[z] + ql].float().view(ql, n_kv, n_q // n_kv, D)
        s = torch.einsum("tgrd,gsd->gtrs", qz, K) / D ** 0.5
        mask = torch.arange(S, device=dev)[None, :] > (p0 + torch.arange(ql, device=dev))[:, None]
        s = s.masked_fill(mask[None, :, None, :], float("-inf"))
        ref = torch.einsum("gtrs,gsd->tgrd", torch.softmax(s, -1), V).reshape(ql, -1)
        err = (out[starts[z]:starts[z] + ql].float() - ref).abs().max().item()
        assert err <= 4e-3 * max(1.0, ref.abs().max().item()), (z, err)


def test_kv_reuse_across_turns(cuda_device):
    """Multi-turn rollouts (BASELINE config 5): turn 2's prompt = turn 1's prompt + answer + new message; every
    full page of the common prefix is reused from the cache even after turn 1 finished."""
    from pipelinerl_b200.engine import SamplingParams
    cfg = tiny_cfg("gqa2")
    w = tiny_weights(cfg)
    eng = make_engine(cfg, w, cuda_device, max_batch=4, max_seq_len=512, max_new_tokens=16, prefill_chunk=128)
    g = torch.Generator().manual_seed(33)
    turn1 = torch.randint(0, cfg.vocab_size, (200,), generator=g).tolist()
    r1 = eng.generate([turn1], SamplingParams(max_tokens=8, greedy=True))[0]
    turn2 = turn1 + r1.output_ids + torch.randint(0, cfg.vocab_size, (90,), generator=g).tolist()
    before = dict(eng.stats)
    r2 = eng.generate([turn2], SamplingParams(max_tokens=8, greedy=True))[0]
    assert eng.stats["prefix_hit_tokens"] - before["prefix_hit_tokens"] == (199 // 64) * 64   # 3 pages of turn 1's prompt
    assert eng.stats["prefill_tokens"] - before["prefill_tokens"] == len(turn2) - 1 - 192
    orc = OracleQwen2(cfg, w)
    logits = orc.forward(torch.tensor(turn2))[-1]
    for tok, lp in zip(r2.output_ids, r2.output_logprobs):
        assert abs(lp - float(torch.log_softmax(logits, -1)[tok])) <= 3e-2
        logits = orc.forward(torch.tensor([tok]))[-1]


@pytest.mark.parametrize("kind", ["gqa2", "gqa7"])
def test_score_reference_logprobs(cuda_device, kind):
    """engine.score() (chunked prefill + fused head with targets) == oracle / HF teacher-forced logprobs."""
    cfg = tiny_cfg(kind)
    w = tiny_weights(cfg)
    eng = make_engine(cfg, w, cuda_device, max_batch=4, max_seq_len=256, max_new_tokens=8, prefill_chunk=64)
    gold = np.load(GOLDEN / f"qwen2_tiny_{kind}_T0.7.npz")
    tokens = gold["tokens"].tolist()
    got = np.array(eng.score([tokens, tokens[:3], [5]], temperature=0.7)[0])
    orc = OracleQwen2(cfg, w)
    want = orc.score(tokens, 0.7).numpy()
    assert got.shape == want.shape
    err = np.abs(got - want)
    assert err.max() <= 3e-2 and err.mean() <= 6e-3, (err.max(), err.mean())
    assert np.abs(got - gold["logprobs"]).max() <= 3e-2
    assert len(eng.free_pages) == eng.n_pages - 1 and len(eng.free_slots) == eng.B


def test_per_request_sampling_parameters_share_one_batch(cuda_device):
    """Requests admitted with different temperature / greedy settings decode in the same batch and each one's sampled
    logprob is log_softmax(logits / ITS temperature) at ITS sampled id (sampling params were engine-global)."""
    from pipelinerl_b200.engine import SamplingParams
    cfg = tiny_cfg("gqa2")
    w = tiny_weights(cfg)
    eng = make_engine(cfg, w, cuda_device, max_batch=4, max_seq_len=128, max_new_tokens=16, use_cuda_graph=False,
                      prefill_chunk=0, fused_head=False)
    g = torch.Generator().manual_seed(3)
    prompt = torch.randint(0, cfg.vocab_size, (9,), generator=g).tolist()
    params = [SamplingParams(max_tokens=8, temperature=1.0, greedy=True), SamplingParams(max_tokens=8, temperature=0.5),
              SamplingParams(max_tokens=8, temperature=2.0), SamplingParams(max_tokens=8, temperature=1.0)]
    reqs = [eng.add_request(prompt, p) for p in params]
    checked = 0
    for _ in range(len(prompt) + 8):      # 9 prompt tokens through the decode path, then 8 generated tokens
        eng.step()
        logits, ids, lps = eng.logits.clone(), eng.sampled.clone(), eng.sampled_lp.clone()
        for r, p in zip(reqs, params):
            s = r.slot
            T = 1.0 if p.greedy else p.temperature
            ref = torch.log_softmax(logits[s] / T, -1)
            assert abs(float(lps[s]) - float(ref[int(ids[s])])) <= 2e-4, (s, T)
            if p.greedy:
                assert int(ids[s]) == int(torch.argmax(logits[s]))
            checked += 1
    assert checked >= 40
    # the three sampled slots see the same logits at the first generated position but draw with different temperatures:
    # their logprobs of one and the same token differ by the temperature, not by noise
    done = {r.req_id: r for r in eng.harvest()}
    assert len(done) == 4 and all(len(r.output_ids) == 8 for r in done.values())
    with pytest.raises(ValueError):
        eng.add_request([0, cfg.vocab_size], SamplingParams(max_tokens=2))      # out-of-range token id
    with pytest.raises(ValueError):
        eng.add_request([1, 2], SamplingParams(max_tokens=2, temperature=0.0))   # T = 0 must be spelt greedy=True


@pytest.mark.parametrize("kind", ["gqa2", "gqa7"])
def test_engine_matches_vllm_golden(cuda_device, kind):
    """Row a1 against the reference's actual sampler engine family: vLLM bf16 on the SAME weights
    (tests/golden/vllm_tiny_<kind>.json, recorded on a B200 by tests/golden/make_golden_vllm.py).  Both engines are
    bf16 with fp32 accumulation and differ in summation order only; bound = 1.5 x the difference measured when the
    golden was recorded (printed below)."""
    import json
    f = GOLDEN / f"vllm_tiny_{kind}.json"
    if not f.exists():
        pytest.skip("vLLM golden not recorded yet (needs a GPU box: tests/golden/make_golden_vllm.py)")
    gold = json.loads(f.read_text())
    cfg = tiny_cfg(kind)
    w = tiny_weights(cfg)
    eng = make_engine(cfg, w, cuda_device, max_batch=4, max_seq_len=512, max_new_tokens=8)
    tf = gold["teacher_forced"]
    got = np.array(eng.score([tf["tokens"]], temperature=1.0)[0])
    err = np.abs(got - np.array(tf["logprobs"]))
    worst, mean = [float(err.max())], [float(err.mean())]
    for pr, gen in zip(gold["prompts"], gold["greedy"]):
        seq = pr + gen["ids"]
        lp = np.array(eng.score([seq], temperature=1.0)[0])[len(pr) - 1:]
        e = np.abs(lp - np.array(gen["logprobs"]))
        worst.append(float(e.max()))
        mean.append(float(e.mean()))
    print(f"[vllm golden {kind}] max |dlogprob| {max(worst):.4f}  mean {np.mean(mean):.5f}  ({gold['engine']})")
    bound_max, bound_mean = VLLM_BOUNDS[kind]
    assert max(worst) <= bound_max and np.mean(mean) <= bound_mean, (worst, mean)


# 1.5 x the measured difference to vLLM 0.22 bf16 on the same weights (two bf16 engines, different summation orders):
# measured max / mean |dlogprob| gqa2 0.0135 / 0.0048, gqa7 0.0235 / 0.0059 on |logprob| ~ 7 -- the same size as the difference
# to the fp32 oracle, i.e. this IS the bf16 floor between the reference's sampler and any other correct engine
VLLM_BOUNDS = {"gqa2": (2.05e-2, 7.2e-3), "gqa7": (3.55e-2, 8.9e-3)}


def test_one_layer_qwen7b_width_decode_step_vs_oracle(cuda_device):
    """One transformer layer at Qwen2.5-7B's real widths (H 3584, I 18944, 28 q / 4 kv heads, qkv bias) through the
    decode step (tcgen05 split-K GEMMs, RoPE + KV write, paged attention, residual RMSNorm, SiLU, fp32-equivalent head)
    against the oracle on identical bf16-valued weights and the same rounding points: with a single layer there is no
    chain of bf16 re-roundings to amplify summation-order noise.  Measured on a B200: mean relative |dlogprob| 4.7e-4, max
    1.39e-3 (one position of 47) -- the mean meets the north star's 1e-3, the max is bounded at 1.5 x measured."""
    from dataclasses import replace
    from pipelinerl_b200.engine import SamplingParams
    from pipelinerl_b200.model import ModelConfig
    cfg = replace(ModelConfig.qwen2_5_7b(), num_layers=1, vocab_size=4096)
    w = tiny_weights(cfg, std=0.02, bias_std=0.1)
    eng = make_engine(cfg, w, cuda_device, max_batch=4, max_seq_len=128, max_new_tokens=4, use_cuda_graph=False,
                      prefill_chunk=0, fused_head=False)
    g = torch.Generator().manual_seed(5)
    tokens = torch.randint(0, cfg.vocab_size, (48,), generator=g).tolist()
    eng.add_request(tokens, SamplingParams(max_tokens=2, temperature=1.0, greedy=True))
    got = []
    for t in range(len(tokens) - 1):
        eng.step()
        got.append(torch.log_softmax(eng.logits[0], -1)[tokens[t + 1]].item())
    got = np.array(got)
    want = OracleQwen2(cfg, w).score(tokens, 1.0).numpy()
    rel = np.abs(got - want) / np.abs(want)
    print(f"[7B-width one layer] max rel |dlogprob| {rel.max():.2e}  mean {rel.mean():.2e}  (|logprob| ~ {np.abs(want).mean():.2f})")
    assert rel.mean() <= 1e-3 and rel.max() <= 2.1e-3, (rel.max(), rel.mean(), int(rel.argmax()))


def test_token_step_with_and_without_the_swiglu_epilogue_is_bit_identical(cuda_device):
    """The token step's gate_up GEMM carries SiLU(gate) * up in its epilogue when it runs without split-K
    (prl_gemm_swiglu_decode; Qwen2.5-7B widths do).  Same one-layer 7B-width model, same tokens: the logits of every step are
    bitwise those of the GEMM + prl_silu_mul pair (engine.fuse_swiglu = False)."""
    from dataclasses import replace
    from pipelinerl_b200.engine import SamplingParams
    from pipelinerl_b200.model import ModelConfig
    cfg = replace(ModelConfig.qwen2_5_7b(), num_layers=1, vocab_size=4096)
    w = tiny_weights(cfg, std=0.02, bias_std=0.1)
    g = torch.Generator().manual_seed(6)
    tokens = torch.randint(0, cfg.vocab_size, (12,), generator=g).tolist()
    logits = {}
    for fuse in (True, False):
        eng = make_engine(cfg, w, cuda_device, max_batch=4, max_seq_len=128, max_new_tokens=4, use_cuda_graph=False,
                          prefill_chunk=0, fused_head=False)
        eng.fuse_swiglu = fuse
        assert eng.split_k["gate_up"] == 1     # otherwise the fused path is not taken and the test proves nothing
        eng.add_request(tokens, SamplingParams(max_tokens=2, temperature=1.0, greedy=True))
        rows = []
        for _ in range(len(tokens) - 1):
            eng.step()
            rows.append(eng.logits[0].clone())
        logits[fuse] = torch.stack(rows)
        del eng
    assert torch.equal(logits[True], logits[False])
