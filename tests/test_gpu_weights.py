"""Hot path (3) on one GPU: the copy kernel is byte-exact, lands in the INACTIVE buffer, and the sampler
flips at a step boundary while requests are in flight (no drain).  Cross-GPU / cross-process IPC is
covered by tests/test_gpu_multi.py (needs >= 2 GPUs)."""
import pytest
import torch

from tests.helpers import tiny_cfg, tiny_weights

pytestmark = pytest.mark.gpu


def test_push_is_byte_exact_and_targets_inactive_buffer(cuda_device):
    from pipelinerl_b200.model import ParamArena
    from pipelinerl_b200.weights import WeightReceiver, WeightUpdateManager
    cfg = tiny_cfg("gqa2")
    recv = WeightReceiver(cfg, cuda_device, n_pushers=2)
    learner = ParamArena(cfg, cuda_device).init_random(seed=1)
    # two "learner ranks" each push half of the bytes
    mgrs = [WeightUpdateManager([recv], learner.data, rank=r, n_learners=2) for r in range(2)]
    assert not recv.maybe_flip()
    for m in mgrs:
        m.send_weight_update(version=5)
    torch.cuda.synchronize()
    assert torch.equal(recv.arenas[1].data, learner.data)          # bit-for-bit
    assert torch.count_nonzero(recv.arenas[0].data) == 0           # the live buffer was not touched
    flipped = False
    for _ in range(50):
        flipped = recv.maybe_flip() or flipped
        torch.cuda.synchronize()
    assert flipped and recv.active == 1 and recv.version == 5 and recv.flips == 1
    # HF-name views of the pushed arena equal the learner's (fused-name mapping q/k/v -> qkv etc.)
    a, b = recv.arena.hf_state_dict(), learner.hf_state_dict()
    assert set(a) == set(b) and all(torch.equal(a[k], b[k]) for k in a)
    # the next update goes to buffer 0 (allowed only after the sampler acknowledged the first flip)
    torch.cuda.synchronize()
    assert int(recv.ctrl[2].item()) == 1
    learner.data.add_(1)
    for m in mgrs:
        m.send_weight_update(version=6)
    torch.cuda.synchronize()
    assert torch.equal(recv.arenas[0].data, learner.data) and not torch.equal(recv.arenas[1].data, learner.data)
    recv.close()


def test_second_update_waits_for_flip_ack(cuda_device):
    """Update 2 would overwrite the buffer the sampler is still reading until it has flipped to update 1."""
    from pipelinerl_b200.model import ParamArena
    from pipelinerl_b200.weights import WeightReceiver, WeightUpdateManager
    cfg = tiny_cfg("gqa2")
    recv = WeightReceiver(cfg, cuda_device, n_pushers=1)
    learner = ParamArena(cfg, cuda_device).init_random(seed=3)
    mgr = WeightUpdateManager([recv], learner.data)
    mgr.send_weight_update(1)
    with pytest.raises(TimeoutError):
        mgr.send_weight_update(2, ack_timeout_s=0.2)     # sampler has not flipped yet
    assert torch.count_nonzero(recv.arenas[0].data) == 0  # live buffer untouched
    for _ in range(50):
        if recv.maybe_flip():
            break
        torch.cuda.synchronize()
    torch.cuda.synchronize()
    mgr.send_weight_update(2, ack_timeout_s=5.0)         # now allowed
    torch.cuda.synchronize()
    assert torch.equal(recv.arenas[0].data, learner.data)
    recv.close()


def test_only_one_of_two_pushers_does_not_flip(cuda_device):
    from pipelinerl_b200.model import ParamArena
    from pipelinerl_b200.weights import WeightReceiver, WeightUpdateManager
    cfg = tiny_cfg("gqa2")
    recv = WeightReceiver(cfg, cuda_device, n_pushers=2)
    learner = ParamArena(cfg, cuda_device).init_random(seed=2)
    WeightUpdateManager([recv], learner.data, rank=0, n_learners=2).send_weight_update(1)
    torch.cuda.synchronize()
    for _ in range(20):
        assert not recv.maybe_flip()
        torch.cuda.synchronize()
    recv.close()


def test_in_flight_update_changes_logprobs_without_draining(cuda_device):
    """Start requests on weights A, push weights B mid-generation, flip between steps: sequences keep their KV
    and finish; tokens generated after the flip are scored by weights B (PipelineRL's defining behaviour,
    README.md:36 / vllm1.py:155-182)."""
    from oracle.decode_oracle import OracleQwen2
    from pipelinerl_b200.engine import DecodeEngine, SamplingParams
    from pipelinerl_b200.weights import WeightReceiver, WeightUpdateManager
    from pipelinerl_b200.model import ParamArena
    cfg = tiny_cfg("gqa2")
    wa, wb = tiny_weights(cfg, seed=42), tiny_weights(cfg, seed=43)
    recv = WeightReceiver(cfg, cuda_device, n_pushers=1)
    for name in recv.arena.names():
        recv.arenas[0].view(name).copy_(wa[name].to(torch.bfloat16))
    learner = ParamArena(cfg, cuda_device)
    for name in learner.names():
        learner.view(name).copy_(wb[name].to(torch.bfloat16))
    eng = DecodeEngine(cfg, recv.arena, max_batch=4, max_seq_len=128, max_new_tokens=32, device=cuda_device,
                       use_cuda_graph=True)
    eng.temperature, eng.greedy = 1.0, True
    prompt = list(range(10, 30))
    req = eng.add_request(prompt, SamplingParams(max_tokens=16, greedy=True))
    mgr = WeightUpdateManager([recv], learner.data)
    n_before = 6     # the prompt is prefilled inside the first step; 6 generated tokens under weights A
    for _ in range(n_before):
        eng.step()
    mgr.send_weight_update(version=1)
    while not recv.maybe_flip(eng):
        torch.cuda.synchronize()
    while eng.slot_req:
        eng.step()
        done = eng.harvest()
    out = done[0]
    assert out.finish_reason == "length" and len(out.output_ids) == 16
    # oracle: same schedule — first 6 generated tokens (and all KV so far) under A, the rest under B with the
    # mixed-version KV cache
    oa, ob = OracleQwen2(cfg, wa), OracleQwen2(cfg, wb)
    logits = oa.forward(torch.tensor(prompt))[-1]
    for i, (tok, lp) in enumerate(zip(out.output_ids, out.output_logprobs)):
        ref = torch.log_softmax(logits, -1)
        assert abs(lp - float(ref[tok])) <= 3e-2, (i, lp, float(ref[tok]))
        model = oa if i + 1 < 6 else ob
        if i + 1 == 6:  # hand the KV cache over: B continues on A's cache
            ob.k_cache, ob.v_cache = oa.k_cache, oa.v_cache
        logits = model.forward(torch.tensor([tok]))[-1]
    recv.close()
