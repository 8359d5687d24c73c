"""Fused AdamW (csrc/adamw.cu) vs torch.optim.AdamW outputs (golden) and vs the numpy oracle."""
import json

import numpy as np
import pytest
import torch

from oracle import adamw_oracle
from tests.helpers import GOLDEN

pytestmark = pytest.mark.gpu


def test_matches_torch_adamw_golden(cuda_device):
    from pipelinerl_b200.finetune.optim import FusedAdamW
    meta = json.loads((GOLDEN / "adamw_case.json").read_text())
    arrs = np.load(GOLDEN / "adamw_case.npz")
    names = meta["names"]
    params = [torch.nn.Parameter(torch.from_numpy(arrs[f"p0/{n}"]).to(cuda_device)) for n in names]
    opt = FusedAdamW(zip(names, params), lr=meta["lr"], weight_decay=meta["weight_decay"],
                     betas=tuple(meta["betas"]), eps=meta["eps"], max_grad_norm=meta["max_grad_norm"])
    for step in range(3):
        opt.zero_grad()
        for p, n in zip(params, names):
            p.grad.copy_(torch.from_numpy(arrs[f"g{step}/{n}"]))
        norm = opt.step()
        assert abs(norm.item() - arrs["grad_norms"][step]) <= 1e-5 * arrs["grad_norms"][step]
        for p, n in zip(params, names):
            np.testing.assert_allclose(p.detach().cpu().numpy(), arrs[f"p{step + 1}/{n}"], rtol=3e-6, atol=1e-8,
                                       err_msg=f"{n} step {step}")
    # bf16 shadow = round-to-nearest-even of the master
    for p, off in zip(params, opt.offsets):
        sh = opt.shadow_bf16[off:off + p.numel()].view(p.shape)
        assert torch.equal(sh, p.detach().to(torch.bfloat16))


@pytest.mark.parametrize("n_elems,bf16_grad", [(1, False), (4097, True), (3 * 1024 * 1024 + 5, True), (1 << 20, False)])
def test_matches_oracle_ragged_sizes(cuda_device, n_elems, bf16_grad):
    from pipelinerl_b200.finetune.optim import FusedAdamW
    g = torch.Generator().manual_seed(n_elems)
    # ragged tensor sizes incl. tiny biases; names decide the decay group
    sizes, left, i = [], n_elems, 0
    while left > 0:
        k = min(left, int(torch.randint(1, max(2, n_elems // 3 + 2), (1,), generator=g)))
        sizes.append(k)
        left -= k
    names = [("b%d.bias" % j) if j % 3 == 1 else ("w%d.weight" % j) for j in range(len(sizes))]
    p0 = [torch.randn(k, generator=g) * 0.05 for k in sizes]
    params = [torch.nn.Parameter(p.to(torch.bfloat16 if bf16_grad else torch.float32).to(cuda_device)) for p in p0]
    opt = FusedAdamW(zip(names, params), lr=3e-4, weight_decay=0.1, max_grad_norm=1.0, keep_lo_residual=True)
    o_p = [p.to(torch.bfloat16).float().numpy().copy() if bf16_grad else p.numpy().copy() for p in p0]
    o_m = [np.zeros_like(x) for x in o_p]
    o_v = [np.zeros_like(x) for x in o_p]
    for step in range(1, 4):
        grads = [torch.randn(k, generator=g) * (10.0 if step == 2 else 0.01) for k in sizes]
        if bf16_grad:
            grads = [x.to(torch.bfloat16) for x in grads]
        for p, x in zip(params, grads):
            p.grad.copy_(x)
        norm = opt.step()
        o_norm = adamw_oracle.adamw_step(o_p, [x.float().numpy() for x in grads], o_m, o_v, names, step, 3e-4, 0.1,
                                         max_grad_norm=1.0)
        assert abs(norm.item() - o_norm) <= 2e-5 * o_norm
    for off, k, want, wm, wv in zip(opt.offsets, sizes, o_p, o_m, o_v):
        np.testing.assert_allclose(opt.master[off:off + k].cpu().numpy(), want, rtol=1e-5, atol=1e-8)
        np.testing.assert_allclose(opt.exp_avg[off:off + k].cpu().numpy(), wm, rtol=1e-5, atol=1e-9)
        np.testing.assert_allclose(opt.exp_avg_sq[off:off + k].cpu().numpy(), wv, rtol=1e-5, atol=1e-12)
    hi = opt.shadow_bf16.float()
    lo = opt.shadow_lo.float()
    assert torch.equal(opt.shadow_bf16, opt.master.to(torch.bfloat16))
    # hi + lo reproduces the fp32 master to ~2^-16 relative (fp32-equivalent head)
    err = (hi + lo - opt.master).abs()
    assert (err <= opt.master.abs() * 2.0 ** -15 + 1e-30).all()


def test_no_clip_and_grad_scale(cuda_device):
    from pipelinerl_b200.finetune.optim import FusedAdamW
    p = torch.nn.Parameter(torch.ones(1000, device=cuda_device))
    opt = FusedAdamW([("w.weight", p)], lr=1e-2, weight_decay=0.0, max_grad_norm=None)
    p.grad.fill_(4.0)
    norm = opt.step(grad_scale=0.25)  # e.g. 1/accumulation passes
    assert abs(norm.item() - (1000 ** 0.5)) < 1e-3
    # first Adam step moves by lr regardless of scale
    assert torch.allclose(p.detach(), torch.full_like(p, 1.0 - 1e-2), rtol=1e-5)


def test_training_state_checkpoint_resume_is_bit_exact(cuda_device, tmp_path):
    """save_training_state -> fresh optimizer -> load_training_checkpoint: the resumed run takes bit-identical steps
    (reference: finetune/checkpoints.py:225-329; here three fp32 arenas in safetensors + a JSON)."""
    from pipelinerl_b200.finetune.checkpoints import load_training_checkpoint, save_training_state
    from pipelinerl_b200.finetune.optim import FusedAdamW, get_scheduler

    def build():
        torch.manual_seed(3)
        ps = [torch.nn.Parameter(torch.randn(s, device=cuda_device)) for s in ((37, 16), (16,), (5, 7))]
        names = ["a.weight", "a.bias", "b.weight"]
        opt = FusedAdamW(zip(names, ps), lr=1e-2, weight_decay=0.01, max_grad_norm=0.5)
        return ps, opt, get_scheduler("cosine", opt, 2, 10)

    def grads(ps, k):
        g = torch.Generator(device=cuda_device).manual_seed(100 + k)
        for p in ps:
            p.grad.copy_(torch.randn(p.shape, generator=g, device=cuda_device))
    ps, opt, sch = build()
    for k in range(3):
        grads(ps, k)
        opt.step()
        sch.step()
    save_training_state(tmp_path / "state", None, opt, sch, {"completed_steps": 3, "samples": 24})
    for k in range(3, 5):
        grads(ps, k)
        opt.step()
        sch.step()
    want = opt.master.clone()
    ps2, opt2, sch2 = build()
    extra = load_training_checkpoint(tmp_path / "state", None, opt2, sch2)
    assert extra == {"completed_steps": 3, "samples": 24} and opt2.step_count == 3 and sch2.last_step == 3
    assert opt2.param_groups[0]["lr"] == pytest.approx(sch.base_lrs[0] * sch2.factor(3))
    for k in range(3, 5):
        grads(ps2, k)
        opt2.step()
        sch2.step()
    torch.cuda.synchronize()
    assert torch.equal(opt2.master, want) and torch.equal(opt2.exp_avg, opt.exp_avg)
    assert torch.equal(opt2.shadow_bf16, opt.shadow_bf16)
