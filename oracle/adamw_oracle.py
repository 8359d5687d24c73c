"""Oracle for hot path (2c): AdamW + global-norm gradient clipping.

TEST INFRASTRUCTURE (see oracle/__init__.py).  numpy fp32, elementwise, in the
operation order of torch.optim.AdamW's single-tensor path, which is what
pipelinerl/finetune/optim.py:29 constructs; decay groups per optim.py:8-22;
clipping as torch.nn.utils.clip_grad_norm_ called at finetune_loop.py:739.
Pinned by tests/test_oracle_golden.py against tests/golden/adamw_case.npz
(three real torch.optim.AdamW steps).
"""
from __future__ import annotations

import numpy as np

NO_DECAY_MARKERS = ("bias", "LayerNorm.weight")


def no_decay(name: str) -> bool:
    return any(m in name for m in NO_DECAY_MARKERS)


def clip_coef(grads, max_norm):
    total = np.sqrt(sum(float(np.sum(g.astype(np.float64) ** 2)) for g in grads))
    if max_norm is None or max_norm <= 0:
        return total, 1.0
    return total, min(1.0, max_norm / (total + 1e-6))


def adamw_step(params, grads, exp_avg, exp_avg_sq, names, step, lr, weight_decay, beta1=0.9, beta2=0.999, eps=1e-8,
               max_grad_norm=None):
    """In-place on lists of fp32 numpy arrays; `step` is 1-based.  Returns the pre-clip grad norm."""
    f = np.float32
    norm, coef = clip_coef(grads, max_grad_norm)
    bc1 = 1.0 - beta1 ** step
    bc2 = 1.0 - beta2 ** step
    step_size = f(lr / bc1)
    bc2_sqrt = f(bc2 ** 0.5)
    for p, g, m, v, name in zip(params, grads, exp_avg, exp_avg_sq, names):
        g = (g * f(coef)).astype(f)
        wd = 0.0 if no_decay(name) else weight_decay
        p *= f(1.0 - lr * wd)
        m += (g - m) * f(1.0 - beta1)
        v *= f(beta2)
        v += f(1.0 - beta2) * g * g
        denom = np.sqrt(v) / bc2_sqrt + f(eps)
        p -= step_size * (m / denom)
    return norm
