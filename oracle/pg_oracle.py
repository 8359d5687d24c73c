"""Oracle for hot path (2): the logprob tail and the policy-gradient loss tail of rl_step.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Torch fp32 on CPU; gradients by
autograd, i.e. derived independently from the hand-written backward in
pipelinerl_b200/csrc/pg_loss.cu.

Follows pipelinerl/finetune/rl/__init__.py:
  logprob_tail   :207-233   logits/T -> gather - logsumexp ; exact entropy
  pg_tail        :237-365   ratios, clip, KL approx, token weights, masked sum
  stats          :388-439
  GSPO           :310-350 + rl/utils.py:106-208 (per_segment_sums)
Pinned by tests/test_oracle_golden.py against tests/golden/rl_step_*.npz, which
hold outputs of the reference itself.
"""
from __future__ import annotations

from dataclasses import dataclass

import torch


@dataclass
class OracleRLConfig:
    """Fields of RLConfig (rl/__init__.py:43-105) that the loss tail reads."""
    policy_loss: str = "ppo"
    use_advantages: bool = True
    epsilon_low: float = 0.2
    epsilon_high: float = 0.2
    batch_size: int = 0
    kl_coef: float = 0.1
    final_kl_coef: float = 0.1
    entropy_bonus: float = 0.0
    final_entropy_bonus: float = 0.0
    relu_log_p_weights: bool = False
    clamp_log_ratio_ref_new_value: float = 10
    overlong_filtering: bool = False
    group_normalization: bool = False
    temperature: float = 1.0

    @classmethod
    def from_dict(cls, d: dict) -> "OracleRLConfig":
        return cls(**{k: v for k, v in d.items() if k in cls.__dataclass_fields__})


def decayed(current_step: int, max_step: int, a: float, b: float) -> float:
    # rl/__init__.py:119-133
    return a + (b - a) * current_step / max_step


def logprob_tail(logits: torch.Tensor, input_ids: torch.Tensor, temperature: float):
    """logits [T, V] fp32, input_ids [T] -> (new_logprobs [T-1], entropy [T-1]).  (:207-233)"""
    z = logits[:-1] / temperature
    nxt = input_ids[1:].unsqueeze(1)
    picked = torch.gather(z, 1, nxt).squeeze(1)
    lse = torch.logsumexp(z, dim=-1)
    new_lp = picked - lse
    lp_all = z - lse.unsqueeze(1)
    entropy = -(torch.exp(lp_all) * lp_all).sum(-1)
    return new_lp, entropy


def pg_tail(new_lp: torch.Tensor, entropy: torch.Tensor, cols: dict, cfg: OracleRLConfig, current_step: int,
            max_step: int, sentinel: bool = False):
    """One packed row.  cols: unshifted 1-D tensors labels, rewards, advantages, ref_logprobs,
    old_logprobs, group_tokens, num_labels, overflow, position_ids, segment_ids(optional).
    Returns (loss 0-d tensor attached to new_lp/entropy, stats dict or {'input_size'} only)."""
    T = cols["labels"].shape[0]
    m = (cols["labels"][1:] != -100)
    mf = m.to(new_lp.dtype)
    sh = {k: cols[k][1:].to(torch.float32) for k in
          ("rewards", "advantages", "ref_logprobs", "old_logprobs", "group_tokens", "num_labels", "overflow")}
    pos = cols["position_ids"]
    starts = (pos == 0).clone()
    starts[0] = True
    n_seq = int(starts.sum())

    if cfg.group_normalization:
        w = 1.0 / sh["group_tokens"]
    else:
        w = torch.ones_like(sh["group_tokens"]) / cfg.batch_size
    if cfg.overlong_filtering:
        w = w * (1 - sh["overflow"])

    lr_no = new_lp - sh["old_logprobs"]
    ratio = torch.exp(lr_no)
    lr_rn = sh["ref_logprobs"] - new_lp
    cv = cfg.clamp_log_ratio_ref_new_value
    lr_rn_c = torch.clamp(lr_rn, -cv, cv)
    kl = torch.exp(lr_rn_c) - lr_rn_c - 1
    kl_no = torch.exp(lr_no) - lr_no - 1
    ent_coef = decayed(current_step, max_step, cfg.entropy_bonus, cfg.final_entropy_bonus)
    kl_coef = decayed(current_step, max_step, cfg.kl_coef, cfg.final_kl_coef)
    use_ent = cfg.entropy_bonus != 0.0 or cfg.final_entropy_bonus != 0.0
    ent_for_loss = entropy if use_ent else entropy.detach()

    adv = sh["advantages"]
    lpw = adv.detach() if cfg.use_advantages else sh["rewards"]
    if cfg.relu_log_p_weights:
        lpw = torch.clamp(lpw, min=0)

    if cfg.policy_loss == "ppo":
        rc = torch.clamp(ratio, 1 - cfg.epsilon_low, 1 + cfg.epsilon_high)
        ind = (rc != ratio).float()
        pol = torch.min(ratio * lpw, rc * lpw)
    elif cfg.policy_loss == "reinforce":
        ind = (ratio > 1 + cfg.epsilon_high).float()
        ratio = torch.clamp(ratio, 0, 1 + cfg.epsilon_high)
        pol = new_lp * lpw * ratio.detach()
    elif cfg.policy_loss == "gspo":
        seg = cols["segment_ids"][1:].long()
        n_seg = int(seg.max()) + 1 if seg.numel() else 0
        cnt = torch.zeros(n_seg).index_add_(0, seg[m], mf[m])
        lsum = torch.zeros(n_seg).index_add_(0, seg[m], lr_no[m])
        asum = torch.zeros(n_seg).index_add_(0, seg[m], adv[m])
        wsum = torch.zeros(n_seg).index_add_(0, seg[m], w[m])
        g_ratio = torch.exp(lsum / cnt.clamp(min=1e-6))
        g_adv = (asum / cnt.clamp(min=1e-6)).detach()
        valid = (cnt > 0) & (wsum > 0)
        g_rc = torch.clamp(g_ratio, 1 - cfg.epsilon_low, 1 + cfg.epsilon_high)
        seg_ind = ((g_rc != g_ratio) & valid).float()
        if sentinel or n_seg == 0:
            total = new_lp[:1].sum() * 0.0
        else:
            total = -(torch.min(g_ratio * g_adv, g_rc * g_adv) * valid.float() * wsum).sum()
        # per-token expansion of the per-segment indicator (:347-350): positional segments from
        # position_ids, applied to the shifted row
        bounds = torch.cat([torch.where(starts)[0], torch.tensor([T])])
        ind = torch.zeros(T - 1)
        for j in range(min(n_seq, n_seg)):
            ind[int(bounds[j]):int(bounds[j + 1])] = seg_ind[j]
    else:
        raise ValueError(f"Unknown algorithm {cfg.policy_loss}")

    if cfg.policy_loss != "gspo":
        tok = pol - kl_coef * kl
        if use_ent:
            tok = tok + ent_coef * ent_for_loss
        total = -((tok * w) * mf).nan_to_num(0).sum()

    n_lab = int(m.sum())
    if n_lab == 0:
        return total, {"input_size": float(T)}

    nl = sh["num_labels"]

    def ssum(x):
        return float(((x * mf).nan_to_num(0)).sum())

    ent_d = entropy.detach()
    new_d = new_lp.detach()
    ratio_d = ratio.detach()
    stats = {
        "loss": float(total.detach()), "max_loss": float(total.detach()), "min_loss": float(total.detach()),
        "reward": ssum(sh["rewards"] / nl), "max_reward": float(sh["rewards"][m].max()),
        "min_reward": float(sh["rewards"][m].min()),
        "entropy": ssum(ent_d / nl), "old_logprobs": ssum(sh["old_logprobs"] / nl),
        "new_logprobs": ssum(new_d / nl), "ref_logprobs": ssum(sh["ref_logprobs"] / nl),
        "advantage": ssum(adv / nl), "max_advantage": float(adv[m].max()), "min_advantage": float(adv[m].min()),
        "kl": ssum(kl.detach() / nl), "kl_new_old": ssum(kl_no.detach() / nl),
        "mean_abs_log_ratio_new_old": ssum(lr_no.detach().abs() / nl),
        "max_kl": float(kl.detach()[m].max()), "min_kl": float(kl.detach()[m].min()),
        "ratio_new_old": ssum(ratio_d / nl), "ratio_new_old_sum": ssum(ratio_d),
        "ratio_new_old_squared_sum": ssum(ratio_d * ratio_d),
        "ratio_ref_new": ssum(torch.exp(lr_rn.detach()) / nl),
        "ratio_ref_old": ssum(torch.exp(sh["ref_logprobs"] - sh["old_logprobs"]) / nl),
        "clamp_log_ratio_ref_new_indicator": ssum((lr_rn.detach().abs() > cv).float() / nl),
        "clamp_log_ratio_new_old_indicator": ssum(ind / nl),
        "token_weight": ssum(w / nl), "max_token_weight": float(w[m].max()), "min_token_weight": float(w[m].min()),
        "kl_coef": n_seq * kl_coef, "entropy_bonus_coef": n_seq * ent_coef,
        "num_output_tokens_sum": float(n_lab), "input_size": float(T),
    }
    return total, stats


def rl_step_oracle(logits: torch.Tensor, cols: dict, cfg: OracleRLConfig, current_step: int, max_step: int,
                   sentinel: bool = False):
    """Full tail on one packed row: logits [T,V] (leaf ok) -> loss, stats, new_lp, entropy."""
    new_lp, ent = logprob_tail(logits, cols["input_ids"], cfg.temperature)
    loss, stats = pg_tail(new_lp, ent, cols, cfg, current_step, max_step, sentinel)
    return loss, stats, new_lp, ent
