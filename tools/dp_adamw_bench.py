#!/usr/bin/env python
"""Learner-DP exchange step on a multi-GPU box: ShardedFusedAdamW (P2P reduce-scatter + AdamW shard + P2P all-gather)
vs the reference arrangement (NCCL all-reduce of the gradient arena, then a full fused AdamW on every rank).

    torchrun --nproc-per-node Ng tools/dp_adamw_bench.py [--params 7.616e9] [--check]
--check: small arena, 3 steps, compares against oracle/adamw_oracle.py on the summed gradients (test mode)."""
import argparse
import json
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from pipelinerl_b200.finetune.optim import FusedAdamW, ShardedFusedAdamW  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--params", type=float, default=7.616e9)
    ap.add_argument("--check", action="store_true")
    args = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    os.environ["NCCL_DEBUG"] = "WARN"
    dist.init_process_group("nccl", device_id=dev)
    if args.check:
        from oracle import adamw_oracle
        g = torch.Generator().manual_seed(7)
        shapes = {"a.weight": (300, 77), "b.bias": (77,), "c.weight": (1000, 129), "d.layernorm.weight": (129,), "e.weight": (5, 3)}
        p0 = {n: (torch.randn(s, generator=g) * 0.1).to(torch.bfloat16) for n, s in shapes.items()}
        params = [(n, torch.nn.Parameter(t.clone().to(dev))) for n, t in p0.items()]
        opt = ShardedFusedAdamW(params, lr=1e-3, weight_decay=0.01, max_grad_norm=0.3)
        o_p = [t.float().numpy().copy() for t in p0.values()]
        o_m = [np.zeros_like(x) for x in o_p]
        o_v = [np.zeros_like(x) for x in o_p]
        ok = True
        for step in range(1, 4):
            grads_all = [[(torch.randn(s, generator=torch.Generator().manual_seed(100 * step + 10 * r + i)) * 0.2).to(torch.bfloat16)
                          for i, s in enumerate(shapes.values())] for r in range(world)]
            for (n, p), gr in zip(params, grads_all[rank]):
                p.grad.copy_(gr)
            norm = opt.step().item()
            summed = [sum(grads_all[r][i].float() for r in range(world)).numpy() for i in range(len(shapes))]
            o_norm = adamw_oracle.adamw_step(o_p, summed, o_m, o_v, list(shapes), step, 1e-3, 0.01, max_grad_norm=0.3)
            ok &= abs(norm - o_norm) <= 1e-4 * o_norm
            for (n, p), want in zip(params, o_p):
                got = p.detach().float().cpu().numpy()
                ok &= bool(np.allclose(got, torch.from_numpy(want).to(torch.bfloat16).float().numpy(), rtol=0, atol=np.abs(want).max() * 2.0 ** -7))  # <= 1 bf16 ulp (fp32 ties)
        # every rank must hold bit-identical parameters
        h = torch.tensor([float(opt.shadow_bf16.view(torch.int16).to(torch.int64).sum())], device=dev, dtype=torch.float64)
        hs = [torch.zeros_like(h) for _ in range(world)]
        dist.all_gather(hs, h)
        ok &= len({float(x) for x in hs}) == 1
        flag = torch.tensor([1.0 if ok else 0.0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if rank == 0:
            print(json.dumps({"bench": "dp_adamw_check", "world": world, "ok": bool(flag.item() == 1.0)}), flush=True)
        dist.barrier()
        opt.close()
        dist.destroy_process_group()
        return

    n = int(args.params) // 4096 * 4096
    p = torch.nn.Parameter(torch.zeros(n, dtype=torch.bfloat16, device=dev))
    opt = ShardedFusedAdamW([("w.weight", p)], lr=1e-6, weight_decay=0.01, max_grad_norm=0.3)
    p.grad.fill_(1e-3)
    for _ in range(2):
        opt.step()
    t_red, t_upd = [], []
    for _ in range(3):
        opt.step()
        t_red.append(opt.last_phase_ms[0]); t_upd.append(opt.last_phase_ms[1])
    ours = torch.tensor([sum(t_red) / 3, sum(t_upd) / 3], device=dev, dtype=torch.float64)
    dist.all_reduce(ours, op=dist.ReduceOp.MAX)
    opt.close()
    del opt, p
    torch.cuda.empty_cache()
    # reference arrangement: NCCL all-reduce of the bf16 gradient arena + full AdamW on every rank
    base = None
    try:
        q = torch.nn.Parameter(torch.zeros(n, dtype=torch.bfloat16, device=dev))
        full = FusedAdamW([("w.weight", q)], lr=1e-6, weight_decay=0.01, max_grad_norm=0.3)
        q.grad.fill_(1e-3)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        for it in range(3):
            torch.cuda.synchronize(); dist.barrier()
            ev[0].record()
            dist.all_reduce(full.grad)
            ev[1].record()
            full.step()
            ev[2].record()
            torch.cuda.synchronize()
        b = torch.tensor([ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2])], device=dev, dtype=torch.float64)
        dist.all_reduce(b, op=dist.ReduceOp.MAX)
        base = b.tolist()
    except RuntimeError as e:
        base = str(e)[:200]
    if rank == 0:
        shard = n / world
        out = {"bench": "dp_adamw_exchange", "learners": world, "params": n,
               "ours": {"reduce_ms": round(ours[0].item(), 3), "update_ms": round(ours[1].item(), 3),
                        "total_ms": round(ours.sum().item(), 3),
                        "p2p_read_GB_per_rank": round(shard * 2 * (world - 1) / 1e9, 2),
                        "p2p_write_GB_per_rank": round(shard * 2 * (world - 1) / 1e9, 2),
                        "optimizer_state_GB_per_rank": round(shard * 12 / 1e9, 1)},
               "reference_arrangement": ({"nccl_allreduce_ms": round(base[0], 3), "full_adamw_ms": round(base[1], 3),
                                          "total_ms": round(base[0] + base[1], 3),
                                          "optimizer_state_GB_per_rank": round(n * 12 / 1e9, 1)}
                                         if isinstance(base, list) else base)}
        print(json.dumps(out), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
