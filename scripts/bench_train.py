#!/usr/bin/env python
"""Training-step timing: FastVLA-1.5B, ALOHA-shaped batch, head-only training.

    python scripts/bench_train.py --batch 16 --steps 10 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 \
        scripts/bench_train.py --batch 16

One step on each rank = frozen backbone forward in the CUDA engine (no_grad, as the reference:
fastvlm_adapter.py:501) -> head forward / MSE / backward in libfvla (fvla_head_forward_backward: gradients written
straight into the flat buffer; `--autograd` runs the head through torch autograd instead) -> ONE all-reduce
of that buffer (NCCL) -> global-norm clip -> AdamW on the head.  The phases are timed separately with CUDA events
on the current stream (the collective is enqueued on it by torch.distributed's wait), max over ranks, and rank 0
prints one JSON line.
"""
from __future__ import annotations

import argparse
import contextlib
import json
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "vla-from-fastvlm_b200"))
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

STATE_DIM = ACTION_DIM = 14  # ALOHA
IMG_HW = (480, 640)


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="fastvlm-1.5b")
    ap.add_argument("--batch", type=int, default=16, help="per-GPU batch")
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--autograd", action="store_true", help="head forward/backward through torch autograd (A/B)")
    args = ap.parse_args()
    if not torch.cuda.is_available():
        raise SystemExit("bench_train.py needs a CUDA device: the FastVLA B200 path has no CPU fallback")
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    from vla_fastvlm.fastvla import FastVLAConfig, FastVLAPolicy
    from vla_fastvlm.training import HeadGradAllReduce, NativeHeadStep

    cfg = FastVLAConfig(vlm_model_name=f"synthetic:{args.model}", state_dim=STATE_DIM, action_dim=ACTION_DIM,
                        compute_dtype="bfloat16", image_token_mode="prefix")
    with contextlib.redirect_stdout(sys.stderr):
        policy = FastVLAPolicy(cfg).to(dev).train()
    g = torch.Generator().manual_seed(11 + rank)
    batch = {"images": torch.rand(args.batch, 3, *IMG_HW, generator=g).to(dev),
             "states": torch.randn(args.batch, STATE_DIM, generator=g).to(dev),
             "tasks": ["transfer the cube to the other arm"] * args.batch,
             "actions": torch.randn(args.batch, ACTION_DIM, generator=g).to(dev)}
    trainable = [p for p in policy.parameters() if p.requires_grad]
    reducer = HeadGradAllReduce(trainable)
    opt = torch.optim.AdamW(trainable, lr=1e-4, weight_decay=0.01, fused=True)
    native = None if args.autograd else NativeHeadStep(policy.model, reducer)

    names = ["backbone_fwd", "head_fwd_bwd", "all_reduce", "clip_adamw"]
    acc = {n: 0.0 for n in names}
    total_ms = 0.0
    first = last = None
    for it in range(args.warmup + args.steps):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        if native is None:
            reducer.zero()
        ev[0].record()
        with torch.no_grad():
            feats = policy.model.backbone(batch["images"], batch["tasks"], device=dev)
        ev[1].record()
        if native is not None:
            loss = native(feats, batch["states"], batch["actions"], train=True)
        else:
            m = policy.model
            fused = m.fusion(torch.cat([feats, m.state_projection(batch["states"])], dim=-1))
            loss = torch.nn.functional.mse_loss(m.action_head(fused), batch["actions"])
            loss.backward()
        ev[2].record()
        reducer.all_reduce()
        ev[3].record()
        reducer.clip_(1.0)
        opt.step()
        ev[4].record()
        torch.cuda.synchronize()
        if it == args.warmup:
            first = float(loss.detach())
        last = float(loss.detach())
        if it >= args.warmup:
            for i, n in enumerate(names):
                acc[n] += ev[i].elapsed_time(ev[i + 1])
            total_ms += ev[0].elapsed_time(ev[4])

    def mx(v: float) -> float:
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    step_ms = mx(total_ms / args.steps)
    phases = {n: mx(acc[n] / args.steps) for n in names}
    if rank == 0:
        print(json.dumps({
            "metric": "training step time (frozen backbone, head-only DP)", "value": step_ms, "unit": "ms/step",
            "higher_is_better": False, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "samples_per_s": args.batch * world / (step_ms / 1e3), "phases_ms": phases,
            "allreduce_bytes": reducer.numel * 4, "head_step": "autograd" if native is None else "libfvla fvla_head_forward_backward", "loss_first": first, "loss_last": last, "dtype": "bf16 backbone, fp32 head",
            "config": {"workload": f"FastVLA-{args.model.split('-')[-1]} train step, batch {args.batch}/GPU, ALOHA-shaped "
                                   "(480x640 frame letterboxed to 1024^2, 14-dim state/action)",
                       "parallelism": f"dp{world}: one all-reduce of the flat head gradient"}}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
