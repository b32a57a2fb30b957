"""Gradient all-reduce for the data-parallel training step (BASELINE config 3).

What the reference does: `accelerator.prepare(model, ...)` wraps the policy in torch DDP and the gradient
all-reduce happens inside `accelerator.backward(loss)` (src/vla_fastvlm/training/trainer.py:68-78, :175).  The
backbone never receives gradients (`@torch.no_grad()` on `FastVLMBackbone.forward`, fastvlm_adapter.py:501;
`freeze_backbone=True`, :63), so the payload is the ~3 M-parameter head: 12.2 MB of fp32 at 0.5B, one DDP bucket.

What this module does instead, one process per GPU: the head parameters' `.grad` tensors are views into ONE
contiguous fp32 buffer (no flatten / copy before the collective), `loss.backward()` accumulates straight into it,
and the step issues exactly one `all_reduce(SUM)` over that buffer (NCCL over NVLink on the GPUs, gloo in the CPU
tests) followed by the 1/world scale — at 12 MB the collective is latency-bound (~60 us on NVSwitch), so there is
nothing to bucket or overlap.  No collective exists on the inference path.
"""
from __future__ import annotations

from typing import Dict, Iterable, List, Optional, Tuple

import torch
import torch.distributed as dist


class HeadGradAllReduce:
    def __init__(self, params: Iterable[torch.nn.Parameter], process_group: Optional[dist.ProcessGroup] = None):
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("no trainable parameters to all-reduce")
        dev = self.params[0].device
        if any(p.device != dev for p in self.params):
            raise ValueError("trainable parameters must live on one device")
        self.group = process_group
        self.numel = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(self.numel, device=dev, dtype=torch.float32)
        off = 0
        for p in self.params:  # gradients become views: autograd accumulates in place, nothing to pack later
            n = p.numel()
            p.grad = self.flat[off:off + n].view_as(p)
            off += n

    @property
    def world_size(self) -> int:
        return dist.get_world_size(self.group) if dist.is_available() and dist.is_initialized() else 1

    def zero(self) -> None:
        """Use instead of optimizer.zero_grad(): keeps the .grad views alive."""
        self.flat.zero_()

    def all_reduce(self) -> None:
        """Average the gradients over the data-parallel ranks (one collective over the flat buffer)."""
        for p in self.params:
            if p.grad is None or p.grad.data_ptr() < self.flat.data_ptr() or \
                    p.grad.data_ptr() >= self.flat.data_ptr() + self.flat.numel() * 4:
                raise RuntimeError("a head gradient no longer aliases the flat buffer "
                                   "(zero_grad(set_to_none=True)?): call HeadGradAllReduce.zero() instead")
        w = self.world_size
        if w > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
            self.flat.mul_(1.0 / w)

    def grad_norm(self) -> torch.Tensor:
        return self.flat.norm()

    def clip_(self, max_norm: float) -> torch.Tensor:
        """Global-norm clipping on the flat buffer (trainer.py:178 `clip_grad_norm_`)."""
        total = self.grad_norm()
        scale = torch.clamp(max_norm / (total + 1e-6), max=1.0)
        self.flat.mul_(scale)
        return total


def train_step(policy, batch: Dict[str, torch.Tensor], optimizer: torch.optim.Optimizer,
               reducer: HeadGradAllReduce, max_grad_norm: Optional[float] = 1.0) -> Tuple[float, float]:
    """One data-parallel step on this rank's shard of the batch: forward (frozen backbone in the engine, head in
    autograd) -> MSE -> backward into the flat buffer -> all-reduce -> clip -> optimizer.  Returns (loss, grad norm)
    of THIS rank before averaging / after averaging respectively."""
    policy.train()
    reducer.zero()
    loss, _ = policy.forward(batch)
    loss.backward()
    reducer.all_reduce()
    norm = reducer.clip_(max_grad_norm) if max_grad_norm is not None else reducer.grad_norm()
    optimizer.step()
    return float(loss.detach()), float(norm)
