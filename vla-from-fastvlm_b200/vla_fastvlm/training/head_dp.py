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


class NativeHeadStep:
    """Forward + MSE + backward of the action head through libfvla (`fvla_head_forward_backward`, csrc/head_train.cu),
    in place of autograd: the kernels read the torch parameters directly and WRITE d loss / d param into the reducer's
    flat buffer — the buffer the NCCL all-reduce runs over — so there is no autograd graph, no per-parameter `.grad`
    allocation and no flatten copy between backward and the collective.  Train-mode semantics of the reference head
    (fastvla/fastvlm_with_expert.py:23-38): Dropout(p) behind the first SiLU of `fusion`, with the keep mask drawn here."""

    def __init__(self, head_owner: torch.nn.Module, reducer: HeadGradAllReduce) -> None:
        from .. import _native as N

        self.N = N
        self.lib = N.load()
        self.reducer = reducer
        mods = (head_owner.state_projection, head_owner.fusion, head_owner.action_head)
        self.params = [p for m in mods for p in m.parameters()]
        if len(self.params) != 12:
            raise ValueError(f"expected the 12 tensors of the FastVLA head, found {len(self.params)}")
        if [id(p) for p in self.params] != [id(p) for p in reducer.params]:
            raise ValueError("the reducer's flat buffer must hold exactly the head parameters, in module order")
        for p in self.params:
            if p.dtype != torch.float32 or not p.is_cuda or not p.is_contiguous():
                raise ValueError("head parameters must be contiguous fp32 CUDA tensors")
        self.S = head_owner.state_projection[1].in_features
        self.Hd = head_owner.state_projection[1].out_features
        self.F = head_owner.fusion[0].out_features
        self.H = head_owner.fusion[0].in_features - self.Hd
        self.A = head_owner.action_head.out_features
        self.drop_p = float(head_owner.fusion[3].p)
        self._scratch: Optional[torch.Tensor] = None
        self.loss = torch.zeros((), device=self.params[0].device, dtype=torch.float32)

    def __call__(self, pooled: torch.Tensor, states: torch.Tensor, target: torch.Tensor, train: bool = True,
                 keep_mask: Optional[torch.Tensor] = None) -> torch.Tensor:
        """pooled (B,H), states (B,S), target (B,A): fp32 CUDA.  Returns the loss (device scalar); the gradients are in
        `reducer.flat`.  `keep_mask` (B,F) uint8 overrides the internally drawn Dropout mask (tests)."""
        import ctypes as C

        N = self.N
        B = pooled.shape[0]
        dev = pooled.device
        pooled, states, target = (t.to(dev, torch.float32).contiguous() for t in (pooled, states, target))
        if pooled.shape != (B, self.H) or states.shape != (B, self.S) or target.shape != (B, self.A):
            raise ValueError("pooled / states / target do not match the head's dimensions")
        p = self.drop_p if train else 0.0
        if p > 0.0 and keep_mask is None:
            keep_mask = (torch.rand(B, self.F, device=dev) >= p).to(torch.uint8)
        if p > 0.0:
            keep_mask = keep_mask.to(dev, torch.uint8).contiguous()
        need = int(self.lib.fvla_head_train_scratch_floats(B, self.H, self.S, self.Hd, self.F, self.A))
        if self._scratch is None or self._scratch.numel() < need or self._scratch.device != dev:
            self._scratch = torch.empty(need, device=dev, dtype=torch.float32)
        ptrs = (C.c_void_p * 12)(*[q.data_ptr() for q in self.params])
        with torch.cuda.device(dev):
            N.check(self.lib.fvla_head_forward_backward(
                B, self.H, self.S, self.Hd, self.F, self.A, ptrs, pooled.data_ptr(), states.data_ptr(),
                target.data_ptr(), keep_mask.data_ptr() if p > 0.0 else None, p, self.reducer.flat.data_ptr(),
                self.loss.data_ptr(), None, self._scratch.data_ptr(), need, N.stream_ptr()),
                "fvla_head_forward_backward")
        return self.loss


def train_step_native(policy, batch: Dict[str, torch.Tensor], optimizer: torch.optim.Optimizer,
                      reducer: HeadGradAllReduce, step: NativeHeadStep,
                      max_grad_norm: Optional[float] = 1.0) -> Tuple[float, float]:
    """`train_step` without autograd: frozen backbone in the engine -> pooled features -> native head
    forward / loss / backward into the flat buffer -> all-reduce -> clip -> optimizer."""
    from ..shared import pick_step

    policy.train()
    images, states, prompts = policy._prepare_inputs(batch)
    with torch.no_grad():
        pooled = policy.model.backbone(images, prompts, device=images.device)
    target = pick_step(batch["action"], 2, 0)
    loss = step(pooled, states.to(pooled.device), target.to(pooled.device), train=True)
    reducer.all_reduce()
    norm = reducer.clip_(max_grad_norm) if max_grad_norm is not None else reducer.grad_norm()
    optimizer.step()
    return float(loss), float(norm)


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
