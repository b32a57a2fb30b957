"""Pieces shared by the standalone (`vla_fastvlm.fastvla`) and LeRobot (`vla_fastvlm.lerobot_fastvla`)
front-ends, so that both stay thin: the field set that defines a FastVLA model, prompt / time-step
normalisation of an observation batch, and the action queue.

Behavioural contract (checked against the reference in tests/test_oracle_vs_reference.py):
  * field names and defaults — fastvla/configuration_fastvla.py:17-32 and
    lerobot_fastvla/configuration_fastvla.py:30-50 of the reference
  * prompt handling — lerobot_fastvla/modeling_fastvla.py:91-105, fastvla/processor_fastvla.py:23-30
  * time-major inputs use the LAST observation step and the FIRST action step
    (lerobot_fastvla/modeling_fastvla.py:84-89, 129-131)
  * queue semantics — lerobot_fastvla/modeling_fastvla.py:78-79, 119-125
"""
from __future__ import annotations

from collections import deque
from dataclasses import dataclass, fields
from typing import Any, Deque, Iterable, List, Optional, Sequence

import torch


@dataclass
class FastVLAModelFields:
    """Everything that determines the network + its preprocessing.  Inherited by both config classes."""

    # FastVLM backbone
    vlm_model_name: str = "apple/FastVLM-0.5B"
    bootstrap_model_name: str = "apple/FastVLM-0.5B"
    freeze_backbone: bool = True
    # action head (state/action dims are overwritten from dataset features by the LeRobot wrapper)
    state_dim: int = 14
    action_dim: int = 14
    hidden_dim: int = 1024
    fusion_dim: int = 1024
    dropout: float = 0.1
    # preprocessing
    tokenizer_max_length: int = 64
    tokenizer_padding_side: str = "right"
    pad_to_max_length: bool = False
    resize_with_padding: bool = True
    image_size: Optional[int] = None
    pad_value: float = 0.0
    add_trailing_newline: bool = True
    # B200 engine switches — not in the reference; the defaults reproduce its behaviour
    compute_dtype: str = "float32"      # "bfloat16" = throughput mode
    image_token_mode: str = "none"      # "prefix" = splice the image tokens in front of the prompt
    pool_merged_last: bool = False
    vision_chunk: int = 0
    skip_unused_vision: bool = True
    synthetic_seed: int = 0
    image_input_scale: float = 1.0      # multiplies raw pixel values inside the ingest kernel: 1/255 takes uint8 camera
                                        # frames (HWC or CHW) straight from the host, 4x fewer PCIe bytes than fp32 [0,1]
    fuse_io_normalization: bool = False  # LeRobot plugin: the STATE normaliser and the ACTION unnormaliser of the
                                         # pre/post pipelines run inside the action-head kernel instead


MODEL_FIELD_NAMES = tuple(f.name for f in fields(FastVLAModelFields))


def copy_model_fields(src: Any, dst_cls):
    """Build `dst_cls` from the model fields of another config object."""
    return dst_cls(**{name: getattr(src, name) for name in MODEL_FIELD_NAMES})


def pick_step(t: torch.Tensor, batched_ndim: int, which: int) -> torch.Tensor:
    """Drop the time axis of a time-major tensor ((B,T,...) -> (B,...)); no-op otherwise."""
    return t[:, which] if t.ndim == batched_ndim + 1 else t


def as_prompt_list(task: Any, batch_size: int, trailing_newline: bool) -> List[str]:
    """None -> "", scalar -> broadcast, single-element list -> broadcast; optional "\\n" terminator."""
    if task is None:
        prompts = [""] * batch_size
    elif isinstance(task, (list, tuple)):
        prompts = [str(t) for t in task]
        if len(prompts) == 1 and batch_size > 1:
            prompts = prompts * batch_size
    else:
        prompts = [str(task)] * batch_size
    if trailing_newline:
        prompts = [p if p.endswith("\n") else p + "\n" for p in prompts]
    return prompts


class ActionQueue:
    """FIFO of per-step action tensors (B, D) shared by the whole batch; refilled from a chunk (B, n, D)."""

    def __init__(self, n_action_steps: int) -> None:
        self._q: Deque[torch.Tensor] = deque([], maxlen=n_action_steps)
        self.n_action_steps = n_action_steps

    def __len__(self) -> int:
        return len(self._q)

    def refill(self, chunk: torch.Tensor) -> None:
        self._q.extend(chunk[:, : self.n_action_steps].transpose(0, 1))

    def pop(self) -> torch.Tensor:
        return self._q.popleft()
