"""`FastVLAProcessor` — prompt / time-step normalisation plus image ingest through the backbone's GPU
kernel.  Same methods as the reference class (src/vla_fastvlm/fastvla/processor_fastvla.py:11-43)."""
from __future__ import annotations

from typing import List, Union

import torch

from vla_fastvlm.shared import as_prompt_list, pick_step


class FastVLAProcessor:
    def __init__(self, config, backbone) -> None:
        self.config = config
        self.backbone = backbone

    def normalize_tasks(self, tasks: Union[List[str], str], batch_size: int) -> List[str]:
        return as_prompt_list(tasks, batch_size, self.config.add_trailing_newline)

    prepare_tasks = normalize_tasks

    def prepare_images(self, images: torch.Tensor, device: torch.device) -> torch.Tensor:
        """(B,[T,]C,H,W) -> (B,3,S,S) on `device` (last time step)."""
        return self.backbone._prepare_images_tensor(pick_step(images, 4, -1), device)

    def prepare_states(self, states: torch.Tensor, device: torch.device) -> torch.Tensor:
        return pick_step(states, 2, -1).to(device)
