"""Standalone (non-LeRobot) FastVLA policy with the reference's public methods
(src/vla_fastvlm/fastvla/modeling_fastvla.py:14-77): forward / compute_loss / select_action / reset."""
from __future__ import annotations

from typing import Dict, List, Optional, Union

import torch
from torch import nn
from torch.nn import functional as F

from vla_fastvlm.fastvla.configuration_fastvla import FastVLAConfig
from vla_fastvlm.fastvla.fastvlm_with_expert import FastVLMWithExpert
from vla_fastvlm.fastvla.processor_fastvla import FastVLAProcessor
from vla_fastvlm.shared import pick_step


class FastVLAPolicy(nn.Module):
    config_class = FastVLAConfig
    name = "fastvla"

    def __init__(self, config: Optional[FastVLAConfig] = None) -> None:
        super().__init__()
        self.config = config if config is not None else FastVLAConfig()
        self.model = FastVLMWithExpert(self.config)
        self.processor = FastVLAProcessor(self.config, self.model.backbone)

    def forward(self, images: torch.Tensor, states: torch.Tensor, tasks: Union[List[str], str],
                device: Optional[torch.device] = None) -> torch.Tensor:
        """(B,[T,]C,H,W) images + (B,[T,]S) states + prompts -> (B, action_dim).

        The reference letterboxes here and once more inside the backbone; the second pass
        is an identity resize, so the single fused ingest inside the engine gives the same pixels."""
        device = images.device if device is None else device
        images = pick_step(images, 4, -1)
        prompts = self.processor.prepare_tasks(tasks, batch_size=images.shape[0])
        return self.model(images, self.processor.prepare_states(states, device), prompts, device=device)

    def compute_loss(self, batch: Dict[str, Union[torch.Tensor, List[str]]]) -> Dict[str, torch.Tensor]:
        pred = self.forward(batch["images"], batch["states"], batch["tasks"])
        mse = F.mse_loss(pred, batch["actions"].to(pred.device))
        return {"loss": mse, "mse": mse.detach()}

    @torch.inference_mode()
    def select_action(self, image: torch.Tensor, state: torch.Tensor, task: str,
                      device: torch.device) -> torch.Tensor:
        self.eval()
        out = self.forward(image[None].to(device), state[None].to(device), task, device=device)
        return out[0]

    def reset(self) -> None:  # API compatibility: the standalone policy keeps no queue
        return None
