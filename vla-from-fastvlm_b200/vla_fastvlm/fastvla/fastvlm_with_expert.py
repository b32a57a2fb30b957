"""`FastVLMWithExpert`: FastVLM backbone + action head, B200-native.

Same constructor, attributes (`backbone`, `state_projection`, `fusion`, `action_head` — hence the same
state_dict keys) and `forward(images, states, tasks, device=None) -> (B, action_dim)` as the reference
(src/vla_fastvlm/fastvla/fastvlm_with_expert.py:12-54).  In eval mode the whole forward, head included,
is ONE call into the CUDA engine (the head runs as its fused kernel on the engine's packed copy of
these parameters).  Whenever the head needs autograd (grad enabled and a head parameter requires it) or
Dropout is active (train mode), the frozen backbone still runs in the engine (the reference computes it
under no_grad too, fastvlm_adapter.py:501) and only the ~3 M-parameter head is evaluated eagerly on the
pooled features — the same numbers the reference's head gives in that mode.
"""
from __future__ import annotations

from typing import Dict, List

import torch
from torch import nn

from vla_fastvlm.fastvla.configuration_fastvla import FastVLAConfig
from vla_fastvlm.model.fastvlm_adapter import FastVLMBackbone


class FastVLMWithExpert(nn.Module):
    def __init__(self, config: FastVLAConfig) -> None:
        super().__init__()
        self.config = config
        self.backbone = FastVLMBackbone(self.config.to_backbone_config())

        self.state_projection = nn.Sequential(
            nn.LayerNorm(self.config.state_dim),
            nn.Linear(self.config.state_dim, self.config.hidden_dim),
            nn.SiLU(),
        )
        fusion_input = self.backbone.output_dim + self.config.hidden_dim
        self.fusion = nn.Sequential(
            nn.Linear(fusion_input, self.config.fusion_dim),
            nn.LayerNorm(self.config.fusion_dim),
            nn.SiLU(),
            nn.Dropout(self.config.dropout),
            nn.Linear(self.config.fusion_dim, self.config.fusion_dim),
            nn.SiLU(),
        )
        self.action_head = nn.Linear(self.config.fusion_dim, self.config.action_dim)

        self._head_version = None  # parameter versions last pushed to the engine
        self.backbone.model.attach_head(self.config.state_dim, self.config.action_dim, self.config.hidden_dim,
                                        self.config.fusion_dim, self._head_state_dict)

    # ---- head parameters <-> engine ----------------------------------------------------------------
    def _head_params(self) -> Dict[str, torch.Tensor]:
        out: Dict[str, torch.Tensor] = {}
        for prefix, mod in (("state_projection", self.state_projection), ("fusion", self.fusion),
                            ("action_head", self.action_head)):
            for k, v in mod.state_dict().items():
                out[f"{prefix}.{k}"] = v
        return out

    @staticmethod
    def _fingerprint(sd: Dict[str, torch.Tensor]):
        # in-place updates bump _version; `.to(dtype)` / `.half()` / `param.data = ...` swap the storage instead
        return tuple((v._version, v.data_ptr(), v.dtype) for v in sd.values())

    def _head_state_dict(self) -> Dict[str, torch.Tensor]:
        sd = self._head_params()
        self._head_version = self._fingerprint(sd)
        return sd

    def _sync_head(self) -> None:
        """Re-upload the head if an optimizer step / load_state_dict / dtype change touched it since the last push."""
        model = self.backbone.model
        _ = model.engine  # builds (and loads the head) on first use
        if self._fingerprint(self._head_params()) != self._head_version:
            model.refresh_head()

    def _head_needs_autograd(self) -> bool:
        """The fused head kernel has no autograd graph and no Dropout: use it only where the reference's eager head
        would give the same numbers — grad disabled (or nothing to differentiate) and Dropout inactive."""
        if self.training and self.config.dropout > 0.0:
            return True
        if not torch.is_grad_enabled():
            return False
        return any(p.requires_grad for mod in (self.state_projection, self.fusion, self.action_head)
                   for p in mod.parameters())

    # ---- forward -----------------------------------------------------------------------------------
    def forward(self, images: torch.Tensor, states: torch.Tensor, tasks: List[str],
                device: torch.device | None = None) -> torch.Tensor:
        if device is None:
            device = images.device if isinstance(images, torch.Tensor) else next(self.parameters()).device
        if self._head_needs_autograd():
            with torch.no_grad():
                backbone_features = self.backbone(images, tasks, device=device)
            state_features = self.state_projection(states.to(backbone_features.device))
            fused = self.fusion(torch.cat([backbone_features, state_features], dim=-1))
            return self.action_head(fused)
        self._sync_head()
        return self.backbone._run(images, tasks, states, device)
