"""LeRobot config for the FastVLA policy, registered under the name `fastvla`
(reference src/vla_fastvlm/lerobot_fastvla/configuration_fastvla.py:11-106).

The model fields come from `vla_fastvlm.shared.FastVLAModelFields`; this class adds what LeRobot
needs: the chunk interface, the normalisation map, optimiser / scheduler presets and delta indices."""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, Optional, Tuple

from lerobot.configs.policies import PreTrainedConfig
from lerobot.configs.types import FeatureType, NormalizationMode
from lerobot.optim.optimizers import AdamWConfig
from lerobot.optim.schedulers import CosineDecayWithWarmupSchedulerConfig

from vla_fastvlm.shared import FastVLAModelFields


def _default_norm_map() -> Dict[str, NormalizationMode]:
    # images pass through; state and action are standardised with dataset statistics
    return {"VISUAL": NormalizationMode.IDENTITY, "STATE": NormalizationMode.MEAN_STD,
            "ACTION": NormalizationMode.MEAN_STD}


@PreTrainedConfig.register_subclass("fastvla")
@dataclass
class FastVLAConfig(FastVLAModelFields, PreTrainedConfig):
    # chunk interface (the head regresses ONE action; chunk_size only widens what the dataset loads)
    n_obs_steps: int = 1
    chunk_size: int = 1
    n_action_steps: int = 1
    normalization_mapping: Dict[str, NormalizationMode] = field(default_factory=_default_norm_map)

    # AdamW + cosine-with-warmup presets
    optimizer_lr: float = 1e-4
    optimizer_betas: Tuple[float, float] = (0.9, 0.95)
    optimizer_eps: float = 1e-8
    optimizer_weight_decay: float = 1e-4
    optimizer_grad_clip_norm: float = 1.0
    scheduler_warmup_steps: int = 500
    scheduler_decay_steps: int = 20_000
    scheduler_decay_lr: float = 2.5e-6

    def __post_init__(self):
        super().__post_init__()
        if self.n_action_steps > self.chunk_size:
            raise ValueError(
                "n_action_steps must be <= chunk_size. "
                f"Got n_action_steps={self.n_action_steps}, chunk_size={self.chunk_size}."
            )

    def _has(self, kind: FeatureType) -> bool:
        return any(ft.type is kind for ft in self.input_features.values())

    def validate_features(self) -> None:
        if not self.input_features:
            return
        if not self._has(FeatureType.VISUAL):
            raise ValueError("FastVLA requires at least one visual observation feature.")
        if not self._has(FeatureType.STATE):
            raise ValueError("FastVLA requires at least one state observation feature.")

    def get_optimizer_preset(self) -> AdamWConfig:
        return AdamWConfig(lr=self.optimizer_lr, betas=self.optimizer_betas, eps=self.optimizer_eps,
                           weight_decay=self.optimizer_weight_decay, grad_clip_norm=self.optimizer_grad_clip_norm)

    def get_scheduler_preset(self) -> CosineDecayWithWarmupSchedulerConfig:
        return CosineDecayWithWarmupSchedulerConfig(
            peak_lr=self.optimizer_lr, decay_lr=self.scheduler_decay_lr,
            num_warmup_steps=self.scheduler_warmup_steps, num_decay_steps=self.scheduler_decay_steps)

    @property
    def observation_delta_indices(self) -> list:
        return [0]

    @property
    def action_delta_indices(self) -> list:
        return list(range(self.chunk_size))

    @property
    def reward_delta_indices(self) -> Optional[list]:
        return None
