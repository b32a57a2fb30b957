"""LeRobot `PreTrainedPolicy` front-end of the B200-native FastVLA model.

Interface parity with the reference wrapper (src/vla_fastvlm/lerobot_fastvla/modeling_fastvla.py:19-133):
`config_class`, `name`, `select_action`, `predict_action_chunk`, `forward`, `reset`, `get_optim_params`,
first-VISUAL / first-STATE feature resolution and the same ValueError texts.  `select_action` is the call
the headline metric (obs -> action chunks/s, p50 latency) is measured through."""
from __future__ import annotations

from typing import Any, Dict, List, Tuple

import torch
from torch import Tensor
from torch.nn import functional as F

from lerobot.configs.types import FeatureType
from lerobot.policies.pretrained import PreTrainedPolicy
from lerobot.utils.constants import ACTION

from vla_fastvlm.fastvla.configuration_fastvla import FastVLAConfig as CoreFastVLAConfig
from vla_fastvlm.fastvla.fastvlm_with_expert import FastVLMWithExpert
from vla_fastvlm.shared import ActionQueue, as_prompt_list, copy_model_fields, pick_step

from .configuration_fastvla import FastVLAConfig


class FastVLAPolicy(PreTrainedPolicy):
    config_class = FastVLAConfig
    name = "fastvla"

    def __init__(self, config: FastVLAConfig, **kwargs: Any):
        super().__init__(config)
        config.validate_features()
        self.config = config
        self._state_key, self._image_keys = self._resolve_input_keys()
        self._infer_io_dims_from_features()
        self.model = FastVLMWithExpert(copy_model_fields(self.config, CoreFastVLAConfig))
        self.reset()

    # ---- feature plumbing ---------------------------------------------------------------------------
    def _keys_of(self, kind: FeatureType) -> List[str]:
        return [name for name, ft in self.config.input_features.items() if ft.type is kind]

    def _resolve_input_keys(self) -> Tuple[str, List[str]]:
        if not self.config.input_features:
            raise ValueError("FastVLA requires input_features to be set.")
        state_keys, image_keys = self._keys_of(FeatureType.STATE), self._keys_of(FeatureType.VISUAL)
        if not state_keys:
            raise ValueError("No state feature found in input_features.")
        if not image_keys:
            raise ValueError("No visual feature found in input_features.")
        return state_keys[0], image_keys

    def _infer_io_dims_from_features(self) -> None:
        feats = self.config.input_features
        if feats and self._state_key in feats:
            self.config.state_dim = feats[self._state_key].shape[0]
        if self.config.action_feature is not None:
            self.config.action_dim = self.config.action_feature.shape[0]

    def get_optim_params(self):
        return self.parameters()

    def reset(self) -> None:
        self._queue = ActionQueue(self.config.n_action_steps)

    @property
    def _action_queue(self):  # name used by the reference implementation
        return self._queue._q

    # ---- observation batch -> model inputs ------------------------------------------------------------
    def _prepare_inputs(self, batch: Dict[str, Tensor]) -> Tuple[Tensor, Tensor, List[str]]:
        """First camera only, last observation step, prompts with trailing newline."""
        images = pick_step(batch[self._image_keys[0]], 4, -1)
        states = pick_step(batch[self._state_key], 2, -1)
        prompts = as_prompt_list(batch.get("task"), images.shape[0], self.config.add_trailing_newline)
        return images, states, prompts

    def _predict_actions(self, batch: Dict[str, Tensor]) -> Tensor:
        images, states, prompts = self._prepare_inputs(batch)
        return self.model(images, states, prompts, device=images.device)

    # ---- public API -------------------------------------------------------------------------------------
    @torch.no_grad()
    def predict_action_chunk(self, batch: Dict[str, Tensor]) -> Tensor:
        self.eval()
        return self._predict_actions(batch)[:, None, :]  # (B, chunk = 1, D): the head emits one step

    @torch.no_grad()
    def select_action(self, batch: Dict[str, Tensor]) -> Tensor:
        self.eval()
        if len(self._queue) == 0:
            self._queue.refill(self.predict_action_chunk(batch))
        return self._queue.pop()

    def forward(self, batch: Dict[str, Tensor]) -> Tuple[Tensor, dict]:
        pred = self._predict_actions(batch)
        target = pick_step(batch[ACTION], 2, 0).to(pred.device)
        loss = F.mse_loss(pred, target)
        value = loss.item()
        return loss, {"loss": value, "mse": value}
