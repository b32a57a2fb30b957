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
from vla_fastvlm.ingest import ObservationStager
from vla_fastvlm.shared import ActionQueue, as_prompt_list, copy_model_fields, pick_step

IO_STATS_KEY = "fastvla.io_stats"  # batch entry through which the fused pre-processor hands the dataset statistics over

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
        self._stager: ObservationStager | None = None
        self._io_stats: Any = None      # statistics object last pushed into the engine (identity comparison)
        self._io_vectors: Dict[str, Tensor] = {}
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

    # ---- fused LeRobot (un)normaliser (config.fuse_io_normalization) ------------------------------------
    def set_dataset_stats(self, stats: Dict[str, Dict[str, Any]] | None) -> None:
        """MEAN_STD statistics of the state feature and of the action, as `make_fastvla_pre_post_processors` receives
        them.  With `fuse_io_normalization` the head kernel applies them (state normalised in front, action
        un-normalised behind) and the pipelines built by `make_fastvla_pre_post_processors` skip those steps."""
        self._io_stats = stats
        vec: Dict[str, Tensor] = {}
        if stats:
            for name, key in (("state", self._state_key), ("action", ACTION)):
                st = stats.get(key)
                if st is not None and "mean" in st and "std" in st:
                    vec[name + "_mean"] = torch.as_tensor(st["mean"], dtype=torch.float32).flatten().cpu()
                    vec[name + "_std"] = torch.as_tensor(st["std"], dtype=torch.float32).flatten().cpu()
        self._io_vectors = vec
        self._io_engine = None

    def _push_io_normalization(self) -> None:
        eng = self.model.backbone.model.engine
        if getattr(self, "_io_engine", None) is eng:  # a rebuilt engine (device move, load_state_dict) starts at identity
            return
        v = self._io_vectors
        eng.set_io_normalization(state_mean=v.get("state_mean"), state_std=v.get("state_std"),
                                 action_mean=v.get("action_mean"), action_std=v.get("action_std"))
        self._io_engine = eng

    def _adopt_batch_stats(self, batch: Dict[str, Any]) -> None:
        if self.config.fuse_io_normalization and IO_STATS_KEY in batch and batch[IO_STATS_KEY] is not self._io_stats:
            self.set_dataset_stats(batch[IO_STATS_KEY])

    def _predict_actions(self, batch: Dict[str, Tensor]) -> Tensor:
        self._adopt_batch_stats(batch)
        images, states, prompts = self._prepare_inputs(batch)
        home = self.model.backbone.model.device
        slot = None
        if home.type == "cuda" and not (images.is_cuda and states.is_cuda):
            # observations still in host memory (pinned uint8 camera frames in a serving loop): staged copy on a side
            # stream, forward ordered behind it, call returns without waiting
            if self._stager is None or self._stager.device != home:
                self._stager = ObservationStager(home)
            staged, slot = self._stager.stage({"images": images, "states": states})
            images, states = staged["images"], staged["states"]
        fused = self.config.fuse_io_normalization and bool(self._io_vectors)
        if fused and self.model._head_needs_autograd():
            # training path (eager head): the same arithmetic in torch, on the raw state
            v = self._io_vectors
            if "state_mean" in v:
                states = (states - v["state_mean"].to(states.device)) / (v["state_std"].to(states.device) + 1e-8)
        elif self.config.fuse_io_normalization:
            self._push_io_normalization()
        out = self.model(images, states, prompts, device=home if home.type == "cuda" else images.device)
        if slot is not None:
            self._stager.release(slot)
        return out

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
        if self.config.fuse_io_normalization and "action_mean" in self._io_vectors:
            # the fused pre-processor leaves the target un-normalised; the loss lives in normalised action space
            v = self._io_vectors
            target = (target - v["action_mean"].to(pred.device)) / (v["action_std"].to(pred.device) + 1e-8)
        loss = F.mse_loss(pred, target)
        value = loss.item()
        return loss, {"loss": value, "mse": value}
