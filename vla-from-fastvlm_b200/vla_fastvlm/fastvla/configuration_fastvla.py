"""Core FastVLA config.  Field names/defaults follow the reference dataclass
(src/vla_fastvlm/fastvla/configuration_fastvla.py:9-46); they live in `vla_fastvlm.shared` so the
LeRobot config cannot drift from this one."""
from __future__ import annotations

from dataclasses import dataclass

from vla_fastvlm.model.fastvlm_adapter import FastVLMBackboneConfig
from vla_fastvlm.shared import FastVLAModelFields

# FastVLAConfig field -> FastVLMBackboneConfig field
_BACKBONE_FIELD_MAP = {
    "vlm_model_name": "model_id",
    "bootstrap_model_name": "bootstrap_model_id",
    "image_size": "force_image_size",
    **{k: k for k in ("freeze_backbone", "resize_with_padding", "pad_value", "tokenizer_max_length",
                      "tokenizer_padding_side", "pad_to_max_length", "compute_dtype", "image_token_mode",
                      "pool_merged_last", "vision_chunk", "skip_unused_vision", "synthetic_seed",
                      "image_input_scale")},
}


@dataclass
class FastVLAConfig(FastVLAModelFields):
    """FastVLM backbone + action-expert head, SmolVLA-style layout."""

    def to_backbone_config(self) -> FastVLMBackboneConfig:
        return FastVLMBackboneConfig(**{dst: getattr(self, src) for src, dst in _BACKBONE_FIELD_MAP.items()})
