"""LeRobot pre/post processor pipelines for FastVLA
(reference src/vla_fastvlm/lerobot_fastvla/processor_fastvla.py:22-61): rename -> add batch dim -> to(device) ->
normalise; unnormalise -> cpu.

With `config.fuse_io_normalization` the two MEAN_STD steps are not separate elementwise passes any more: the
pre-processor hands the dataset statistics to the policy through the batch (`fastvla.io_stats`) and leaves STATE /
ACTION untouched, the action-head kernel normalises the state in front and un-normalises the action behind
(include/fvla.h: fvla_set_io_normalization), and the post-processor only moves the result to the host.  The
observation is also left in host memory (the policy stages it on a side stream, vla_fastvlm/ingest.py)."""
from __future__ import annotations

from typing import Any

import torch

from lerobot.processor import (
    AddBatchDimensionProcessorStep,
    DeviceProcessorStep,
    NormalizerProcessorStep,
    PolicyAction,
    PolicyProcessorPipeline,
    RenameObservationsProcessorStep,
    UnnormalizerProcessorStep,
)
from lerobot.processor.converters import policy_action_to_transition, transition_to_policy_action
from lerobot.utils.constants import POLICY_POSTPROCESSOR_DEFAULT_NAME, POLICY_PREPROCESSOR_DEFAULT_NAME

from .configuration_fastvla import FastVLAConfig


class _AttachStatsStep:
    """Fused mode: pass the statistics object along with the batch instead of applying it."""

    def __init__(self, stats) -> None:
        self.stats = stats

    def __call__(self, batch):
        out = dict(batch)
        out["fastvla.io_stats"] = self.stats
        return out


def make_fastvla_pre_post_processors(
    config: FastVLAConfig,
    dataset_stats: dict[str, dict[str, torch.Tensor]] | None = None,
) -> tuple[
    PolicyProcessorPipeline[dict[str, Any], dict[str, Any]],
    PolicyProcessorPipeline[PolicyAction, PolicyAction],
]:
    if getattr(config, "fuse_io_normalization", False):
        pre = PolicyProcessorPipeline[dict[str, Any], dict[str, Any]](
            steps=[RenameObservationsProcessorStep(rename_map={}), AddBatchDimensionProcessorStep(),
                   _AttachStatsStep(dataset_stats)],
            name=POLICY_PREPROCESSOR_DEFAULT_NAME,
        )
        post = PolicyProcessorPipeline[PolicyAction, PolicyAction](
            steps=[DeviceProcessorStep(device="cpu")],
            name=POLICY_POSTPROCESSOR_DEFAULT_NAME,
            to_transition=policy_action_to_transition,
            to_output=transition_to_policy_action,
        )
        return pre, post
    pre = PolicyProcessorPipeline[dict[str, Any], dict[str, Any]](
        steps=[
            RenameObservationsProcessorStep(rename_map={}),
            AddBatchDimensionProcessorStep(),
            DeviceProcessorStep(device=config.device),
            NormalizerProcessorStep(
                features={**config.input_features, **config.output_features},
                norm_map=config.normalization_mapping,
                stats=dataset_stats,
                device=config.device,
            ),
        ],
        name=POLICY_PREPROCESSOR_DEFAULT_NAME,
    )
    post = PolicyProcessorPipeline[PolicyAction, PolicyAction](
        steps=[
            UnnormalizerProcessorStep(
                features=config.output_features,
                norm_map=config.normalization_mapping,
                stats=dataset_stats,
            ),
            DeviceProcessorStep(device="cpu"),
        ],
        name=POLICY_POSTPROCESSOR_DEFAULT_NAME,
        to_transition=policy_action_to_transition,
        to_output=transition_to_policy_action,
    )
    return pre, post
