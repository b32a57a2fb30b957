"""Processor steps reduced to what the FastVLA pipelines need (identity VISUAL, mean/std STATE/ACTION)."""
from typing import Any, Dict, Generic, TypeVar

import torch

PolicyAction = torch.Tensor
TIn = TypeVar("TIn")
TOut = TypeVar("TOut")


class _Step:
    def __call__(self, x):
        return x


class RenameObservationsProcessorStep(_Step):
    def __init__(self, rename_map=None):
        self.rename_map = rename_map or {}

    def __call__(self, x):
        return {self.rename_map.get(k, k): v for k, v in x.items()}


class AddBatchDimensionProcessorStep(_Step):
    def __call__(self, x):
        out = {}
        for k, v in x.items():
            if isinstance(v, torch.Tensor) and ((k.startswith("observation.image") and v.ndim == 3) or
                                                (k.endswith("state") and v.ndim == 1)):
                v = v.unsqueeze(0)
            out[k] = v
        return out


class DeviceProcessorStep(_Step):
    def __init__(self, device=None):
        self.device = device

    def __call__(self, x):
        if self.device is None:
            return x
        if isinstance(x, torch.Tensor):
            return x.to(self.device)
        return {k: (v.to(self.device) if isinstance(v, torch.Tensor) else v) for k, v in x.items()}


class NormalizerProcessorStep(_Step):
    def __init__(self, features=None, norm_map=None, stats=None, device=None):
        self.features, self.norm_map, self.stats = features or {}, norm_map or {}, stats or {}

    def __call__(self, x):
        out = dict(x)
        for k, ft in self.features.items():
            mode = self.norm_map.get(ft.type.value)
            if k in out and k in self.stats and mode is not None and mode.value == "MEAN_STD":
                s = self.stats[k]
                out[k] = (out[k] - s["mean"].to(out[k].device)) / (s["std"].to(out[k].device) + 1e-8)
        return out


class UnnormalizerProcessorStep(_Step):
    def __init__(self, features=None, norm_map=None, stats=None):
        self.features, self.norm_map, self.stats = features or {}, norm_map or {}, stats or {}

    def __call__(self, x):
        for k, ft in self.features.items():
            mode = self.norm_map.get(ft.type.value)
            if k in self.stats and mode is not None and mode.value == "MEAN_STD":
                s = self.stats[k]
                return x * s["std"].to(x.device) + s["mean"].to(x.device)  # LeRobot: eps only on the forward side
        return x


class PolicyProcessorPipeline(Generic[TIn, TOut]):
    def __init__(self, steps=None, name="", to_transition=None, to_output=None):
        self.steps, self.name = list(steps or []), name

    def __call__(self, x):
        for s in self.steps:
            x = s(x)
        return x
