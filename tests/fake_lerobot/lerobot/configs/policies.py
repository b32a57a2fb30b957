from dataclasses import dataclass, field
from typing import Dict, Optional

from .types import FeatureType, PolicyFeature

_REGISTRY: Dict[str, type] = {}


@dataclass
class PreTrainedConfig:
    n_obs_steps: int = 1
    input_features: Dict[str, PolicyFeature] = field(default_factory=dict)
    output_features: Dict[str, PolicyFeature] = field(default_factory=dict)
    device: Optional[str] = None
    use_amp: bool = False

    def __post_init__(self):
        pass

    @classmethod
    def register_subclass(cls, name: str):
        def deco(sub):
            _REGISTRY[name] = sub
            sub._choice_name = name
            return sub

        return deco

    @classmethod
    def get_choice_class(cls, name: str):
        return _REGISTRY[name]

    @property
    def type(self) -> str:
        return getattr(self, "_choice_name", "")

    @property
    def action_feature(self) -> Optional[PolicyFeature]:
        for ft in self.output_features.values():
            if ft.type is FeatureType.ACTION:
                return ft
        return None
