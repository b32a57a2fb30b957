from dataclasses import dataclass
from typing import Tuple


@dataclass
class AdamWConfig:
    lr: float = 1e-3
    betas: Tuple[float, float] = (0.9, 0.999)
    eps: float = 1e-8
    weight_decay: float = 1e-2
    grad_clip_norm: float = 10.0
