from dataclasses import dataclass


@dataclass
class CosineDecayWithWarmupSchedulerConfig:
    num_warmup_steps: int
    num_decay_steps: int
    peak_lr: float
    decay_lr: float
