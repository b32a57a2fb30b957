"""Data-parallel training step of the FastVLA head (the backbone is frozen and runs in the CUDA engine)."""
from .head_dp import HeadGradAllReduce, train_step

__all__ = ["HeadGradAllReduce", "train_step"]
