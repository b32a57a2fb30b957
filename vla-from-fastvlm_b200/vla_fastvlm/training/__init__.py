"""Data-parallel training step of the FastVLA head (the backbone is frozen and runs in the CUDA engine)."""
from .head_dp import HeadGradAllReduce, NativeHeadStep, train_step, train_step_native

__all__ = ["HeadGradAllReduce", "NativeHeadStep", "train_step", "train_step_native"]
