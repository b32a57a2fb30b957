from .fastvlm_adapter import FastVLMBackbone, FastVLMBackboneConfig

__all__ = ["FastVLMBackbone", "FastVLMBackboneConfig"]
