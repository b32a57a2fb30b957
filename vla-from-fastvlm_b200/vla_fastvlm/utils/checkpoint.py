"""Policy checkpoint directories in the reference's format: `policy_config.json` (the dataclass fields of the policy
config) + `policy_state_dict.pt` (the policy's full `state_dict()`, frozen backbone included).

Written by the reference trainer (src/vla_fastvlm/training/trainer.py:246-255) and read back by
`load_policy_from_checkpoint` (src/vla_fastvlm/utils/checkpoint.py:14-47): the policy is rebuilt from the stored
config — which re-loads the backbone named by `vlm_model_name` — and then `load_state_dict(strict)` overwrites every
tensor, backbone included, with the checkpoint's.  Same protocol here; the legacy `FastVLMPolicy` branch of the
reference (configs without `vlm_model_name`) is out of scope and raises.
"""
from __future__ import annotations

import json
from dataclasses import asdict, fields
from pathlib import Path
from typing import Optional, Tuple, Union

import torch

from vla_fastvlm.fastvla import FastVLAConfig, FastVLAPolicy

CONFIG_NAME = "policy_config.json"
WEIGHTS_NAME = "policy_state_dict.pt"


def save_policy_checkpoint(policy: FastVLAPolicy, checkpoint_dir: Union[str, Path]) -> Path:
    """What `Trainer._save_checkpoint` leaves behind for inference (trainer.py:250-255)."""
    path = Path(checkpoint_dir)
    path.mkdir(parents=True, exist_ok=True)
    with open(path / CONFIG_NAME, "w", encoding="utf-8") as f:
        json.dump(asdict(policy.config), f, indent=2)
    torch.save({k: v.detach().cpu() for k, v in policy.state_dict().items()}, path / WEIGHTS_NAME)
    return path


def load_policy_from_checkpoint(checkpoint_dir: Union[str, Path], device_preference: Optional[str] = None,
                                strict: bool = True) -> Tuple[FastVLAPolicy, torch.device]:
    """Rebuild a FastVLA policy from a checkpoint directory; returns (policy in eval mode, device).
    `device_preference`: a CUDA device string; anything else raises (there is no CPU path)."""
    path = Path(checkpoint_dir)
    config_path, weights_path = path / CONFIG_NAME, path / WEIGHTS_NAME
    if not config_path.exists():
        raise FileNotFoundError(f"Missing {CONFIG_NAME} in {checkpoint_dir}")
    if not weights_path.exists():
        raise FileNotFoundError(f"Missing {WEIGHTS_NAME} in {checkpoint_dir}")
    with open(config_path, "r", encoding="utf-8") as f:
        config_dict = json.load(f)
    if "vlm_model_name" not in config_dict:
        raise ValueError("legacy FastVLMPolicy checkpoints (no `vlm_model_name` in policy_config.json) are not "
                         "supported by the B200 path")
    known = {f.name for f in fields(FastVLAConfig)}
    unknown = sorted(set(config_dict) - known)
    if unknown:
        raise TypeError(f"policy_config.json has fields FastVLAConfig does not know: {unknown}")
    policy = FastVLAPolicy(FastVLAConfig(**config_dict))
    state_dict = torch.load(weights_path, map_location="cpu", weights_only=True)
    policy.load_state_dict(state_dict, strict=strict)
    device = torch.device(device_preference) if device_preference else torch.device("cuda", torch.cuda.current_device())
    if device.type != "cuda":
        raise RuntimeError(f"the FastVLA B200 path runs on CUDA only, got device_preference={device_preference!r}")
    policy.to(device)
    policy.eval()
    return policy, device
