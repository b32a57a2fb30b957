"""Definition of the golden cases: inputs are regenerated from seeds (torch CPU generator), so the
fixtures only need to store checksums + reference outputs.  Imported by make_golden.py (reference
side) and by the tests (oracle / CUDA side)."""
from __future__ import annotations

from typing import Dict, List, Union

import torch

PROMPTS = ["pick up the red block", "open the drawer", "push", "insert the peg into the hole"]

# name -> spec
CASES: Dict[str, dict] = {
    "prefix_letterbox": dict(seed=11, image=("rand", (3, 3, 120, 160)), state=(3, 6), tasks=PROMPTS[:3], prefix=True,
                             pool="last_token"),
    "none_ragged": dict(seed=12, image=("rand", (3, 3, 96, 64)), state=(3, 6), tasks=PROMPTS[1:4], prefix=False,
                        pool="last_token"),
    "none_mean_pool": dict(seed=13, image=("rand", (2, 3, 256, 256)), state=(2, 6), tasks=PROMPTS[:2], prefix=False,
                           pool="mean_pool"),
    "prefix_bhwc_uint8_values": dict(seed=14, image=("u8", (2, 72, 100, 3)), state=(2, 6), tasks="stack the cubes",
                                     prefix=True, pool="last_token"),
    "prefix_time_major": dict(seed=15, image=("rand", (2, 2, 3, 80, 80)), state=(2, 2, 6), tasks=PROMPTS[2:4],
                              prefix=True, pool="last_token"),
}

TINY_HEAD = dict(state_dim=6, action_dim=5, hidden_dim=64, fusion_dim=64)


def case_inputs(name: str):
    spec = CASES[name]
    g = torch.Generator().manual_seed(spec["seed"])
    kind, shape = spec["image"]
    if kind == "rand":
        images = torch.rand(*shape, generator=g)
    else:  # integer-valued pixels 0..255 stored as float (what LeRobot gives before /255)
        images = torch.randint(0, 256, shape, generator=g).float()
    states = torch.randn(*spec["state"], generator=g)
    tasks: Union[str, List[str]] = spec["tasks"]
    return images, states, tasks


def checksum(t: torch.Tensor) -> float:
    return float(t.double().sum() + (t.double() * torch.arange(1, t.numel() + 1, dtype=torch.float64).reshape(t.shape) % 7).sum())
