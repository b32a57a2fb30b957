"""Shared builders for the parity tests: seeded weights, seeded inputs, oracle and engine set-up."""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch

from vla_fastvlm.model.arch import PRESETS, BackboneArch
from vla_fastvlm.model.synthetic import synthetic_backbone_state_dict, synthetic_head_state_dict

IMAGE_TOKEN_INDEX = -200
TINY_HEAD = dict(state_dim=6, action_dim=5, hidden_dim=64, fusion_dim=64)


def tiny_weights(seed: int = 0) -> Tuple[BackboneArch, Dict[str, torch.Tensor], Dict[str, torch.Tensor]]:
    arch = PRESETS["tiny"]
    sd = synthetic_backbone_state_dict(arch, seed)
    hsd = synthetic_head_state_dict(arch.text.hidden, TINY_HEAD["state_dim"], TINY_HEAD["action_dim"],
                                    TINY_HEAD["hidden_dim"], TINY_HEAD["fusion_dim"], seed + 1)
    return arch, sd, hsd


def make_inputs(B: int, h: int, w: int, T: int, vocab: int, state_dim: int, seed: int = 1,
                image_mode: str = "prefix", ragged: bool = True):
    """images (B,3,h,w) in [0,1], states, token ids (right padded; optional leading -200), mask."""
    g = torch.Generator().manual_seed(seed)
    images = torch.rand(B, 3, h, w, generator=g)
    states = torch.randn(B, state_dim, generator=g)
    ids = torch.randint(0, vocab, (B, T), generator=g)
    mask = torch.ones(B, T, dtype=torch.long)
    if ragged:
        for b in range(B):
            n = T if b == 0 else max(2, T - 1 - (b * 3) % (T - 1))
            mask[b, n:] = 0
    if image_mode == "prefix":
        ids[:, 0] = IMAGE_TOKEN_INDEX
    ids = ids * mask  # padded slots hold token 0 like a real pad id
    return images, states, ids, mask


def make_engine(arch, sd, hsd, dtype, head=TINY_HEAD, pool_mode="last_token", vision_chunk=0,
                skip_unused_vision=True):
    from vla_fastvlm.model.engine import BACKBONE_KEY_PREFIX, NativeEngine

    eng = NativeEngine(arch, dtype=dtype, pool_mode=pool_mode, vision_chunk=vision_chunk,
                       skip_unused_vision=skip_unused_vision, **head)
    eng.load_state_dict(sd, prefix=BACKBONE_KEY_PREFIX)
    eng.load_state_dict(hsd)
    assert eng.missing_tensors() == []
    eng.finalize()
    return eng


def rel_err(out: torch.Tensor, ref: torch.Tensor) -> float:
    out, ref = out.detach().float().cpu(), ref.detach().float().cpu()
    return float((out - ref).abs().max() / (ref.abs().max() + 1e-12))
