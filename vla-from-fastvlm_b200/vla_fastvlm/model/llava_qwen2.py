"""Native stand-in for the `LlavaQwen2ForCausalLM` the reference pulls from the HF hub.

The reference instantiates the VLM with `AutoModelForCausalLM.from_pretrained(model_id,
trust_remote_code=True)` (src/vla_fastvlm/model/fastvlm_adapter.py:183-201) and then only ever calls
`model(**text_inputs, images=x, output_hidden_states=True)` and reads `hidden_states[-1]`
(:530-556).  This module provides that object for the B200 path: it owns the checkpoint tensors
until the CUDA engine is built, exposes the `config` attributes the adapter inspects
(`hidden_size`, `mm_vision_tower`, `model_type`), and forwards to libfvla.  There is no PyTorch
implementation of the network here — without the CUDA library nothing runs.
"""
from __future__ import annotations

import json
import logging
from pathlib import Path
from types import SimpleNamespace
from typing import Callable, Dict, Optional, Tuple

import torch
from torch import nn

from .arch import PRESETS, BackboneArch, arch_from_hf_config
from .engine import BACKBONE_KEY_PREFIX, NativeEngine
from .synthetic import synthetic_backbone_state_dict

logger = logging.getLogger(__name__)

SYNTHETIC_PREFIX = "synthetic:"


def _read_checkpoint_tensors(path: Path) -> Dict[str, torch.Tensor]:
    """All tensors of a local HF checkpoint directory (safetensors shards preferred, then *.bin)."""
    tensors: Dict[str, torch.Tensor] = {}
    st_files = sorted(path.glob("*.safetensors"))
    if st_files:
        from safetensors.torch import load_file

        for f in st_files:
            tensors.update(load_file(str(f), device="cpu"))
        return tensors
    bin_files = sorted(path.glob("pytorch_model*.bin")) or sorted(path.glob("*.pt"))
    for f in bin_files:
        tensors.update(torch.load(str(f), map_location="cpu", weights_only=True))
    if not tensors:
        raise FileNotFoundError(f"no *.safetensors / pytorch_model*.bin weights under {path}")
    return tensors


class LlavaQwen2Native(nn.Module):
    """Holds the architecture + weights of a llava_qwen2 checkpoint and the CUDA engine built from them."""

    def __init__(self, arch: BackboneArch, state_dict: Dict[str, torch.Tensor], name_or_path: str,
                 compute_dtype: torch.dtype = torch.float32) -> None:
        super().__init__()
        self.arch = arch
        self.compute_dtype = compute_dtype
        self._pending_sd: Optional[Dict[str, torch.Tensor]] = state_dict
        self._engine: Optional[NativeEngine] = None
        self._head_spec: Optional[Tuple[int, int, int, int]] = None
        self._head_sd_fn: Optional[Callable[[], Dict[str, torch.Tensor]]] = None
        self._engine_opts: Dict[str, object] = {}
        # what FastVLMBackbone reads from `model.config` (fastvlm_adapter.py:96-107, 280-298)
        self.config = SimpleNamespace(
            model_type="llava_qwen2",
            hidden_size=arch.text.hidden,
            mm_vision_tower=arch.mm_vision_tower,
            vocab_size=arch.text.vocab,
            num_hidden_layers=arch.text.layers,
            tokenizer_padding_side=arch.tokenizer_padding_side,
            output_hidden_states=False,
            _name_or_path=name_or_path,
        )
        # lets callers do `next(model.parameters()).device` like on the HF module (adapter :510)
        dev = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")
        self._device_anchor = nn.Parameter(torch.zeros(1, device=dev), requires_grad=False)

    # ---- construction ---------------------------------------------------------------------------
    @classmethod
    def from_pretrained(cls, model_id: str, compute_dtype: torch.dtype = torch.float32, seed: int = 0,
                        **_unused) -> "LlavaQwen2Native":
        """`synthetic:<preset>` -> seeded random init; otherwise a local checkpoint directory with
        config.json (model_type llava_qwen2) and safetensors / bin weights.  Hub ids cannot be
        resolved offline: that raises the same ValueError text transformers raises for an
        unregistered architecture, so the adapter's bootstrap branch sees what it expects."""
        if model_id.startswith(SYNTHETIC_PREFIX):
            key = model_id[len(SYNTHETIC_PREFIX):]
            if key not in PRESETS:
                raise ValueError(f"unknown synthetic preset {key!r}; choose from {sorted(PRESETS)}")
            arch = PRESETS[key]
            return cls(arch, synthetic_backbone_state_dict(arch, seed), model_id, compute_dtype)
        path = Path(model_id)
        cfg_path = path / "config.json"
        if not (path.is_dir() and cfg_path.is_file()):
            raise OSError(
                f"{model_id!r} is not a local checkpoint directory with config.json; hub downloads are not "
                "available to the native loader (download with scripts/download_fastvlm.sh first)."
            )
        with open(cfg_path, encoding="utf-8") as f:
            hf_cfg = json.load(f)
        if hf_cfg.get("model_type") != "llava_qwen2":
            raise ValueError(
                f"The checkpoint you are trying to load has model type `{hf_cfg.get('model_type')}` but the native "
                "FastVLA engine does not recognize this architecture (only `llava_qwen2`)."
            )
        arch = arch_from_hf_config(hf_cfg)
        return cls(arch, _read_checkpoint_tensors(path), model_id, compute_dtype)

    # ---- engine ---------------------------------------------------------------------------------
    def configure_engine(self, **opts) -> None:
        """pool_mode / vision_chunk / skip_unused_vision; must precede the first forward."""
        self._engine_opts.update(opts)

    def attach_head(self, state_dim: int, action_dim: int, hidden_dim: int, fusion_dim: int,
                    head_state_dict_fn: Callable[[], Dict[str, torch.Tensor]]) -> None:
        spec = (int(state_dim), int(action_dim), int(hidden_dim), int(fusion_dim))
        if self._engine is not None and self._head_spec != spec:
            raise RuntimeError("the CUDA engine was already built with a different head")
        self._head_spec = spec
        self._head_sd_fn = head_state_dict_fn

    @property
    def engine(self) -> NativeEngine:
        if self._engine is None:
            self._build_engine()
        return self._engine

    def _build_engine(self) -> None:
        if self._pending_sd is None:
            raise RuntimeError("backbone weights were already released")
        spec = self._head_spec or (0, 0, 0, 0)
        eng = NativeEngine(self.arch, dtype=self.compute_dtype, state_dim=spec[0], action_dim=spec[1],
                           hidden_dim=spec[2], fusion_dim=spec[3], **self._engine_opts)
        eng.load_state_dict(self._pending_sd, prefix=BACKBONE_KEY_PREFIX)
        if self._head_sd_fn is not None:
            eng.load_state_dict(self._head_sd_fn())
        missing = eng.missing_tensors()
        if missing:
            raise RuntimeError(f"checkpoint is missing {len(missing)} tensors, e.g. {missing[:3]}")
        eng.finalize()
        self._engine = eng
        self._pending_sd = None  # packed copies live on the device now

    def refresh_head(self) -> None:
        """Push the current (possibly just-trained) head parameters into the engine."""
        if self._engine is not None and self._head_sd_fn is not None:
            self._engine.load_state_dict(self._head_sd_fn())

    def forward(self, *args, **kwargs):  # pragma: no cover - guidance only
        raise RuntimeError(
            "LlavaQwen2Native has no eager forward; use FastVLMBackbone.forward / FastVLMWithExpert.forward, "
            "which call the CUDA engine."
        )
