"""Native stand-in for the `LlavaQwen2ForCausalLM` the reference pulls from the HF hub.

The reference instantiates the VLM with `AutoModelForCausalLM.from_pretrained(model_id,
trust_remote_code=True)` (src/vla_fastvlm/model/fastvlm_adapter.py:183-201) and then only ever calls
`model(**text_inputs, images=x, output_hidden_states=True)` and reads `hidden_states[-1]`
(:530-556).  This module provides that object for the B200 path: it owns the checkpoint tensors
until the CUDA engine is built, exposes the `config` attributes the adapter inspects
(`hidden_size`, `mm_vision_tower`, `model_type`), and forwards to libfvla.  There is no PyTorch
implementation of the network here — without the CUDA library nothing runs.

Checkpoint contract (reference `policy_state_dict.pt`, training/trainer.py:246-255 -> utils/checkpoint.py:14-47):
the HF module's tensors appear in the policy's `state_dict()` under `model.backbone.model.<hf name>` and
`load_state_dict(strict=True)` consumes them.  The module therefore keeps the checkpoint tensors (host memory, stored
dtype) next to the engine's packed device copies, serialises them under their HF names, and rebuilds the engine when a
`load_state_dict` replaces them.
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
        # checkpoint tensors under their HF names (host memory): source of the engine build AND of state_dict()
        self._weights: Dict[str, torch.Tensor] = {k: v.detach() for k, v in state_dict.items()}
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
        # follows .to(device) / .cuda() like a parameter would, but never shows up in state_dict()
        dev = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")
        self.register_buffer("_device_anchor", torch.zeros(1, device=dev), persistent=False)

    @property
    def device(self) -> torch.device:
        return self._device_anchor.device

    # ---- state dict: the HF tensors under their HF names -----------------------------------------
    def _save_to_state_dict(self, destination, prefix, keep_vars):
        super()._save_to_state_dict(destination, prefix, keep_vars)
        for k, v in self._weights.items():
            destination[prefix + k] = v

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys,
                              error_msgs):
        touched = False
        for k, cur in self._weights.items():
            key = prefix + k
            if key not in state_dict:
                if not k.startswith("lm_head."):  # never computed: kept for round trips, never required
                    missing_keys.append(key)
                continue
            new = state_dict[key]
            if tuple(new.shape) != tuple(cur.shape):
                error_msgs.append(f"size mismatch for {key}: copying a param with shape {tuple(new.shape)} from "
                                  f"checkpoint, the shape in current model is {tuple(cur.shape)}.")
                continue
            self._weights[k] = new.detach().to("cpu")
            touched = True
        for key in state_dict:
            if key.startswith(prefix):
                k = key[len(prefix):]
                # the LM head is never computed; tied checkpoints list it next to embed_tokens
                if k not in self._weights and not k.startswith("lm_head."):
                    unexpected_keys.append(key)
        if touched:
            self._engine = None  # packed device copies are stale: rebuilt from the new tensors on the next forward

    def _apply(self, fn, recurse=True):
        before = self._device_anchor.device
        out = super()._apply(fn, recurse)
        if self._engine is not None and self._device_anchor.device != before:
            self._engine = None  # the engine lives on ONE device: rebuild where the module now is
        return out

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
        if "mm_vision_tower" not in hf_cfg and "vision" not in hf_cfg:
            # Same situation as the reference's stock path on a local export without `auto_map`
            # (fastvlm_adapter.py:183-206): the config names the architecture but does not describe it; the defaults
            # live with the bootstrap model.  Same message shape, so `_needs_llava_qwen2_bootstrap` routes it.
            raise ValueError(
                "The checkpoint you are trying to load has model type `llava_qwen2` but its config.json does not "
                "describe the vision tower (no `mm_vision_tower`): the loader does not recognize this architecture "
                "without the bootstrap model's defaults."
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
        dev = self.device
        if dev.type != "cuda":
            raise RuntimeError("the FastVLA B200 path runs on CUDA only: move the policy to a CUDA device first "
                               "(no CPU fallback)")
        spec = self._head_spec or (0, 0, 0, 0)
        eng = NativeEngine(self.arch, dtype=self.compute_dtype, state_dim=spec[0], action_dim=spec[1],
                           hidden_dim=spec[2], fusion_dim=spec[3], device=dev, **self._engine_opts)
        eng.load_state_dict({k: v for k, v in self._weights.items() if not k.startswith("lm_head.")},
                            prefix=BACKBONE_KEY_PREFIX)
        if self._head_sd_fn is not None:
            eng.load_state_dict(self._head_sd_fn())
        missing = eng.missing_tensors()
        if missing:
            raise RuntimeError(f"checkpoint is missing {len(missing)} tensors, e.g. {missing[:3]}")
        eng.finalize()
        self._engine = eng

    def refresh_head(self) -> None:
        """Push the current (possibly just-trained) head parameters into the engine."""
        if self._engine is not None and self._head_sd_fn is not None:
            self._engine.load_state_dict(self._head_sd_fn())

    def forward(self, *args, **kwargs):  # pragma: no cover - guidance only
        raise RuntimeError(
            "LlavaQwen2Native has no eager forward; use FastVLMBackbone.forward / FastVLMWithExpert.forward, "
            "which call the CUDA engine."
        )
