"""Architecture description of the `llava_qwen2` backbone (FastViTHD + mlp2x_gelu + Qwen2).

The reference never spells these numbers out: it obtains them from the HF repo's remote code and
`config.json` at run time (src/vla_fastvlm/model/fastvlm_adapter.py:183-201).  This module is the
single place where they are written down for the B200 engine and where a
checkpoint `config.json` is translated into them.
"""
from __future__ import annotations

import json
import re
from dataclasses import asdict, dataclass, field
from pathlib import Path
from typing import Any, Dict, Optional, Tuple


@dataclass(frozen=True)
class VisionArch:
    """FastViTHD (`fastvithd`, timm `fastvit_mci3`), inference-mode (reparameterised) form."""

    image_size: int = 1024
    layers: Tuple[int, ...] = (2, 12, 24, 4, 2)
    dims: Tuple[int, ...] = (96, 192, 384, 768, 1536)
    attention: Tuple[bool, ...] = (False, False, False, True, True)
    pos_emb: Tuple[bool, ...] = (False, False, False, True, True)
    mlp_ratio: int = 4
    head_dim: int = 32
    cls_ratio: float = 2.0
    se_ratio: float = 0.0625

    @property
    def out_channels(self) -> int:
        return int(self.dims[-1] * self.cls_ratio)

    @property
    def se_reduced(self) -> int:
        return int(self.out_channels * self.se_ratio)

    @property
    def total_stride(self) -> int:
        return 4 * 2 ** (len(self.dims) - 1)

    @property
    def grid(self) -> int:
        return self.image_size // self.total_stride

    @property
    def num_tokens(self) -> int:
        return self.grid * self.grid


@dataclass(frozen=True)
class TextArch:
    """Qwen2 decoder shape (take every value from the checkpoint's config.json)."""

    hidden: int = 896
    layers: int = 24
    q_heads: int = 14
    kv_heads: int = 2
    head_dim: int = 64
    intermediate: int = 4864
    vocab: int = 151936
    rms_eps: float = 1e-6
    rope_theta: float = 1e6
    max_position: int = 32768


@dataclass(frozen=True)
class BackboneArch:
    name: str = "fastvlm-0.5b"
    vision: VisionArch = field(default_factory=VisionArch)
    text: TextArch = field(default_factory=TextArch)
    mm_vision_tower: str = "mobileclip_l_1024"
    tokenizer_padding_side: str = "right"

    @property
    def hidden_size(self) -> int:
        return self.text.hidden

    def to_json(self) -> Dict[str, Any]:
        d = asdict(self)
        d["model_type"] = "llava_qwen2"
        return d


PRESETS: Dict[str, BackboneArch] = {
    "fastvlm-0.5b": BackboneArch(),
    "fastvlm-1.5b": BackboneArch(
        name="fastvlm-1.5b",
        text=TextArch(hidden=1536, layers=28, q_heads=12, kv_heads=2, head_dim=128, intermediate=8960),
    ),
    "fastvlm-7b": BackboneArch(
        name="fastvlm-7b",
        text=TextArch(hidden=3584, layers=28, q_heads=28, kv_heads=4, head_dim=128, intermediate=18944,
                      vocab=152064),
    ),
    # Small shape with every block type, used by the parity tests (CPU oracle finishes in < 1 s).
    "tiny": BackboneArch(
        name="tiny",
        vision=VisionArch(image_size=256, layers=(1, 2, 2, 2, 1), dims=(16, 32, 64, 128, 256)),
        text=TextArch(hidden=128, layers=2, q_heads=2, kv_heads=1, head_dim=64, intermediate=256,
                      vocab=512, max_position=2048),
        mm_vision_tower="mobileclip_l_256",
    ),
    # Decoder-shape parity presets: the Qwen2-1.5B / 7B layer geometry (hidden, heads, head_dim 128, intermediate)
    # at depth 2 behind the small tower at 1024^2 (256 image tokens => T' = 256 + T_text like the full models), with a
    # small vocabulary so the CPU oracle and the host weight generation stay cheap.
    "dec-1.5b-2l": BackboneArch(
        name="dec-1.5b-2l",
        vision=VisionArch(image_size=1024, layers=(1, 1, 1, 1, 1), dims=(16, 32, 64, 128, 256)),
        text=TextArch(hidden=1536, layers=2, q_heads=12, kv_heads=2, head_dim=128, intermediate=8960, vocab=2048),
        mm_vision_tower="mobileclip_l_1024",
    ),
    "dec-7b-2l": BackboneArch(
        name="dec-7b-2l",
        vision=VisionArch(image_size=1024, layers=(1, 1, 1, 1, 1), dims=(16, 32, 64, 128, 256)),
        text=TextArch(hidden=3584, layers=2, q_heads=28, kv_heads=4, head_dim=128, intermediate=18944, vocab=2048),
        mm_vision_tower="mobileclip_l_1024",
    ),
}


def arch_from_hf_config(cfg: Dict[str, Any]) -> BackboneArch:
    """Translate a llava_qwen2 `config.json` (Apple FastVLM checkpoints) into a BackboneArch."""
    if cfg.get("model_type") not in ("llava_qwen2", None):
        raise ValueError(f"expected model_type llava_qwen2, got {cfg.get('model_type')!r}")
    if "vision" in cfg and "text" in cfg:  # our own to_json() form
        v, t = dict(cfg["vision"]), dict(cfg["text"])
        for k in ("layers", "dims", "attention", "pos_emb"):
            v[k] = tuple(v[k])
        return BackboneArch(name=cfg.get("name", "custom"), vision=VisionArch(**v), text=TextArch(**t),
                            mm_vision_tower=cfg.get("mm_vision_tower", "mobileclip_l_1024"),
                            tokenizer_padding_side=cfg.get("tokenizer_padding_side", "right"))
    hidden = int(cfg["hidden_size"])
    heads = int(cfg["num_attention_heads"])
    tower = str(cfg.get("mm_vision_tower", "mobileclip_l_1024"))
    m = re.search(r"(\d{2,4})$", tower)
    image_size = int(m.group(1)) if m else 1024
    text = TextArch(
        hidden=hidden,
        layers=int(cfg["num_hidden_layers"]),
        q_heads=heads,
        kv_heads=int(cfg.get("num_key_value_heads", heads)),
        head_dim=int(cfg.get("head_dim", hidden // heads)),
        intermediate=int(cfg["intermediate_size"]),
        vocab=int(cfg["vocab_size"]),
        rms_eps=float(cfg.get("rms_norm_eps", 1e-6)),
        rope_theta=float(cfg.get("rope_theta", 1e6)),
        max_position=int(cfg.get("max_position_embeddings", 32768)),
    )
    vis = dict(image_size=image_size)
    if isinstance(cfg.get("vision_arch"), dict):
        # Not an Apple field: the checkpoints imply FastViTHD through the tower name alone.  Exports of OTHER tower
        # geometries (the test checkpoints, distilled towers) spell the geometry out under this key.
        for k, val in cfg["vision_arch"].items():
            vis[k] = tuple(val) if isinstance(val, list) else val
    return BackboneArch(name=str(cfg.get("_name_or_path", "llava_qwen2")), vision=VisionArch(**vis),
                        text=text, mm_vision_tower=tower,
                        tokenizer_padding_side=str(cfg.get("tokenizer_padding_side", "right")))


def load_arch(model_id: str) -> Optional[BackboneArch]:
    """Preset name (`synthetic:<preset>` or bare preset) or a local checkpoint directory."""
    key = model_id.split(":", 1)[1] if model_id.startswith("synthetic:") else model_id
    if key in PRESETS:
        return PRESETS[key]
    p = Path(model_id) / "config.json"
    if p.is_file():
        with open(p, encoding="utf-8") as f:
            return arch_from_hf_config(json.load(f))
    return None
