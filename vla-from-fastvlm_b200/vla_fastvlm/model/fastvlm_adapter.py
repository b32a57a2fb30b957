"""B200-native `FastVLMBackbone`: same surface as the reference adapter
(src/vla_fastvlm/model/fastvlm_adapter.py), different engine.

Kept from the reference (names, argument meaning, error types and key phrases):
  * `FastVLMBackboneConfig` fields and defaults (:58-80)
  * `FastVLMBackbone(config)`, `.forward(images, tasks, device=None) -> (B, H)`, `.backbone(...)`
    alias, `.output_dim`, `.expected_size`, `.tokenizer`, `.model`, `._prepare_images_tensor`
  * the loader protocol incl. the `llava_qwen2` bootstrap fallback (:183-241), expected-size
    resolution order (:245-335) and the too-small `image_size` guard (:145-154)
Changed underneath:
  * the VLM is the CUDA engine (libfvla) instead of HF remote code; the LM head is never computed
  * image canonicalisation runs in one GPU kernel — no GPU->CPU->GPU bounce (:485-488)
  * pooling + final RMSNorm are fused in-kernel; the action head can be fused too (FastVLMWithExpert)
New, opt-in fields (defaults reproduce the reference): `compute_dtype`, `image_token_mode`,
`vision_chunk`, `skip_unused_vision`.  There is no CPU fallback: without CUDA, construction raises.
"""
from __future__ import annotations

import json
import logging
import re
from dataclasses import dataclass
from pathlib import Path
from typing import Any, Dict, List, Optional, Tuple, Union

import torch
from torch import nn

from .. import _native as N
from .llava_qwen2 import SYNTHETIC_PREFIX, LlavaQwen2Native
from .tokenizer import SimpleByteTokenizer

Tensor = torch.Tensor
ImageLike = Union[Tensor, "numpy.ndarray", "PIL.Image.Image"]  # type: ignore  # noqa: F821
logger = logging.getLogger(__name__)

IMAGE_TOKEN_INDEX = N.IMAGE_TOKEN_INDEX
_IMAGENET_MEAN = (0.485, 0.456, 0.406)
_IMAGENET_STD = (0.229, 0.224, 0.225)
_DTYPES = {"float32": torch.float32, "fp32": torch.float32, "bfloat16": torch.bfloat16, "bf16": torch.bfloat16}


@dataclass
class FastVLMBackboneConfig:
    model_id: str = "apple/FastVLM-0.5B"
    # Used only when loading local llava_qwen2 checkpoints missing `auto_map`.
    bootstrap_model_id: str = "apple/FastVLM-0.5B"
    freeze_backbone: bool = True
    image_feature_pool: str = "last_token"  # "last_token" | "mean_pool"
    fallback_image_size: int = 512
    force_image_size: Optional[int] = None
    normalize_imagenet: bool = False
    resize_with_padding: bool = True
    pad_value: float = 0.0
    tokenizer_max_length: int = 64
    pad_to_max_length: bool = False
    tokenizer_padding_side: str = "right"
    image_key_order: Tuple[str, ...] = ("images", "pixel_values", "pixel_values_vit")
    # ---- B200 engine options (not in the reference; defaults keep reference behaviour) ----
    compute_dtype: str = "float32"     # reference loads fp32 (:187); "bfloat16" = throughput mode
    image_token_mode: str = "none"     # "none": prompt as tokenised (reference-literal);
                                       # "prefix": one image placeholder first => T' = n_img + T_text
    pool_merged_last: bool = False     # prefix mode: pool the last valid MERGED position instead of text_len-1 (F5)
    vision_chunk: int = 0              # images per FastViTHD pass (0 = engine default)
    skip_unused_vision: bool = True    # do not run the tower when no prompt holds a placeholder
    synthetic_seed: int = 0
    image_input_scale: float = 1.0     # pixel scale applied by the ingest kernel (1/255: raw uint8 frames)


def resize_with_pad(img: Tensor, width: int, height: int, pad_value: float = 0.0) -> Tensor:
    """Letterbox on the GPU: aspect-preserving bilinear resize, pad left/top (reference :36-55)."""
    if img.ndim != 4:
        raise ValueError(f"(B,C,H,W) expected, but got shape {tuple(img.shape)}")
    if width != height:
        raise ValueError("the native letterbox kernel produces square outputs")
    if not img.is_cuda:
        N.require_cuda()
        img = img.cuda()
    out = N.op_preprocess(img.contiguous(), int(width), torch.float32, nhwc=False, letterbox=True,
                          pad_value=float(pad_value))
    return out[..., :3].permute(0, 3, 1, 2)


class FastVLMBackbone(nn.Module):
    """LLaVA/FastVLM VLM used as a feature extractor; always feeds the tower (B,3,S,S)."""

    def __init__(self, config: FastVLMBackboneConfig | None = None) -> None:
        super().__init__()
        self.config = config or FastVLMBackboneConfig()
        N.require_cuda()  # fail loudly: no CPU path
        if self.config.compute_dtype not in _DTYPES:
            raise ValueError(f"compute_dtype must be one of {sorted(_DTYPES)}")
        if self.config.image_token_mode not in ("none", "prefix"):
            raise ValueError("image_token_mode must be 'none' or 'prefix'")

        self.model = self._load_model()
        if hasattr(self.model, "config"):
            self.model.config.output_hidden_states = True

        hidden_size = getattr(self.model.config, "hidden_size", None)
        if hidden_size is None:
            hs_list = getattr(self.model.config, "hidden_sizes", None)
            if isinstance(hs_list, (list, tuple)) and len(hs_list) > 0:
                hidden_size = int(hs_list[-1])
        if hidden_size is None:
            raise ValueError("Could not infer hidden size from model config.")
        self.output_dim = int(hidden_size)

        self.processor = None
        self.image_processor = None  # resizing is done by the engine's ingest kernel
        self.tokenizer = self._load_tokenizer()
        try:
            self.tokenizer.padding_side = self.config.tokenizer_padding_side
        except Exception:
            pass

        self.expected_size = self._resolve_expected_image_size()
        declared_size, tower_name = self._resolve_declared_tower_size()
        if (
            declared_size is not None
            and self.config.force_image_size is not None
            and int(self.expected_size) < int(declared_size)
        ):
            raise ValueError(
                "Configured image_size is too small for this FastVLM vision tower. "
                f"force_image_size={self.expected_size}, tower={tower_name}, required>={declared_size}. "
                "Set image_size to the declared tower size (e.g. 1024) or leave it unset (None) for auto-detection."
            )
        if int(self.expected_size) != int(self.model.arch.vision.image_size):
            # the CUDA tower is specialised to one input size; make it the expected one
            import dataclasses

            arch = self.model.arch
            if int(self.expected_size) % arch.vision.total_stride != 0:
                raise ValueError(f"image_size {self.expected_size} is not a multiple of the tower stride "
                                 f"{arch.vision.total_stride}")
            self.model.arch = dataclasses.replace(
                arch, vision=dataclasses.replace(arch.vision, image_size=int(self.expected_size)))

        self.model.configure_engine(pool_mode=self.config.image_feature_pool,
                                    vision_chunk=self.config.vision_chunk,
                                    skip_unused_vision=self.config.skip_unused_vision)
        if not self.config.freeze_backbone:
            # The reference cannot train the backbone either: its forward is @torch.no_grad (:501).  Say so once.
            logger.warning("freeze_backbone=False has no effect: the backbone forward runs without autograd "
                           "(as in the reference, fastvlm_adapter.py:501); only the action head is trainable.")
        for p in self.model.parameters():
            p.requires_grad = False
        self._token_cache: Dict[Tuple[str, ...], Tuple[Tensor, Tensor]] = {}
        print(f"[FastVLMBackbone] expected (S,S) = ({self.expected_size},{self.expected_size})")

    # -------------------- loading --------------------
    def _model_kwargs(self) -> dict[str, Any]:
        return {"compute_dtype": _DTYPES[self.config.compute_dtype], "seed": self.config.synthetic_seed}

    def _load_model(self) -> nn.Module:
        """Load the backbone; local llava_qwen2 checkpoints that the generic path rejects go through
        the bootstrap fallback (reference :183-201)."""
        model_kwargs = self._model_kwargs()
        try:
            return LlavaQwen2Native.from_pretrained(self.config.model_id, **model_kwargs)
        except ValueError as err:
            if not self._needs_llava_qwen2_bootstrap(err):
                raise
            logger.warning(
                "Falling back to llava_qwen2 bootstrap loader for local checkpoint '%s'. "
                "Using bootstrap model config from '%s'.",
                self.config.model_id,
                self.config.bootstrap_model_id,
            )
            return self._load_llava_qwen2_with_bootstrap(model_kwargs)

    @staticmethod
    def _needs_llava_qwen2_bootstrap(err: Exception) -> bool:
        message = str(err)
        return "model type `llava_qwen2`" in message and "does not recognize this architecture" in message

    def _load_llava_qwen2_with_bootstrap(self, model_kwargs: dict[str, Any]) -> nn.Module:
        """Local llava_qwen2 directory + architecture defaults from the bootstrap model
        (reference :208-241; same preconditions and error texts)."""
        model_path = Path(self.config.model_id)
        config_path = model_path / "config.json"
        if not model_path.is_dir() or not config_path.is_file():
            raise RuntimeError(
                "llava_qwen2 bootstrap fallback only supports local checkpoint directories containing config.json. "
                f"Got model_id='{self.config.model_id}'."
            )
        with open(config_path, encoding="utf-8") as f:
            local_config = json.load(f)
        if local_config.get("model_type") != "llava_qwen2":
            raise RuntimeError(
                "Bootstrap fallback was triggered, but the local model_type is not llava_qwen2. "
                f"Got '{local_config.get('model_type')}'."
            )
        try:
            from .arch import arch_from_hf_config, load_arch
            from .llava_qwen2 import _read_checkpoint_tensors

            boot = load_arch(self.config.bootstrap_model_id.rsplit("/", 1)[-1].lower())  # "apple/FastVLM-0.5B" -> preset
            if boot is None:
                boot = load_arch(self.config.bootstrap_model_id)  # a local directory / "synthetic:<preset>"
            merged = dict(local_config)
            if boot is not None:
                from dataclasses import asdict

                merged.setdefault("mm_vision_tower", boot.mm_vision_tower)
                geometry = asdict(boot.vision)
                geometry.pop("image_size")  # follows the tower name / force_image_size, like a described checkpoint
                merged.setdefault("vision_arch", geometry)
            if boot is None and "mm_vision_tower" not in merged:
                raise ValueError(f"unknown bootstrap model {self.config.bootstrap_model_id!r}")
            arch = arch_from_hf_config(merged)
            return LlavaQwen2Native(arch, _read_checkpoint_tensors(model_path), self.config.model_id,
                                    model_kwargs["compute_dtype"])
        except Exception as exc:
            raise RuntimeError(
                "Failed to load local llava_qwen2 checkpoint with bootstrap config. "
                f"model_id='{self.config.model_id}', bootstrap_model_id='{self.config.bootstrap_model_id}'."
            ) from exc

    def _load_tokenizer(self):
        if self.config.model_id.startswith(SYNTHETIC_PREFIX):
            return SimpleByteTokenizer(self.model.arch.text.vocab, padding_side=self.config.tokenizer_padding_side)
        from transformers import AutoTokenizer

        return AutoTokenizer.from_pretrained(self.config.model_id, trust_remote_code=False)

    # -------------------- helpers --------------------
    def _resolve_expected_image_size(self) -> int:
        """force -> vision_config.image_size -> size in the tower name -> fallback (reference :245-278)."""
        if self.config.force_image_size is not None:
            return int(self.config.force_image_size)
        cfg = getattr(self.model, "config", None)
        if cfg is not None:
            vcfg = getattr(cfg, "vision_config", None)
            if vcfg is not None:
                img_size = getattr(vcfg, "image_size", None)
                if isinstance(img_size, (int, float)):
                    return int(img_size)
                if isinstance(img_size, (tuple, list)) and len(img_size) > 0:
                    return int(img_size[0])
            tower_size, _ = self._resolve_declared_tower_size()
            if tower_size is not None:
                return int(tower_size)
        return int(self.config.fallback_image_size)

    def _resolve_declared_tower_size(self) -> tuple[Optional[int], Optional[str]]:
        cfg = getattr(self.model, "config", None)
        if cfg is None:
            return None, None
        candidates = [getattr(cfg, "mm_vision_tower", None), getattr(cfg, "vision_tower", None)]
        vcfg = getattr(cfg, "vision_config", None)
        if vcfg is not None:
            candidates += [getattr(vcfg, "model_name", None), getattr(vcfg, "name_or_path", None)]
        for tower_name in candidates:
            tower_size = self._infer_size_from_tower_name(tower_name)
            if tower_size is not None:
                return tower_size, str(tower_name)
        return None, None

    @staticmethod
    def _infer_size_from_tower_name(tower_name: Any) -> Optional[int]:
        """`mobileclip_l_1024` -> 1024, `...patch14-384` -> 384; ignores scale suffixes such as
        `so400m` (same accept/reject rules as reference :300-335)."""
        if not isinstance(tower_name, str):
            return None
        name = tower_name.lower()
        plausible = lambda v: 64 <= v <= 4096  # noqa: E731
        for pattern in (r"(?:^|[_-])(\d{2,4})$", r"patch\d+[-_](\d{2,4})(?:$|[_-])"):
            m = re.search(pattern, name)
            if m is not None and plausible(int(m.group(1))):
                return int(m.group(1))
        picked = None
        for m in re.finditer(r"(\d{2,4})", name):
            if plausible(int(m.group(1))) and name[m.end(): m.end() + 1] not in {"m", "b"}:
                picked = int(m.group(1))
        return picked

    @staticmethod
    def _pool_hidden(hidden: Tensor, attention_mask: Optional[Tensor], mode: str) -> Tensor:
        """Kept for API compatibility (reference :337-359).  The engine pools in-kernel; this helper is
        only for callers that hold a (B,T,H) tensor of their own."""
        if mode == "mean_pool":
            if attention_mask is None:
                return hidden.mean(dim=1)
            mask = attention_mask.float().unsqueeze(-1)
            return (hidden * mask).sum(dim=1) / mask.sum(dim=1).clamp_min(1e-6)
        if attention_mask is not None:
            idx = (attention_mask.long().sum(dim=1) - 1).clamp_min(0)
            b, _, h = hidden.size()
            return hidden.gather(dim=1, index=idx.view(b, 1, 1).expand(b, 1, h)).squeeze(1)
        return hidden[:, -1, :]

    def _tokenize(self, tasks: List[str]) -> Tuple[Tensor, Tensor]:
        """CPU (input_ids, attention_mask), cached per prompt tuple (the tokenizer is the only CPU work
        left on the path)."""
        if self.tokenizer is None:
            raise RuntimeError("Tokenizer is missing; ensure AutoTokenizer/AutoProcessor is available.")
        key = tuple(tasks)
        hit = self._token_cache.get(key)
        if hit is not None:
            return hit
        padding = "max_length" if self.config.pad_to_max_length else "longest"
        try:
            self.tokenizer.padding_side = self.config.tokenizer_padding_side
        except Exception:
            pass
        tok = self.tokenizer(list(tasks), padding=padding, truncation=True,
                             max_length=self.config.tokenizer_max_length, return_tensors="pt")
        ids, mask = tok["input_ids"].cpu(), tok["attention_mask"].cpu()
        if len(self._token_cache) > 4096:
            self._token_cache.clear()
        self._token_cache[key] = (ids, mask)
        return ids, mask

    def _prep_text(self, tasks: List[str], device: torch.device) -> Dict[str, Tensor]:
        ids, mask = self._tokenize(tasks)
        return {"input_ids": ids.to(device), "attention_mask": mask.to(device)}

    # -------- image canonicalisation: always (B,3,S,S) --------
    def _as_bchw(self, images: Union[List[ImageLike], ImageLike, Tensor]) -> Tensor:
        """PIL / NumPy / Tensor (BCHW, BHWC, CHW, HWC) / list -> BCHW; values untouched.
        Same layout heuristics as reference :384-442 (4-D is BHWC only if last dim in {1,3} and dim 1 not)."""
        x, nhwc = self._as_batch(images)
        return x.permute(0, 3, 1, 2) if nhwc else x

    def _as_batch(self, images) -> Tuple[Tensor, bool]:
        """-> (4-D tensor, is_nhwc) without copying device tensors."""
        try:
            import numpy as np
        except Exception:  # pragma: no cover
            np = None
        try:
            from PIL import Image as PILImage  # type: ignore
        except Exception:  # pragma: no cover
            PILImage = None

        def one(x) -> Tuple[Tensor, bool]:  # -> (3-D tensor, is_hwc)
            if np is not None and isinstance(x, np.ndarray):
                if x.ndim == 3 and (x.shape[0] in (1, 3) or x.shape[-1] in (1, 3)):
                    x = torch.from_numpy(x)
                else:
                    raise ValueError(f"Unsupported numpy array shape: {x.shape}")
            elif PILImage is not None and isinstance(x, PILImage.Image):
                arr = torch.from_numpy(np.array(x))
                if arr.ndim == 3 and arr.shape[-1] in (1, 3):
                    return arr, True
                raise ValueError(f"Unsupported PIL image shape: {tuple(arr.shape)}")
            if isinstance(x, torch.Tensor):
                if x.ndim == 3:
                    return (x, False) if x.shape[0] in (1, 3) else (x, True)
                if x.ndim == 2:
                    return x.unsqueeze(0), False
                raise ValueError(f"Unsupported tensor shape: {tuple(x.shape)}")
            raise TypeError(f"Unsupported image type: {type(x)}")

        if isinstance(images, torch.Tensor) and images.ndim == 4:
            nhwc = images.shape[-1] in (1, 3) and images.shape[1] not in (1, 3)
            return images, bool(nhwc)
        if isinstance(images, (list, tuple)):
            items = [one(i) for i in images]
            chw = [t.permute(2, 0, 1) if hwc else t for t, hwc in items]
            return torch.stack([c.to(torch.float32) for c in chw], dim=0), False
        t, hwc = one(images)
        return t.unsqueeze(0), hwc

    def _normalize_channels(self, x_bchw: Tensor) -> Tensor:
        if x_bchw.shape[1] == 1:
            return x_bchw.repeat(1, 3, 1, 1)
        if x_bchw.shape[1] > 3:
            return x_bchw[:, :3]
        return x_bchw

    def _ingest_args(self, x: Tensor) -> Dict[str, Any]:
        """Kernel arguments equivalent to `_resize_image` + `_maybe_normalize_imagenet` (:451-477)."""
        args: Dict[str, Any] = dict(letterbox=bool(self.config.resize_with_padding),
                                    pad_value=float(self.config.pad_value),
                                    img_scale=float(self.config.image_input_scale), mean=None, std=None)
        if self.config.normalize_imagenet:
            # 0..255 input (reference :472-473).  uint8 frames answer this from the dtype; a float batch needs its
            # maximum, i.e. one device->host sync per call — only on this opt-in path.
            if self.config.image_input_scale == 1.0 and (x.dtype == torch.uint8 or float(x.max()) > 1.5):
                args["img_scale"] = 1.0 / 255.0
            args["mean"], args["std"] = _IMAGENET_MEAN, _IMAGENET_STD
        return args

    @staticmethod
    def _device_image(x: Tensor, device: torch.device) -> Tensor:
        if x.dtype not in (torch.float32, torch.bfloat16, torch.uint8):
            x = x.to(torch.float32)
        return x.to(device, non_blocking=True).contiguous()

    def _prepare_images_tensor(self, images: Union[List[ImageLike], ImageLike, Tensor], device: torch.device) -> Tensor:
        """-> (B,3,S,S) float32 on `device`, computed by the ingest kernel (reference :479-488)."""
        S = int(self.expected_size)
        x, nhwc = self._as_batch(images)
        x = self._device_image(x, torch.device(device))
        a = self._ingest_args(x)
        out = N.op_preprocess(x, S, torch.float32, nhwc=nhwc, letterbox=a["letterbox"], pad_value=a["pad_value"],
                              scale=a["img_scale"], mean=a["mean"], std=a["std"])
        return out[..., :3].permute(0, 3, 1, 2)

    def _pack_image_inputs(self, x_bchw: Tensor, prefer_keys: Tuple[str, ...]) -> List[Dict[str, Tensor]]:
        return [{k: x_bchw} for k in prefer_keys]

    # -------------------- forward --------------------
    def _prompt_ids(self, tasks: List[str]) -> Tuple[Tensor, Tensor, Optional[Tensor]]:
        """(token ids with optional image placeholder, text_len, pool_idx or None) on the CPU."""
        if self.config.tokenizer_padding_side != "right":
            raise ValueError("the native engine supports tokenizer_padding_side='right' only")
        ids, mask = self._tokenize(tasks)
        lens = mask.long().sum(dim=1)
        pool_idx = None
        if self.config.image_token_mode == "prefix":
            ph = torch.full((ids.shape[0], 1), IMAGE_TOKEN_INDEX, dtype=ids.dtype)
            ids = torch.cat([ph, ids], dim=1)
            lens = lens + 1
            if self.config.pool_merged_last:
                pool_idx = lens - 1 + (self.model.arch.vision.num_tokens - 1)
        return ids, lens, pool_idx

    @torch.no_grad()
    def forward(self, images: Union[List[ImageLike], ImageLike, Tensor], tasks: List[str],
                device: torch.device | None = None) -> Tensor:  # (B, H=self.output_dim)
        return self._run(images, tasks, None, device)

    def _run(self, images, tasks: List[str], states: Optional[Tensor], device: torch.device | None) -> Tensor:
        # The engine lives where the module lives (`policy.to("cuda:1")` moves it); `device` is accepted for API
        # compatibility (reference :501-513 moves its inputs there) but cannot disagree with the module.
        home = self.model.device
        if home.type != "cuda":
            raise N.NativeError("the FastVLA B200 path runs on CUDA only (no CPU fallback): move the policy to a "
                                "CUDA device")
        if device is not None:
            device = torch.device(device)
            if device.type != "cuda":
                raise N.NativeError("the FastVLA B200 path runs on CUDA only (no CPU fallback)")
            if device.index is not None and home.index is not None and device.index != home.index:
                raise ValueError(f"device={device} but the backbone lives on {home}; move the policy with .to(device)")
        device = home
        x, nhwc = self._as_batch(images)
        x = self._device_image(x, device)
        if len(tasks) != x.shape[0]:
            raise ValueError(f"got {x.shape[0]} images but {len(tasks)} task strings")
        ids, lens, pool_idx = self._prompt_ids(list(tasks))
        a = self._ingest_args(x)
        return self.model.engine.forward(x, ids, lens, states=states, pool_idx=pool_idx, nhwc=nhwc,
                                         letterbox=a["letterbox"], pad_value=a["pad_value"],
                                         img_scale=a["img_scale"], mean=a["mean"], std=a["std"])

    # compat: old call style `self.backbone(images, tasks, device=...)`
    def backbone(self, images, tasks, device: Optional[torch.device] = None, **kwargs):
        return self.forward(images, tasks, device=device)
