"""Python handle on the native FastVLA engine (libfvla.so).

Host-side responsibilities only: hand the state dict over by name, keep the handle alive, turn
torch tensors into raw pointers for `fvla_forward`.  All arithmetic happens in the CUDA library.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Sequence

import torch

from .. import _native as N
from .arch import BackboneArch

BACKBONE_KEY_PREFIX = "backbone.model."  # FastVLMWithExpert.backbone (FastVLMBackbone) .model (HF module)


def make_config(arch: BackboneArch, dtype: torch.dtype, state_dim: int, action_dim: int, hidden_dim: int,
                fusion_dim: int, pool_mode: str = "last_token", vision_chunk: int = 0,
                skip_unused_vision: bool = True) -> N.FvlaConfig:
    v, t = arch.vision, arch.text
    if len(v.dims) > N.MAX_VIS_STAGES:
        raise ValueError("too many vision stages")
    if pool_mode not in ("last_token", "mean_pool"):
        raise ValueError(f"unknown image_feature_pool mode {pool_mode!r}")
    cfg = N.FvlaConfig()
    cfg.dtype = N.dtype_code(dtype)
    cfg.image_size = v.image_size
    cfg.vis_num_stages = len(v.dims)
    for i in range(len(v.dims)):
        cfg.vis_layers[i] = v.layers[i]
        cfg.vis_dims[i] = v.dims[i]
        cfg.vis_attention[i] = int(v.attention[i])
        cfg.vis_pos_emb[i] = int(v.pos_emb[i])
    cfg.vis_mlp_ratio = v.mlp_ratio
    cfg.vis_head_dim = v.head_dim
    cfg.vis_se_reduced = v.se_reduced
    cfg.hidden, cfg.n_layers = t.hidden, t.layers
    cfg.n_q_heads, cfg.n_kv_heads, cfg.head_dim = t.q_heads, t.kv_heads, t.head_dim
    cfg.intermediate, cfg.vocab = t.intermediate, t.vocab
    cfg.rms_eps, cfg.rope_theta = t.rms_eps, t.rope_theta
    cfg.state_dim, cfg.action_dim = state_dim, action_dim
    cfg.hidden_dim, cfg.fusion_dim = hidden_dim, fusion_dim
    cfg.pool_mode = N.POOL_LAST_TOKEN if pool_mode == "last_token" else N.POOL_MEAN
    cfg.vision_chunk = vision_chunk
    cfg.skip_unused_vision = int(skip_unused_vision)
    return cfg


class NativeEngine:
    """One engine per process/device, like one policy instance in the reference."""

    def __init__(self, arch: BackboneArch, dtype: torch.dtype = torch.bfloat16, state_dim: int = 14,
                 action_dim: int = 14, hidden_dim: int = 1024, fusion_dim: int = 1024,
                 pool_mode: str = "last_token", vision_chunk: int = 0, skip_unused_vision: bool = True,
                 device: Optional[torch.device] = None) -> None:
        N.require_cuda()
        self.lib = N.load()
        self.arch = arch
        self.dtype = dtype
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        if self.device.type != "cuda":
            raise N.NativeError(f"the engine needs a CUDA device, got {self.device}")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.state_dim, self.action_dim = state_dim, action_dim
        self.hidden_dim, self.fusion_dim = hidden_dim, fusion_dim
        self.cfg = make_config(arch, dtype, state_dim, action_dim, hidden_dim, fusion_dim, pool_mode,
                               vision_chunk, skip_unused_vision)
        self._h = C.c_void_p()
        with torch.cuda.device(self.device):
            N.check(self.lib.fvla_create(C.byref(self.cfg), C.byref(self._h)), "fvla_create")
        self.finalized = False
        self._taps: Dict[int, torch.Tensor] = {}

    def __del__(self) -> None:
        try:
            if getattr(self, "_h", None) is not None and self._h.value:
                self.lib.fvla_destroy(self._h)
                self._h = C.c_void_p()
        except Exception:
            pass

    # ---- weights ------------------------------------------------------------------------------
    def load_tensor(self, name: str, t: torch.Tensor) -> None:
        t = t.detach()
        if t.dtype not in (torch.float32, torch.bfloat16):
            t = t.to(torch.float32)
        t = t.cpu().contiguous()
        shape = (C.c_int64 * max(t.dim(), 1))(*t.shape)
        N.check(self.lib.fvla_load_tensor(self._h, name.encode(), t.data_ptr(), N.dtype_code(t.dtype), t.dim(),
                                          shape), f"fvla_load_tensor({name})")

    def load_state_dict(self, sd: Dict[str, torch.Tensor], prefix: str = "") -> None:
        for k, v in sd.items():
            self.load_tensor(prefix + k, v)

    def missing_tensors(self) -> List[str]:
        n = C.c_int32(0)
        buf = C.create_string_buffer(1 << 20)
        N.check(self.lib.fvla_missing_tensors(self._h, buf, len(buf), C.byref(n)), "fvla_missing_tensors")
        names = buf.value.decode().split("\n")
        return [x for x in names if x]

    def finalize(self) -> None:
        with torch.cuda.device(self.device):
            N.check(self.lib.fvla_finalize(self._h), "fvla_finalize")
        self.finalized = True

    def reserve(self, batch: int, n_tokens: int) -> None:
        with torch.cuda.device(self.device):
            N.check(self.lib.fvla_reserve(self._h, batch, n_tokens), "fvla_reserve")

    # ---- LeRobot (un)normaliser steps fused into the head kernel -------------------------------
    def set_io_normalization(self, state_mean=None, state_std=None, action_mean=None, action_std=None,
                             eps: float = 1e-8) -> None:
        """state' = (state - mean) / (std + eps) in front of the head, action' = action * std + mean behind it — the
        MEAN_STD arithmetic of LeRobot's Normalizer / Unnormalizer steps
        (lerobot_fastvla/processor_fastvla.py:30-48).  None = identity for that part."""
        if not self.finalized:
            raise N.NativeError("engine not finalized")

        def vec(t, n, fn):
            if t is None:
                return None
            v = fn(torch.as_tensor(t, dtype=torch.float64).flatten().cpu()).to(torch.float32).contiguous()
            if v.numel() != n:
                raise N.NativeError(f"normalisation vector has {v.numel()} elements, expected {n}")
            return v

        keep = [vec(state_mean, self.state_dim, lambda m: m),
                vec(state_std, self.state_dim, lambda s: 1.0 / (s + eps)),
                vec(action_std, self.action_dim, lambda s: s),
                vec(action_mean, self.action_dim, lambda m: m)]
        if state_mean is None and state_std is not None:
            raise N.NativeError("state_std without state_mean")
        ptrs = [C.cast(v.data_ptr(), C.POINTER(C.c_float)) if v is not None else None for v in keep]
        with torch.cuda.device(self.device):
            N.check(self.lib.fvla_set_io_normalization(self._h, *ptrs), "fvla_set_io_normalization")

    # ---- taps ---------------------------------------------------------------------------------
    def set_tap(self, stage: int, buf: Optional[torch.Tensor]) -> None:
        if buf is None:
            self._taps.pop(stage, None)
            N.check(self.lib.fvla_set_tap(self._h, stage, None, 0), "fvla_set_tap")
            return
        self._taps[stage] = buf
        N.check(self.lib.fvla_set_tap(self._h, stage, buf.data_ptr(), buf.numel() * buf.element_size()),
                "fvla_set_tap")

    # ---- forward ------------------------------------------------------------------------------
    def forward(self, images: torch.Tensor, token_ids: torch.Tensor, text_len: torch.Tensor,
                states: Optional[torch.Tensor] = None, pool_idx: Optional[torch.Tensor] = None,
                nhwc: bool = False, letterbox: bool = True, pad_value: float = 0.0, img_scale: float = 1.0,
                mean: Optional[Sequence[float]] = None, std: Optional[Sequence[float]] = None,
                out: Optional[torch.Tensor] = None, pooled_out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """images: CUDA (B,C,h,w) or (B,h,w,C) fp32/bf16/uint8.  token_ids: CPU int (B,T) with -200 image
        placeholders, right padded.  text_len: CPU int (B).  Returns actions (B,A) fp32 (or pooled
        features (B,H) when `states` is None)."""
        if not self.finalized:
            raise N.NativeError("engine not finalized")
        if not images.is_cuda or images.dim() != 4:
            raise N.NativeError("images must be a 4-D CUDA tensor")
        if images.device != self.device:
            raise N.NativeError(f"images live on {images.device} but this engine was built on {self.device}: one "
                                "engine per device (move the policy, or the batch)")
        images = images.contiguous()
        B = images.shape[0]
        if nhwc:
            _, h, w, c = images.shape
        else:
            _, c, h, w = images.shape
        ids = token_ids.to(dtype=torch.int32, device="cpu").contiguous()
        lens = text_len.to(dtype=torch.int32, device="cpu").contiguous()
        if ids.dim() != 2 or ids.shape[0] != B or lens.numel() != B:
            raise N.NativeError("token_ids must be (B,T) and text_len (B)")
        a = N.FvlaForwardArgs()
        a.batch = B
        a.images = images.data_ptr()
        a.img_dtype = N.dtype_code(images.dtype)
        a.img_nhwc = int(nhwc)
        a.img_c, a.img_h, a.img_w = c, h, w
        a.letterbox = int(letterbox)
        a.pad_value = pad_value
        a.img_scale = img_scale
        a.normalize = int(mean is not None)
        for i in range(3):
            a.mean[i] = float(mean[i]) if mean is not None else 0.0
            a.inv_std[i] = 1.0 / float(std[i]) if std is not None else 1.0
        a.token_ids = C.cast(ids.data_ptr(), C.POINTER(C.c_int32))
        a.text_len = C.cast(lens.data_ptr(), C.POINTER(C.c_int32))
        a.n_tokens = ids.shape[1]
        pidx = None
        if pool_idx is not None:
            pidx = pool_idx.to(dtype=torch.int32, device="cpu").contiguous()
            a.pool_idx = C.cast(pidx.data_ptr(), C.POINTER(C.c_int32))
        result: torch.Tensor
        if states is not None:
            st = states.to(device=self.device, dtype=torch.float32).contiguous()
            if st.shape != (B, self.state_dim):
                raise N.NativeError(f"states must be ({B},{self.state_dim}), got {tuple(st.shape)}")
            if out is None:
                out = torch.empty((B, self.action_dim), device=self.device, dtype=torch.float32)
            a.states = st.data_ptr()
            a.actions = out.data_ptr()
            result = out
        if pooled_out is None and states is None:
            pooled_out = torch.empty((B, self.arch.text.hidden), device=self.device, dtype=torch.float32)
        if pooled_out is not None:
            a.pooled = pooled_out.data_ptr()
        if states is None:
            result = pooled_out
        for name, t in (("out", out), ("pooled_out", pooled_out)):
            if t is not None and t.device != self.device:
                raise N.NativeError(f"{name} lives on {t.device}, engine on {self.device}")
        with torch.cuda.device(self.device):
            N.check(self.lib.fvla_forward(self._h, C.byref(a), N.stream_ptr()), "fvla_forward")
        return result

    # ---- profiling ----------------------------------------------------------------------------
    def set_profile(self, on: bool) -> None:
        N.check(self.lib.fvla_set_profile(self._h, int(on)), "fvla_set_profile")

    def profile_report(self) -> List[Dict[str, float]]:
        """Per-kernel/shape totals since the last report: label, count, total_ms, flops, bytes."""
        buf = C.create_string_buffer(1 << 20)
        N.check(self.lib.fvla_profile_report(self._h, buf, len(buf)), "fvla_profile_report")
        rows = []
        for line in buf.value.decode().strip().split("\n")[1:]:
            if not line:
                continue
            label, cnt, ms, fl, by = line.rsplit(",", 4)
            rows.append(dict(label=label, count=int(cnt), total_ms=float(ms), flops=float(fl), bytes=float(by)))
        return rows

    # ---- introspection ------------------------------------------------------------------------
    @property
    def last_launch_count(self) -> int:
        return int(self.lib.fvla_last_launch_count(self._h))

    @property
    def last_forward_flops(self) -> float:
        return float(self.lib.fvla_last_forward_flops(self._h))

    @property
    def merged_len(self) -> int:
        return int(self.lib.fvla_merged_len(self._h))

    @property
    def workspace_bytes(self) -> int:
        return int(self.lib.fvla_workspace_bytes(self._h))

    @property
    def weight_bytes(self) -> int:
        return int(self.lib.fvla_weight_bytes(self._h))
