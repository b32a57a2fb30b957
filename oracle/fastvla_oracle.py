"""CPU fp32 ORACLE for the FastVLA policy forward — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of
`bench.py` may import this module, and only as the checker / the CPU baseline.  The product path
(`vla-from-fastvlm_b200/`) never imports it and has no CPU fallback.

PARITY STATUS — "parity unpinned" for three of the four stages: the reference
(syun88/VLA-from-FastVLM) ships no tests, golden vectors or fixtures, and the
arithmetic of FastViTHD, the mlp2x_gelu projector and the LLaVA splice lives in the HF-hub remote
code of `apple/FastVLM-0.5B` (pulled with trust_remote_code=True and no revision pin at
src/vla_fastvlm/model/fastvlm_adapter.py:191), which is absent from /root/reference and cannot be
fetched offline.  Those stages are restated here from the published architecture (SURVEY.md
App. A-C).  What IS pinned:
  * the in-tree parts (image canonicalisation, pooling, action head, batch handling) are checked
    bit-for-bit against the real reference classes imported from /root/reference
    (tests/test_oracle_vs_reference.py; fixtures in tests/golden/ made by tests/golden/make_golden.py),
  * the Qwen2 decoder restatement is checked against the `transformers` Qwen2Model installed in the
    image (tests/test_oracle_qwen2.py).

Every function cites the reference file:line (or the upstream [EXT] source) it restates.
Tensors are NCHW / (B,T,H) exactly as in the reference, weights are the un-folded checkpoint
tensors (BatchNorm, layer scale, separate q/k/v and gate/up), so the engine's folding and repacking
is checked against the original formulation.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

IMAGE_TOKEN_INDEX = -200  # LLaVA constants.py [EXT]

VIS = "model.vision_tower.vision_tower.model."
PROJ = "model.mm_projector."
LLM = "model."

Tensor = torch.Tensor
Taps = Dict[str, Tensor]


# ------------------------------------------------------------------------------------------------
# Image canonicalisation — src/vla_fastvlm/model/fastvlm_adapter.py:36-55, 444-488
# ------------------------------------------------------------------------------------------------
def resize_with_pad(img: Tensor, width: int, height: int, pad_value: float = 0.0) -> Tensor:
    """fastvlm_adapter.py:36-55 — aspect-preserving bilinear resize, then pad LEFT/TOP."""
    if img.ndim != 4:
        raise ValueError(f"(B,C,H,W) expected, but got shape {tuple(img.shape)}")
    cur_h, cur_w = img.shape[2:]
    ratio = max(cur_w / width, cur_h / height)
    rh, rw = int(cur_h / ratio), int(cur_w / ratio)
    resized = F.interpolate(img, size=(rh, rw), mode="bilinear", align_corners=False)
    return F.pad(resized, (max(0, int(width - rw)), 0, max(0, int(height - rh)), 0), value=pad_value)


def prepare_images(images: Tensor, S: int, resize_with_padding: bool = True, pad_value: float = 0.0,
                   normalize_imagenet: bool = False) -> Tensor:
    """fastvlm_adapter.py:479-488 for tensor inputs (`_as_bchw` :423-431, `_normalize_channels`
    :444-449, `_resize_image` :451-461, `_maybe_normalize_imagenet` :463-477). -> (B,3,S,S) fp32."""
    x = images
    if x.ndim == 3:
        x = (x if x.shape[0] in (1, 3) else x.permute(2, 0, 1)).unsqueeze(0)
    elif x.ndim == 4 and x.shape[-1] in (1, 3) and x.shape[1] not in (1, 3):
        x = x.permute(0, 3, 1, 2)
    x = x.to(dtype=torch.float32, device="cpu")
    if x.shape[1] == 1:
        x = x.repeat(1, 3, 1, 1)
    elif x.shape[1] > 3:
        x = x[:, :3]
    if resize_with_padding:
        x = resize_with_pad(x, width=S, height=S, pad_value=pad_value)
    elif x.shape[-2:] != (S, S):
        x = F.interpolate(x, size=(S, S), mode="bilinear", align_corners=False)
    if normalize_imagenet:
        if x.max() > 1.5:
            x = x / 255.0
        mean = torch.tensor([0.485, 0.456, 0.406]).view(1, 3, 1, 1)
        std = torch.tensor([0.229, 0.224, 0.225]).view(1, 3, 1, 1)
        x = (x - mean) / std
    return x


# ------------------------------------------------------------------------------------------------
# FastViTHD — [EXT] apple/ml-fastvlm llava/model/multimodal_encoder/mobileclip/mci.py,
# inference-mode (reparameterised) form; SURVEY.md App. A
# ------------------------------------------------------------------------------------------------
def _bn(sd: Dict[str, Tensor], p: str, x: Tensor) -> Tensor:
    """nn.BatchNorm2d in eval mode (eps 1e-5)."""
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"],
                        sd[p + ".bias"], training=False, eps=1e-5)


def _ln_channel(sd: Dict[str, Tensor], p: str, x: Tensor) -> Tensor:
    """mci.py LayerNormChannel [EXT]: per-pixel statistics over the channel dimension of (B,C,H,W), biased variance,
    eps 1e-5, per-channel weight and bias."""
    u = x.mean(1, keepdim=True)
    s = (x - u).pow(2).mean(1, keepdim=True)
    x = (x - u) / torch.sqrt(s + 1e-5)
    return sd[p + ".weight"].view(1, -1, 1, 1) * x + sd[p + ".bias"].view(1, -1, 1, 1)


def _attn_norm(sd: Dict[str, Tensor], p: str, x: Tensor) -> Tensor:
    """AttentionBlock.norm: the checkpoint's key set says which layer it is — BatchNorm2d carries running statistics,
    LayerNormChannel does not (SURVEY App. A marks the choice unverified; both are restated)."""
    return _bn(sd, p, x) if p + ".running_mean" in sd else _ln_channel(sd, p, x)


def _conv_ffn(sd: Dict[str, Tensor], p: str, x: Tensor) -> Tensor:
    """mci.py ConvFFN.forward: dw7x7(no bias) -> BN -> fc1 -> GELU -> fc2."""
    d = x.shape[1]
    x = F.conv2d(x, sd[p + ".conv.conv.weight"], None, padding=3, groups=d)
    x = _bn(sd, p + ".conv.bn", x)
    x = F.gelu(F.conv2d(x, sd[p + ".fc1.weight"], sd[p + ".fc1.bias"]))
    return F.conv2d(x, sd[p + ".fc2.weight"], sd[p + ".fc2.bias"])


def _mhsa(sd: Dict[str, Tensor], p: str, x: Tensor, head_dim: int) -> Tensor:
    """mci.py MHSA.forward: qkv Linear (no bias), heads of `head_dim`, softmax(q*scale @ k^T) @ v,
    proj Linear."""
    B, C, H, W = x.shape
    N = H * W
    heads = C // head_dim
    t = x.flatten(2).transpose(-2, -1)  # (B, N, C)
    qkv = F.linear(t, sd[p + ".qkv.weight"]).reshape(B, N, 3, heads, head_dim).permute(2, 0, 3, 1, 4)
    q, k, v = qkv.unbind(0)
    attn = (q * head_dim ** -0.5) @ k.transpose(-2, -1)
    attn = attn.softmax(dim=-1)
    t = (attn @ v).transpose(1, 2).reshape(B, N, C)
    t = F.linear(t, sd[p + ".proj.weight"], sd[p + ".proj.bias"])
    return t.transpose(-2, -1).reshape(B, C, H, W)


def fastvithd_forward(sd: Dict[str, Tensor], x: Tensor, layers: Sequence[int], dims: Sequence[int],
                      attention: Sequence[bool], pos_emb: Sequence[bool], head_dim: int = 32,
                      taps: Optional[Taps] = None) -> Tensor:
    """(B,3,S,S) -> (B, n_tokens, 2*dims[-1]) image features, as MobileCLIPVisionTower returns them
    (feature map of conv_exp, flattened; transformers fast_vlm modeling :143-146 does the same)."""
    P = VIS
    # stem: three MobileOneBlocks, each reparam_conv -> GELU (mci.py convolutional_stem)
    x = F.gelu(F.conv2d(x, sd[P + "patch_embed.0.reparam_conv.weight"], sd[P + "patch_embed.0.reparam_conv.bias"],
                        stride=2, padding=1))
    x = F.gelu(F.conv2d(x, sd[P + "patch_embed.1.reparam_conv.weight"], sd[P + "patch_embed.1.reparam_conv.bias"],
                        stride=2, padding=1, groups=dims[0]))
    x = F.gelu(F.conv2d(x, sd[P + "patch_embed.2.reparam_conv.weight"], sd[P + "patch_embed.2.reparam_conv.bias"]))
    if taps is not None:
        taps["stem"] = x
    idx = 0
    for i, d in enumerate(dims):
        if pos_emb[i]:  # RepCPE, reparameterised: one depthwise 7x7 (+ identity folded in)
            p = P + f"network.{idx}.reparam_conv"
            x = F.conv2d(x, sd[p + ".weight"], sd[p + ".bias"], padding=3, groups=d)
            idx += 1
        for j in range(layers[i]):
            p = P + f"network.{idx}.{j}"
            if attention[i]:
                # AttentionBlock: x + ls1*MHSA(BN(x)); x + ls2*ConvFFN(x)
                x = x + sd[p + ".layer_scale_1"] * _mhsa(sd, p + ".token_mixer", _attn_norm(sd, p + ".norm", x), head_dim)
                x = x + sd[p + ".layer_scale_2"] * _conv_ffn(sd, p + ".convffn", x)
            else:
                # RepMixerBlock: x = reparam dw3x3(x); x + ls*ConvFFN(x)
                x = F.conv2d(x, sd[p + ".token_mixer.reparam_conv.weight"],
                             sd[p + ".token_mixer.reparam_conv.bias"], padding=1, groups=d)
                x = x + sd[p + ".layer_scale"] * _conv_ffn(sd, p + ".convffn", x)
        idx += 1
        if taps is not None:
            taps[f"vis_stage{i}"] = x
        if i + 1 < len(dims):
            # PatchEmbed: ReparamLargeKernelConv 7x7 s2 groups=d (-> GELU), MobileOneBlock 1x1 (-> GELU)
            p = P + f"network.{idx}"
            x = F.gelu(F.conv2d(x, sd[p + ".proj.0.lkb_reparam.weight"], sd[p + ".proj.0.lkb_reparam.bias"],
                                stride=2, padding=3, groups=d))
            x = F.gelu(F.conv2d(x, sd[p + ".proj.1.reparam_conv.weight"], sd[p + ".proj.1.reparam_conv.bias"]))
            idx += 1
    # conv_exp: MobileOneBlock 3x3 groups=d_last, x2 channels, SE, GELU
    d = dims[-1]
    x = F.conv2d(x, sd[P + "conv_exp.reparam_conv.weight"], sd[P + "conv_exp.reparam_conv.bias"], padding=1, groups=d)
    # SEBlock: avg_pool2d over the (16x16 at 1024^2) final map == global mean; reduce->ReLU->expand->sigmoid
    s = x.mean(dim=(2, 3), keepdim=True)
    s = F.relu(F.conv2d(s, sd[P + "conv_exp.se.reduce.weight"], sd[P + "conv_exp.se.reduce.bias"]))
    s = torch.sigmoid(F.conv2d(s, sd[P + "conv_exp.se.expand.weight"], sd[P + "conv_exp.se.expand.bias"]))
    x = F.gelu(x * s)
    feats = x.flatten(2).transpose(1, 2)  # (B, HW, C)
    if taps is not None:
        taps["image_features"] = feats
    return feats


def mm_projector(sd: Dict[str, Tensor], feats: Tensor) -> Tensor:
    """mlp2x_gelu: Linear -> GELU(erf) -> Linear (LLaVA multimodal_projector/builder.py [EXT];
    transformers fast_vlm modeling :39-57 agrees)."""
    h = F.gelu(F.linear(feats, sd[PROJ + "0.weight"], sd[PROJ + "0.bias"]))
    return F.linear(h, sd[PROJ + "2.weight"], sd[PROJ + "2.bias"])


# ------------------------------------------------------------------------------------------------
# LLaVA splice — llava_arch.py prepare_inputs_labels_for_multimodal [EXT]; SURVEY.md App. C
# ------------------------------------------------------------------------------------------------
def splice_inputs(embed: Tensor, input_ids: Tensor, attention_mask: Tensor, image_tokens: Tensor,
                  padding_side: str = "right") -> Tuple[Tensor, Tensor]:
    """-> inputs_embeds (B,T',H), attention_mask (B,T') bool.  A sample without a placeholder gets
    `image_features[0:0]` appended, i.e. nothing."""
    B = input_ids.shape[0]
    rows: List[Tensor] = []
    cur_image = 0
    for b in range(B):
        ids = input_ids[b][attention_mask[b].bool()]
        n_img = int((ids == IMAGE_TOKEN_INDEX).sum())
        if n_img == 0:
            rows.append(torch.cat([F.embedding(ids, embed), image_tokens[cur_image][0:0]], dim=0))
            cur_image += 1
            continue
        pieces: List[Tensor] = []
        start = 0
        pos = torch.where(ids == IMAGE_TOKEN_INDEX)[0].tolist()
        for p in pos:
            pieces.append(F.embedding(ids[start:p], embed))
            pieces.append(image_tokens[cur_image])
            cur_image += 1
            start = p + 1
        pieces.append(F.embedding(ids[start:], embed))
        rows.append(torch.cat(pieces, dim=0))
    Tm = max(r.shape[0] for r in rows)
    H = embed.shape[1]
    out = torch.zeros(B, Tm, H, dtype=embed.dtype)
    mask = torch.zeros(B, Tm, dtype=torch.bool)
    for b, r in enumerate(rows):
        n = r.shape[0]
        if padding_side == "left":
            out[b, Tm - n:] = r
            mask[b, Tm - n:] = True
        else:
            out[b, :n] = r
            mask[b, :n] = True
    return out, mask


# ------------------------------------------------------------------------------------------------
# Qwen2 decoder — transformers/models/qwen2/modeling_qwen2.py (container copy, v5.5.0)
# ------------------------------------------------------------------------------------------------
def _rms_norm(x: Tensor, w: Tensor, eps: float) -> Tensor:
    """Qwen2RMSNorm.forward."""
    xf = x.to(torch.float32)
    xf = xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + eps)
    return w * xf.to(x.dtype)


def _rope(T: int, head_dim: int, theta: float) -> Tuple[Tensor, Tensor]:
    """Qwen2RotaryEmbedding (default rope init): inv_freq = theta^(-2i/d); emb = cat(freqs, freqs)."""
    inv_freq = 1.0 / (theta ** (torch.arange(0, head_dim, 2, dtype=torch.int64).float() / head_dim))
    freqs = torch.arange(T, dtype=torch.float32)[:, None] * inv_freq[None, :]
    emb = torch.cat([freqs, freqs], dim=-1)
    return emb.cos(), emb.sin()


def _rotate_half(x: Tensor) -> Tensor:
    x1, x2 = x[..., : x.shape[-1] // 2], x[..., x.shape[-1] // 2:]
    return torch.cat((-x2, x1), dim=-1)


def qwen2_forward(sd: Dict[str, Tensor], x: Tensor, mask: Tensor, n_layers: int, q_heads: int,
                  kv_heads: int, head_dim: int, eps: float, theta: float,
                  taps: Optional[Taps] = None) -> Tensor:
    """inputs_embeds (B,T,H) + bool padding mask (B,T) -> final-norm hidden states (B,T,H).
    Pre-norm GQA causal attention (q/k/v bias, rotate-half RoPE, fp32 softmax) + SwiGLU MLP."""
    B, T, H = x.shape
    cos, sin = _rope(T, head_dim, theta)
    causal = torch.ones(T, T, dtype=torch.bool).tril()
    allow = causal[None, None] & mask[:, None, None, :]  # (B,1,T,T)
    bias = torch.zeros(B, 1, T, T).masked_fill(~allow, float("-inf"))
    # rows that are pure padding would be all -inf; keep them finite (their outputs are never read)
    dead = ~allow.any(-1, keepdim=True)
    bias = bias.masked_fill(dead, 0.0)
    rep = q_heads // kv_heads
    for l in range(n_layers):
        p = LLM + f"layers.{l}"
        h = _rms_norm(x, sd[p + ".input_layernorm.weight"], eps)
        q = F.linear(h, sd[p + ".self_attn.q_proj.weight"], sd[p + ".self_attn.q_proj.bias"])
        k = F.linear(h, sd[p + ".self_attn.k_proj.weight"], sd[p + ".self_attn.k_proj.bias"])
        v = F.linear(h, sd[p + ".self_attn.v_proj.weight"], sd[p + ".self_attn.v_proj.bias"])
        q = q.view(B, T, q_heads, head_dim).transpose(1, 2)
        k = k.view(B, T, kv_heads, head_dim).transpose(1, 2)
        v = v.view(B, T, kv_heads, head_dim).transpose(1, 2)
        q = q * cos + _rotate_half(q) * sin
        k = k * cos + _rotate_half(k) * sin
        k = k.repeat_interleave(rep, dim=1)
        v = v.repeat_interleave(rep, dim=1)
        att = (q @ k.transpose(2, 3)) * head_dim ** -0.5 + bias
        att = att.softmax(dim=-1, dtype=torch.float32)
        a = (att @ v).transpose(1, 2).reshape(B, T, q_heads * head_dim)
        x = x + F.linear(a, sd[p + ".self_attn.o_proj.weight"])
        h = _rms_norm(x, sd[p + ".post_attention_layernorm.weight"], eps)
        g = F.silu(F.linear(h, sd[p + ".mlp.gate_proj.weight"])) * F.linear(h, sd[p + ".mlp.up_proj.weight"])
        x = x + F.linear(g, sd[p + ".mlp.down_proj.weight"])
        if taps is not None:
            taps[f"layer{l}"] = x
    return _rms_norm(x, sd[LLM + "norm.weight"], eps)


# ------------------------------------------------------------------------------------------------
# Pooling + action head — fastvlm_adapter.py:337-359, fastvla/fastvlm_with_expert.py:23-38,50-54
# ------------------------------------------------------------------------------------------------
def pool_hidden(hidden: Tensor, attention_mask: Optional[Tensor], mode: str) -> Tensor:
    """fastvlm_adapter.py:337-359.  NB the mask is the TEXT mask even when image tokens were
    spliced in."""
    if mode == "mean_pool":
        if attention_mask is None:
            return hidden.mean(dim=1)
        m = attention_mask.float().unsqueeze(-1)
        return (hidden[:, : m.shape[1]] * m).sum(dim=1) / m.sum(dim=1).clamp_min(1e-6)
    if attention_mask is not None:
        idx = (attention_mask.long().sum(dim=1) - 1).clamp_min(0)
        b, _, h = hidden.size()
        return hidden.gather(dim=1, index=idx.view(b, 1, 1).expand(b, 1, h)).squeeze(1)
    return hidden[:, -1, :]


def action_head(hsd: Dict[str, Tensor], pooled: Tensor, states: Tensor, taps: Optional[Taps] = None) -> Tensor:
    """fastvlm_with_expert.py:23-38, 50-54 in eval mode (Dropout = identity)."""
    s = F.layer_norm(states, (states.shape[-1],), hsd["state_projection.0.weight"], hsd["state_projection.0.bias"])
    s = F.silu(F.linear(s, hsd["state_projection.1.weight"], hsd["state_projection.1.bias"]))
    f = torch.cat([pooled, s], dim=-1)
    f = F.linear(f, hsd["fusion.0.weight"], hsd["fusion.0.bias"])
    f = F.silu(F.layer_norm(f, (f.shape[-1],), hsd["fusion.1.weight"], hsd["fusion.1.bias"]))
    f = F.silu(F.linear(f, hsd["fusion.4.weight"], hsd["fusion.4.bias"]))
    if taps is not None:
        taps["state_feat"] = s
        taps["fused"] = f
    return F.linear(f, hsd["action_head.weight"], hsd["action_head.bias"])


# ------------------------------------------------------------------------------------------------
# Whole path
# ------------------------------------------------------------------------------------------------
class FastVLAOracle:
    """obs images + token ids + state -> action, CPU fp32.  `arch` is a
    vla_fastvlm.model.arch.BackboneArch (only its plain fields are read)."""

    def __init__(self, arch, backbone_sd: Dict[str, Tensor], head_sd: Optional[Dict[str, Tensor]],
                 pool_mode: str = "last_token", resize_with_padding: bool = True, pad_value: float = 0.0):
        self.arch = arch
        self.sd = {k: v.detach().to(torch.float32).cpu() for k, v in backbone_sd.items()}
        self.hsd = None if head_sd is None else {k: v.detach().to(torch.float32).cpu() for k, v in head_sd.items()}
        self.pool_mode = pool_mode
        self.resize_with_padding = resize_with_padding
        self.pad_value = pad_value

    @torch.no_grad()
    def vlm_hidden(self, pixel: Tensor, input_ids: Tensor, attention_mask: Tensor,
                   taps: Optional[Taps] = None, inject: Optional[Taps] = None) -> Tuple[Tensor, Tensor]:
        """LlavaQwen2ForCausalLM.forward(input_ids, attention_mask, images=pixel) minus the LM head:
        returns hidden_states[-1] (B,T',H) and the merged mask.  `inject["image_features"]` (test hook) replaces
        the tower's output so that the segments after an ill-conditioned tower input can be checked on their own."""
        v, t = self.arch.vision, self.arch.text
        if inject is not None and "image_features" in inject:
            feats = inject["image_features"].to(torch.float32).cpu()
        else:
            feats = fastvithd_forward(self.sd, pixel, v.layers, v.dims, v.attention, v.pos_emb, v.head_dim, taps)
        img_tok = mm_projector(self.sd, feats)
        if taps is not None:
            taps["projector"] = img_tok
        embeds, mask = splice_inputs(self.sd[LLM + "embed_tokens.weight"], input_ids, attention_mask, img_tok,
                                     self.arch.tokenizer_padding_side)
        if taps is not None:
            taps["embeds"] = embeds
        hidden = qwen2_forward(self.sd, embeds, mask, t.layers, t.q_heads, t.kv_heads, t.head_dim, t.rms_eps,
                               t.rope_theta, taps)
        return hidden, mask

    @torch.no_grad()
    def forward(self, images: Tensor, states: Optional[Tensor], input_ids: Tensor, attention_mask: Tensor,
                pool_idx: Optional[Tensor] = None, taps: Optional[Taps] = None,
                inject: Optional[Taps] = None) -> Tensor:
        pixel = prepare_images(images, self.arch.vision.image_size, self.resize_with_padding, self.pad_value)
        if taps is not None:
            taps["preprocess"] = pixel
        hidden, _ = self.vlm_hidden(pixel, input_ids, attention_mask, taps, inject)
        if pool_idx is not None:
            b, _, h = hidden.shape
            pooled = hidden.gather(1, pool_idx.view(b, 1, 1).expand(b, 1, h).long()).squeeze(1)
        else:
            pooled = pool_hidden(hidden, attention_mask, self.pool_mode)
        if taps is not None:
            taps["pooled"] = pooled
        if states is None or self.hsd is None:
            return pooled
        return action_head(self.hsd, pooled, states.to(torch.float32).cpu(), taps)


def flops_per_sample(arch, t_merged: int, state_dim: int, action_dim: int, hidden_dim: int = 1024,
                     fusion_dim: int = 1024) -> float:
    """Algorithmic FLOPs of one sample: GEMMs + depthwise + attention, LM head excluded,
    causal attention counted as half the square."""
    v, t = arch.vision, arch.text
    S = v.image_size
    f = 2.0 * 27 * (S // 2) ** 2 * v.dims[0] + 2.0 * 9 * (S // 4) ** 2 * v.dims[0] + 2.0 * (S // 4) ** 2 * v.dims[0] ** 2
    side = S // 4
    for i, d in enumerate(v.dims):
        px = side * side
        if v.pos_emb[i]:
            f += 2.0 * 49 * px * d
        for _ in range(v.layers[i]):
            if v.attention[i]:
                f += 2.0 * px * d * 3 * d + 4.0 * px * px * d + 2.0 * px * d * d
            else:
                f += 2.0 * 9 * px * d
            f += 2.0 * 49 * px * d + 2 * 2.0 * px * d * d * v.mlp_ratio
        if i + 1 < len(v.dims):
            side //= 2
            px2 = side * side
            f += 2.0 * 49 * px2 * 2 * d + 2.0 * px2 * (2 * d) ** 2
    px = side * side
    ce = v.out_channels
    f += 2.0 * 9 * px * ce
    H = t.hidden
    f += 2.0 * px * (ce * H + H * H)
    w_layer = 2 * H * t.q_heads * t.head_dim + 2 * H * t.kv_heads * t.head_dim + 3 * H * t.intermediate
    f += t.layers * (2.0 * t_merged * w_layer + 2.0 * t_merged * t_merged * t.q_heads * t.head_dim)
    f += 2.0 * (state_dim * hidden_dim + (H + hidden_dim) * fusion_dim + fusion_dim * fusion_dim + fusion_dim * action_dim)
    return f
