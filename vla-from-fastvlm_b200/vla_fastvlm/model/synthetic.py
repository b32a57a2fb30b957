"""Seeded random-init weights for the `llava_qwen2` backbone and the FastVLA head.

No checkpoints can be downloaded where this is developed and benchmarked, so the benchmark and the
parity tests run on random weights of the real architecture.  The init is deliberately
non-degenerate (SURVEY.md trap T5): FastViT's default layer-scale (1e-5) and identity BatchNorm
would hide residual-branch bugs and HF's std-0.02 init makes attention uniform, so layer scales are
O(0.3), BatchNorm statistics/affines, every bias and every norm weight are randomised, and matrices
use 1/sqrt(fan_in) with residual-branch gains tuned so activations stay O(1) at every tap of the
FULL-DEPTH tower (2/12/24/4/2 blocks) — otherwise the late attention stages see one-hot softmaxes and the
whole forward becomes chaotically sensitive to rounding (measured; see scripts/parity_fullsize.py).

Keys are those of `LlavaQwen2ForCausalLM.state_dict()`; head keys are those of
`FastVLMWithExpert` (src/vla_fastvlm/fastvla/fastvlm_with_expert.py:23-38).
"""
from __future__ import annotations

import math
from typing import TYPE_CHECKING, Dict

import torch

if TYPE_CHECKING:  # keep this module loadable by file path (tests/golden/make_golden.py)
    from .arch import BackboneArch

VIS_PREFIX = "model.vision_tower.vision_tower.model."
PROJ_PREFIX = "model.mm_projector."
LLM_PREFIX = "model."


class _Init:
    def __init__(self, seed: int) -> None:
        self.g = torch.Generator(device="cpu").manual_seed(seed)

    def normal(self, *shape: int, std: float = 1.0, mean: float = 0.0) -> torch.Tensor:
        return torch.randn(*shape, generator=self.g, dtype=torch.float32) * std + mean

    def uniform(self, *shape: int, lo: float, hi: float) -> torch.Tensor:
        return torch.rand(*shape, generator=self.g, dtype=torch.float32) * (hi - lo) + lo

    def linear(self, n: int, k: int, gain: float = 1.0) -> torch.Tensor:
        return self.normal(n, k, std=gain / math.sqrt(k))

    def bn(self, sd: Dict[str, torch.Tensor], prefix: str, c: int) -> None:
        sd[prefix + ".weight"] = self.uniform(c, lo=0.6, hi=1.4)
        sd[prefix + ".bias"] = self.normal(c, std=0.1)
        sd[prefix + ".running_mean"] = self.normal(c, std=0.1)
        sd[prefix + ".running_var"] = self.uniform(c, lo=0.6, hi=1.4)


def synthetic_backbone_state_dict(arch: "BackboneArch", seed: int = 0,
                                  attn_norm: str = "batchnorm") -> Dict[str, torch.Tensor]:
    """`attn_norm`: the layer in front of MHSA — "batchnorm" (BatchNorm2d: weight, bias, running stats) or "layernorm"
    (LayerNormChannel: weight and bias only); the engine and the oracle tell them apart by the key set."""
    if attn_norm not in ("batchnorm", "layernorm"):
        raise ValueError("attn_norm must be 'batchnorm' or 'layernorm'")
    r = _Init(seed)
    sd: Dict[str, torch.Tensor] = {}
    v, t = arch.vision, arch.text
    P = VIS_PREFIX
    d0 = v.dims[0]
    # ---- stem (three reparameterised MobileOne blocks) ----
    sd[P + "patch_embed.0.reparam_conv.weight"] = r.normal(d0, 3, 3, 3, std=1.5 / math.sqrt(27))
    sd[P + "patch_embed.0.reparam_conv.bias"] = r.normal(d0, std=0.1)
    sd[P + "patch_embed.1.reparam_conv.weight"] = r.normal(d0, 1, 3, 3, std=1.0 / 3)
    sd[P + "patch_embed.1.reparam_conv.bias"] = r.normal(d0, std=0.1)
    sd[P + "patch_embed.2.reparam_conv.weight"] = r.linear(d0, d0, 1.5).view(d0, d0, 1, 1)
    sd[P + "patch_embed.2.reparam_conv.bias"] = r.normal(d0, std=0.1)

    def convffn(base: str, d: int) -> None:
        hd = d * v.mlp_ratio
        sd[base + ".convffn.conv.conv.weight"] = r.normal(d, 1, 7, 7, std=1.0 / 7)
        r.bn(sd, base + ".convffn.conv.bn", d)
        sd[base + ".convffn.fc1.weight"] = r.linear(hd, d, 1.2).view(hd, d, 1, 1)
        sd[base + ".convffn.fc1.bias"] = r.normal(hd, std=0.1)
        sd[base + ".convffn.fc2.weight"] = r.linear(d, hd, 1.2).view(d, hd, 1, 1)
        sd[base + ".convffn.fc2.bias"] = r.normal(d, std=0.1)

    idx = 0
    for i, d in enumerate(v.dims):
        if v.pos_emb[i]:
            # RepCPE reparameterised: depthwise 7x7 of pe(x) + x  -> identity tap + perturbation
            w = r.normal(d, 1, 7, 7, std=0.15 / 7)
            w[:, 0, 3, 3] += 1.0
            sd[P + f"network.{idx}.reparam_conv.weight"] = w
            sd[P + f"network.{idx}.reparam_conv.bias"] = r.normal(d, std=0.05)
            idx += 1
        for j in range(v.layers[i]):
            base = P + f"network.{idx}.{j}"
            if v.attention[i]:
                r.bn(sd, base + ".norm", d)
                if attn_norm == "layernorm":
                    del sd[base + ".norm.running_mean"], sd[base + ".norm.running_var"]
                    # LayerNormChannel emits unit-variance rows whatever the input scale; the q/k gains below assume
                    # O(0.2) inputs (as the BatchNorm variant delivers), so the gain lives in the norm weight.  With
                    # unit weights the logits have std ~50, softmax is an argmax and the tower is ill-conditioned:
                    # rounding the norm INPUT to bf16 once, inside the fp32 oracle, moves stage 4 by 12 %.
                    sd[base + ".norm.weight"] = sd[base + ".norm.weight"] * 0.2
                # q/k gains chosen so softmax logits have std ~2 (not one-hot, not uniform) on O(0.2) inputs
                sd[base + ".token_mixer.qkv.weight"] = torch.cat(
                    [r.linear(d, d, 7.0), r.linear(d, d, 7.0), r.linear(d, d, 3.0)], dim=0)
                sd[base + ".token_mixer.proj.weight"] = r.linear(d, d, 1.0)
                sd[base + ".token_mixer.proj.bias"] = r.normal(d, std=0.1)
                sd[base + ".layer_scale_1"] = r.uniform(d, 1, 1, lo=0.1, hi=0.3)
                sd[base + ".layer_scale_2"] = r.uniform(d, 1, 1, lo=0.1, hi=0.3)
            else:
                # RepMixer reparameterised: x + ls*(mixer(x) - norm(x)) as one depthwise 3x3
                w = r.normal(d, 1, 3, 3, std=0.1 / 3)
                w[:, 0, 1, 1] += 0.97
                sd[base + ".token_mixer.reparam_conv.weight"] = w
                sd[base + ".token_mixer.reparam_conv.bias"] = r.normal(d, std=0.05)
                sd[base + ".layer_scale"] = r.uniform(d, 1, 1, lo=0.1, hi=0.3)
            convffn(base, d)
        idx += 1
        if i + 1 < len(v.dims):
            d2 = v.dims[i + 1]
            base = P + f"network.{idx}"
            sd[base + ".proj.0.lkb_reparam.weight"] = r.normal(d2, 1, 7, 7, std=1.0 / 7)
            sd[base + ".proj.0.lkb_reparam.bias"] = r.normal(d2, std=0.1)
            sd[base + ".proj.1.reparam_conv.weight"] = r.linear(d2, d2, 1.3).view(d2, d2, 1, 1)
            sd[base + ".proj.1.reparam_conv.bias"] = r.normal(d2, std=0.1)
            idx += 1
    ce, cr = v.out_channels, v.se_reduced
    sd[P + "conv_exp.reparam_conv.weight"] = r.normal(ce, 1, 3, 3, std=2.0 / 3)
    sd[P + "conv_exp.reparam_conv.bias"] = r.normal(ce, std=0.1)
    sd[P + "conv_exp.se.reduce.weight"] = r.linear(cr, ce, 1.0).view(cr, ce, 1, 1)
    sd[P + "conv_exp.se.reduce.bias"] = r.normal(cr, std=0.1)
    sd[P + "conv_exp.se.expand.weight"] = r.linear(ce, cr, 1.0).view(ce, cr, 1, 1)
    sd[P + "conv_exp.se.expand.bias"] = r.normal(ce, std=0.5)
    # ---- mlp2x_gelu projector ----
    H = t.hidden
    sd[PROJ_PREFIX + "0.weight"] = r.linear(H, ce, 2.0)
    sd[PROJ_PREFIX + "0.bias"] = r.normal(H, std=0.1)
    sd[PROJ_PREFIX + "2.weight"] = r.linear(H, H, 2.0)
    sd[PROJ_PREFIX + "2.bias"] = r.normal(H, std=0.1)
    # ---- Qwen2 ----
    L = LLM_PREFIX
    sd[L + "embed_tokens.weight"] = r.normal(t.vocab, H, std=0.5)
    for l in range(t.layers):
        b = L + f"layers.{l}"
        sd[b + ".input_layernorm.weight"] = r.uniform(H, lo=0.7, hi=1.3)
        sd[b + ".self_attn.q_proj.weight"] = r.linear(t.q_heads * t.head_dim, H, 1.0)
        sd[b + ".self_attn.q_proj.bias"] = r.normal(t.q_heads * t.head_dim, std=0.2)
        sd[b + ".self_attn.k_proj.weight"] = r.linear(t.kv_heads * t.head_dim, H, 1.0)
        sd[b + ".self_attn.k_proj.bias"] = r.normal(t.kv_heads * t.head_dim, std=0.2)
        sd[b + ".self_attn.v_proj.weight"] = r.linear(t.kv_heads * t.head_dim, H, 1.0)
        sd[b + ".self_attn.v_proj.bias"] = r.normal(t.kv_heads * t.head_dim, std=0.1)
        sd[b + ".self_attn.o_proj.weight"] = r.linear(H, t.q_heads * t.head_dim, 0.5)
        sd[b + ".post_attention_layernorm.weight"] = r.uniform(H, lo=0.7, hi=1.3)
        sd[b + ".mlp.gate_proj.weight"] = r.linear(t.intermediate, H, 1.0)
        sd[b + ".mlp.up_proj.weight"] = r.linear(t.intermediate, H, 1.0)
        sd[b + ".mlp.down_proj.weight"] = r.linear(H, t.intermediate, 0.7)
    sd[L + "norm.weight"] = r.uniform(H, lo=0.7, hi=1.3)
    return sd


def synthetic_head_state_dict(hidden: int, state_dim: int, action_dim: int, hidden_dim: int,
                              fusion_dim: int, seed: int = 1) -> Dict[str, torch.Tensor]:
    r = _Init(seed)
    sd: Dict[str, torch.Tensor] = {}
    sd["state_projection.0.weight"] = r.uniform(state_dim, lo=0.7, hi=1.3)
    sd["state_projection.0.bias"] = r.normal(state_dim, std=0.1)
    sd["state_projection.1.weight"] = r.linear(hidden_dim, state_dim, 1.0)
    sd["state_projection.1.bias"] = r.normal(hidden_dim, std=0.1)
    sd["fusion.0.weight"] = r.linear(fusion_dim, hidden + hidden_dim, 1.0)
    sd["fusion.0.bias"] = r.normal(fusion_dim, std=0.1)
    sd["fusion.1.weight"] = r.uniform(fusion_dim, lo=0.7, hi=1.3)
    sd["fusion.1.bias"] = r.normal(fusion_dim, std=0.1)
    sd["fusion.4.weight"] = r.linear(fusion_dim, fusion_dim, 1.0)
    sd["fusion.4.bias"] = r.normal(fusion_dim, std=0.1)
    sd["action_head.weight"] = r.linear(action_dim, fusion_dim, 1.0)
    sd["action_head.bias"] = r.normal(action_dim, std=0.1)
    return sd
