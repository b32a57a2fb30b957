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


def write_checkpoint_dir(path, arch: BackboneArch, sd: Dict[str, torch.Tensor], describe_tower: bool = True,
                         dtype: torch.dtype = torch.float32, with_lm_head: bool = True):
    """A local HF-style `llava_qwen2` checkpoint directory the way Apple's exports look: config.json with the Qwen2
    fields (+ `mm_vision_tower`; the tower GEOMETRY only when it is not FastViTHD's), model.safetensors under the HF
    tensor names (a tied `lm_head.weight` included), and a word-level `tokenizer.json` so that AutoTokenizer works
    offline.  `describe_tower=False` leaves the tower undescribed, like a local export without `auto_map`."""
    import json
    from dataclasses import asdict
    from pathlib import Path

    from safetensors.torch import save_file
    from tokenizers import Tokenizer, models, pre_tokenizers

    path = Path(path)
    path.mkdir(parents=True, exist_ok=True)
    t = arch.text
    cfg = dict(model_type="llava_qwen2", architectures=["LlavaQwen2ForCausalLM"], hidden_size=t.hidden,
               num_hidden_layers=t.layers, num_attention_heads=t.q_heads, num_key_value_heads=t.kv_heads,
               intermediate_size=t.intermediate, vocab_size=t.vocab, rms_norm_eps=t.rms_eps, rope_theta=t.rope_theta,
               max_position_embeddings=t.max_position, tokenizer_padding_side="right", tie_word_embeddings=True)
    if t.head_dim != t.hidden // t.q_heads:
        cfg["head_dim"] = t.head_dim
    if describe_tower:
        cfg["mm_vision_tower"] = arch.mm_vision_tower
        v = asdict(arch.vision)
        v.pop("image_size")
        cfg["vision_arch"] = v
    (path / "config.json").write_text(json.dumps(cfg, indent=1))
    tensors = {k: v.to(dtype).contiguous() for k, v in sd.items()}
    if with_lm_head:
        tensors["lm_head.weight"] = tensors["model.embed_tokens.weight"].clone()
    save_file(tensors, str(path / "model.safetensors"))
    words = ("pick up the red block open drawer push insert peg into hole stack cubes close door button "
             "press slide turn place").split()
    vocab = {"[PAD]": 0, "[UNK]": 1, "\n": 2}
    for w in words:
        vocab.setdefault(w, len(vocab))
    tok = Tokenizer(models.WordLevel(vocab, unk_token="[UNK]"))
    tok.pre_tokenizer = pre_tokenizers.Sequence([pre_tokenizers.Split(" ", "removed"),
                                                 pre_tokenizers.Split("\n", "isolated")])
    tok.save(str(path / "tokenizer.json"))
    (path / "tokenizer_config.json").write_text(json.dumps(
        {"tokenizer_class": "PreTrainedTokenizerFast", "pad_token": "[PAD]", "unk_token": "[UNK]",
         "model_max_length": 2048}))
    return path
