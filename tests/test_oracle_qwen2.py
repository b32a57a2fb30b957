"""Pins the oracle's Qwen2 restatement to the `transformers` Qwen2Model shipped in the image
(transformers/models/qwen2/modeling_qwen2.py): same weights, inputs_embeds path, right-padded batch."""
import pytest
import torch

from helpers import tiny_weights


def test_qwen2_restatement_matches_transformers():
    transformers = pytest.importorskip("transformers")
    from transformers import Qwen2Config, Qwen2Model

    from oracle.fastvla_oracle import LLM, qwen2_forward

    arch, sd, _ = tiny_weights(0)
    t = arch.text
    cfg = Qwen2Config(vocab_size=t.vocab, hidden_size=t.hidden, intermediate_size=t.intermediate,
                      num_hidden_layers=t.layers, num_attention_heads=t.q_heads, num_key_value_heads=t.kv_heads,
                      rms_norm_eps=t.rms_eps, rope_theta=t.rope_theta, max_position_embeddings=t.max_position,
                      tie_word_embeddings=True, attn_implementation="eager")
    if hasattr(cfg, "rope_parameters") and isinstance(cfg.rope_parameters, dict):
        cfg.rope_parameters["rope_theta"] = t.rope_theta
    model = Qwen2Model(cfg).eval()
    hf_sd = {k[len(LLM):]: v for k, v in sd.items()
             if k.startswith(LLM) and not k.startswith(LLM + "vision_tower") and not k.startswith(LLM + "mm_projector")}
    missing, unexpected = model.load_state_dict(hf_sd, strict=False)
    assert not unexpected and not [m for m in missing if "rotary" not in m], (missing, unexpected)

    g = torch.Generator().manual_seed(0)
    B, T = 3, 21
    x = torch.randn(B, T, t.hidden, generator=g)
    mask = torch.ones(B, T, dtype=torch.bool)
    mask[1, 15:] = False
    mask[2, 6:] = False
    with torch.no_grad():
        ref = model(inputs_embeds=x, attention_mask=mask.long(), output_hidden_states=True)
        got = qwen2_forward(sd, x, mask, t.layers, t.q_heads, t.kv_heads, t.head_dim, t.rms_eps, t.rope_theta)
    want = ref.hidden_states[-1]
    assert torch.equal(want, ref.last_hidden_state)  # hidden_states[-1] is post-norm (adapter :551-556)
    m = mask[..., None]
    err = ((got - want) * m).abs().max().item()
    assert err < 2e-5, err
