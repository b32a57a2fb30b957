"""CPU tests of the host-side logic: prompt/time-step normalisation, action queue, tokenizer protocol,
architecture presets, FLOP model, batch sharding."""
import sys
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parents[1]


def test_prompt_normalisation_rules():
    from vla_fastvlm.shared import as_prompt_list

    assert as_prompt_list(None, 2, True) == ["\n", "\n"]                      # None -> "" (+ newline)
    assert as_prompt_list("go", 3, True) == ["go\n"] * 3                     # scalar broadcast
    assert as_prompt_list(["a"], 2, False) == ["a", "a"]                      # single-element list broadcast
    assert as_prompt_list(["a\n", "b"], 2, True) == ["a\n", "b\n"]           # no double newline
    assert as_prompt_list(7, 1, False) == ["7"]                              # anything else is str()-ed


def test_pick_step_and_queue():
    from vla_fastvlm.shared import ActionQueue, pick_step

    x = torch.arange(2 * 3 * 4).reshape(2, 3, 4)
    assert torch.equal(pick_step(x, 2, -1), x[:, -1]) and torch.equal(pick_step(x, 2, 0), x[:, 0])
    assert pick_step(x[:, 0], 2, -1).shape == (2, 4)
    q = ActionQueue(2)
    chunk = torch.arange(2 * 3 * 5).reshape(2, 3, 5).float()   # (B, n, D)
    q.refill(chunk)
    assert len(q) == 2
    assert torch.equal(q.pop(), chunk[:, 0]) and torch.equal(q.pop(), chunk[:, 1]) and len(q) == 0


def test_byte_tokenizer_protocol():
    from vla_fastvlm.model.tokenizer import SimpleByteTokenizer

    tok = SimpleByteTokenizer(512)
    out = tok(["ab\n", "abcdef\n"], padding="longest", truncation=True, max_length=5, return_tensors="pt")
    assert out["input_ids"].shape == (2, 5) and out["attention_mask"].sum(1).tolist() == [3, 5]
    assert (out["input_ids"][0, 3:] == 0).all()                               # right padded
    assert (out["input_ids"] < 512).all() and (out["input_ids"][out["attention_mask"].bool()] >= 3).all()
    tok.padding_side = "left"
    left = tok(["ab\n", "abcdef\n"], padding="max_length", max_length=8)
    assert left["input_ids"].shape == (2, 8) and left["attention_mask"][0].tolist() == [0] * 5 + [1] * 3


def test_arch_presets_and_flop_model():
    from oracle.fastvla_oracle import flops_per_sample
    from vla_fastvlm.model.arch import PRESETS, arch_from_hf_config

    a = PRESETS["fastvlm-0.5b"]
    assert a.vision.num_tokens == 256 and a.vision.out_channels == 3072 and a.vision.se_reduced == 192
    assert abs(flops_per_sample(a, 272, 14, 14) / 1e9 - 686.0) < 2.0       # ~686 GFLOP / sample
    hf = dict(model_type="llava_qwen2", hidden_size=1536, num_attention_heads=12, num_key_value_heads=2,
              num_hidden_layers=28, intermediate_size=8960, vocab_size=151936, mm_vision_tower="mobileclip_l_1024")
    b = arch_from_hf_config(hf)
    assert b.text == PRESETS["fastvlm-1.5b"].text and b.vision.image_size == 1024


def test_oracle_attention_norm_variants():
    """The oracle's two restatements of AttentionBlock.norm: LayerNormChannel equals torch's layer_norm over the channel
    axis; the key set selects the variant (SURVEY App. A leaves the layer type unverified)."""
    from oracle.fastvla_oracle import _attn_norm, _bn, _ln_channel

    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, 24, 5, 7, generator=g) * 2 + 1
    sd = {"n.weight": torch.rand(24, generator=g) + 0.5, "n.bias": torch.randn(24, generator=g)}
    want = torch.nn.functional.layer_norm(x.permute(0, 2, 3, 1), (24,), sd["n.weight"], sd["n.bias"], 1e-5)
    assert torch.allclose(_ln_channel(sd, "n", x), want.permute(0, 3, 1, 2), atol=1e-5)
    assert torch.equal(_attn_norm(sd, "n", x), _ln_channel(sd, "n", x))
    sd.update({"n.running_mean": torch.randn(24, generator=g), "n.running_var": torch.rand(24, generator=g) + 0.5})
    assert torch.equal(_attn_norm(sd, "n", x), _bn(sd, "n", x))


def test_shard_range_partitions_the_batch():
    sys.path.insert(0, str(ROOT))
    import bench

    for total, world in [(64, 8), (64, 3), (5, 8), (1, 1)]:
        spans = [bench.shard_range(total, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == total
        assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
        assert max(e - s for s, e in spans) - min(e - s for s, e in spans) <= 1
    pool = bench.make_prompt_pool()
    assert len(pool) == 50 and len(set(pool)) == 50 and all(8 <= len(p) + 1 <= 16 for p in pool)
