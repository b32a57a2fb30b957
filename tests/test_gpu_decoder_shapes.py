"""Engine-level parity for the Qwen2-1.5B and Qwen2-7B decoder geometries (BASELINE configs 3/4): hidden 1536 / 3584,
head_dim 128, 12q/2kv and 28q/4kv heads, intermediate 8960 / 18944 — at depth 2 behind a small tower whose 1024^2 input
gives the real 256 image tokens (T' = 256 + T_text), so the CPU oracle stays cheap.  Every decoder layer is tapped.
fp32: actions <= 1e-3 max-abs; bf16: <= 2e-2 relative (north_star)."""
import pytest
import torch

from helpers import make_inputs, rel_err

pytestmark = pytest.mark.gpu
HEAD = dict(state_dim=14, action_dim=14, hidden_dim=256, fusion_dim=256)


@pytest.mark.parametrize("preset", ["dec-1.5b-2l", "dec-7b-2l"])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_wide_decoder_matches_oracle(preset, dtype):
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    from oracle.fastvla_oracle import FastVLAOracle
    from vla_fastvlm import _native as N
    from vla_fastvlm.model.arch import PRESETS
    from vla_fastvlm.model.engine import BACKBONE_KEY_PREFIX, NativeEngine
    from vla_fastvlm.model.synthetic import synthetic_backbone_state_dict, synthetic_head_state_dict

    arch = PRESETS[preset]
    t = arch.text
    sd = synthetic_backbone_state_dict(arch, 0)
    hsd = synthetic_head_state_dict(t.hidden, HEAD["state_dim"], HEAD["action_dim"], HEAD["hidden_dim"],
                                    HEAD["fusion_dim"], 1)
    B, T = 4, 17   # M = 4 * 272 = 1088 rows: the fused-norm ("emit") decoder path; B = 1 runs the split-K path below
    images, states, ids, mask = make_inputs(B, 240, 320, T, t.vocab, HEAD["state_dim"], seed=21, image_mode="prefix")
    taps = {}
    ref = FastVLAOracle(arch, sd, hsd).forward(images, states, ids, mask, taps=taps)
    Tm = taps["embeds"].shape[1]
    assert Tm == arch.vision.num_tokens + T - 1 == 272

    eng = NativeEngine(arch, dtype=dtype, **HEAD)
    eng.load_state_dict(sd, prefix=BACKBONE_KEY_PREFIX)
    eng.load_state_dict(hsd)
    eng.finalize()
    dev = eng.device
    bufs = {N.TAP_PROJECTOR: torch.zeros(B, arch.vision.num_tokens, t.hidden, device=dev, dtype=dtype),
            N.TAP_EMBEDS: torch.zeros(B, Tm, t.hidden, device=dev, dtype=dtype),
            N.TAP_POOLED: torch.zeros(B, t.hidden, device=dev, dtype=torch.float32)}
    for l in range(t.layers):
        bufs[N.TAP_LAYER0 + l] = torch.zeros(B, Tm, t.hidden, device=dev, dtype=dtype)
    for k, b in bufs.items():
        eng.set_tap(k, b)
    out = eng.forward(images.to(dev), ids, mask.sum(1), states=states.to(dev)).float().cpu()
    assert eng.merged_len == Tm
    valid = torch.zeros(B, Tm, 1)
    for b in range(B):
        valid[b, : int(mask[b].sum()) + arch.vision.num_tokens - 1] = 1
    errs = {"projector": rel_err(bufs[N.TAP_PROJECTOR], taps["projector"]),
            "embeds": rel_err(bufs[N.TAP_EMBEDS], taps["embeds"])}
    for l in range(t.layers):
        errs[f"layer{l}"] = rel_err(bufs[N.TAP_LAYER0 + l].float().cpu() * valid, taps[f"layer{l}"] * valid)
    errs["pooled"] = rel_err(bufs[N.TAP_POOLED], taps["pooled"])
    err = float((out - ref).abs().max())
    rel = err / float(ref.abs().max())
    print(f"{preset} {dtype}: actions max-abs {err:.3e} rel {rel:.3e}; stages {errs}")
    stage_tol = 2e-4 if dtype == torch.float32 else 4e-2
    assert all(e <= stage_tol for e in errs.values()), errs
    if dtype == torch.float32:
        assert err <= 1e-3, (err, errs)
    else:
        assert rel <= 2e-2, (rel, errs)
    # a single observation (M = T' rows): split-K stream updates + weight-less norm kernel; same actions as in the batch
    one = eng.forward(images[:1].to(dev), ids[:1], mask[:1].sum(1), states=states[:1].to(dev)).float().cpu()
    if dtype == torch.float32:
        assert (one[0] - ref[0]).abs().max() <= 1e-3
    else:
        assert float((one[0] - ref[0]).abs().max() / ref.abs().max()) <= 2e-2
