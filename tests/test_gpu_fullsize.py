"""North-star parity at FULL size: FastVLA on FastVLM-0.5B shapes (FastViTHD 2/12/24/4/2 @1024^2, Qwen2-0.5B,
T' = 256 + 16), seeded random init, one synthetic frame + prompt + 14-dim state — BASELINE.json configs[0] —
engine (through the C ABI) vs the CPU fp32 oracle.  fp32: actions within 1e-3 max-abs; bf16: within 2e-2
relative; per-stage tensors checked.  ~40 s (the oracle needs ~2 s/sample on the host cores)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def fullsize():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    from oracle.fastvla_oracle import IMAGE_TOKEN_INDEX, FastVLAOracle
    from vla_fastvlm.model.arch import PRESETS
    from vla_fastvlm.model.synthetic import synthetic_backbone_state_dict, synthetic_head_state_dict

    arch = PRESETS["fastvlm-0.5b"]
    sd = synthetic_backbone_state_dict(arch, 0)
    hsd = synthetic_head_state_dict(arch.text.hidden, 14, 14, 1024, 1024, 1)
    g = torch.Generator().manual_seed(1)
    images = torch.rand(1, 3, 1024, 1024, generator=g)          # 1024^2: resize is the identity, no padding
    states = torch.randn(1, 14, generator=g)
    ids = torch.randint(0, 151643, (1, 17), generator=torch.Generator().manual_seed(2))
    ids[:, 0] = IMAGE_TOKEN_INDEX                                # prefix mode: T' = 256 + 16 = 272
    mask = torch.ones(1, 17, dtype=torch.long)
    taps = {}
    ref = FastVLAOracle(arch, sd, hsd).forward(images, states, ids, mask, taps=taps)
    assert taps["embeds"].shape[1] == 272
    return dict(arch=arch, sd=sd, hsd=hsd, images=images, states=states, ids=ids, mask=mask, ref=ref, taps=taps)


@pytest.mark.parametrize("dtype,stage_tol", [(torch.float32, 1e-4), (torch.bfloat16, 4e-2)])
def test_fastvla_05b_matches_oracle(fullsize, dtype, stage_tol):
    from vla_fastvlm import _native as N
    from vla_fastvlm.model.engine import BACKBONE_KEY_PREFIX, NativeEngine

    f = fullsize
    arch, v = f["arch"], f["arch"].vision
    eng = NativeEngine(arch, dtype=dtype, state_dim=14, action_dim=14)
    eng.load_state_dict(f["sd"], prefix=BACKBONE_KEY_PREFIX)
    eng.load_state_dict(f["hsd"])
    eng.finalize()
    dev = eng.device
    bufs = {}
    side = v.image_size // 4
    for i, d in enumerate(v.dims):
        bufs[N.TAP_VIS_STAGE0 + i] = torch.zeros(1, side, side, d, device=dev, dtype=dtype)
        side //= 2
    bufs[N.TAP_IMAGE_FEATURES] = torch.zeros(1, v.num_tokens, v.out_channels, device=dev, dtype=dtype)
    bufs[N.TAP_LAYER0 + arch.text.layers - 1] = torch.zeros(1, 272, arch.text.hidden, device=dev, dtype=dtype)
    bufs[N.TAP_POOLED] = torch.zeros(1, arch.text.hidden, device=dev, dtype=torch.float32)
    for k, b in bufs.items():
        eng.set_tap(k, b)
    out = eng.forward(f["images"].to(dev), f["ids"], f["mask"].sum(1), states=f["states"].to(dev)).float().cpu()
    assert eng.merged_len == 272

    def rel(x, y):
        return float((x.float().cpu() - y).abs().max() / y.abs().max())

    errs = {f"vis_stage{i}": rel(bufs[N.TAP_VIS_STAGE0 + i], f["taps"][f"vis_stage{i}"].permute(0, 2, 3, 1))
            for i in range(len(v.dims))}
    errs["image_features"] = rel(bufs[N.TAP_IMAGE_FEATURES], f["taps"]["image_features"])
    errs["last_layer"] = rel(bufs[N.TAP_LAYER0 + arch.text.layers - 1], f["taps"][f"layer{arch.text.layers - 1}"])
    errs["pooled"] = rel(bufs[N.TAP_POOLED], f["taps"]["pooled"])
    err = (out - f["ref"]).abs().max().item()
    print(f"fullsize {dtype}: action max-abs {err:.3e} rel {err / f['ref'].abs().max().item():.3e} stages {errs}")
    assert all(e <= stage_tol for e in errs.values()), errs
    if dtype == torch.float32:
        assert err <= 1e-3, (err, errs)
    else:
        assert err / f["ref"].abs().max().item() <= 2e-2, (err, errs)
