"""Product path (public policy API -> C ABI -> sm_100a kernels) against the golden fixtures recorded
from the reference's own classes (tests/golden/make_golden.py).  fp32: actions within 1e-3 max-abs;
bf16: within 2e-2 relative (north_star)."""
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent / "golden"))
from cases import CASES, TINY_HEAD, case_inputs  # noqa: E402

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).resolve().parent / "golden"


def _policy(dtype, spec):
    from vla_fastvlm.fastvla import FastVLAConfig, FastVLAPolicy
    from helpers import tiny_weights

    cfg = FastVLAConfig(vlm_model_name="synthetic:tiny", compute_dtype=dtype,
                        image_token_mode="prefix" if spec["prefix"] else "none", **TINY_HEAD)
    pol = FastVLAPolicy(cfg)
    pol.model.backbone.config.image_feature_pool = spec["pool"]
    pol.model.backbone.model.configure_engine(pool_mode=spec["pool"])
    _, _, hsd = tiny_weights(0)
    missing = pol.model.load_state_dict(hsd, strict=False)
    assert not missing.unexpected_keys
    return pol.cuda().eval()


def _check_ill_conditioned_bf16(pol, images, states, tasks, actions, want):
    """Raw 0..255 pixels drive the random-init tower's attention stages into near one-hot softmaxes: rounding only the
    WEIGHTS to bf16 moves the fp32 oracle's actions by O(10 %) (asserted below, so the exemption cannot outlive its
    reason).  The end-to-end figure therefore says nothing about the bf16 arithmetic; the case is pinned end to end in
    fp32, and in bf16 every well-conditioned segment is held to the north_star tolerance separately:
      (1) ingest + stem + the three RepMixer stages (before the first attention block) against the oracle taps,
      (2) everything after the tower — projector, splice, Qwen2, pooling, head — against the oracle continued from the
          ENGINE's own image features (oracle on bf16-rounded weights: what a bf16 run holds)."""
    from helpers import tiny_weights
    from oracle.fastvla_oracle import FastVLAOracle
    from vla_fastvlm import _native as N

    arch, sd, hsd = tiny_weights(0)
    sd_q = {k: (v.bfloat16().float() if v.is_floating_point() and v.ndim >= 2 else v) for k, v in sd.items()}
    be = pol.model.backbone
    prompts = pol.processor.prepare_tasks(tasks, actions.shape[0])
    ids, lens, _ = be._prompt_ids(prompts)
    mask = (torch.arange(ids.shape[1])[None, :] < lens[:, None]).long()
    taps, taps_q = {}, {}
    ref = FastVLAOracle(arch, sd, hsd).forward(images, states, ids.long(), mask, taps=taps)
    ref_q = FastVLAOracle(arch, sd_q, hsd).forward(images, states, ids.long(), mask, taps=taps_q)
    assert np.abs(ref.numpy() - want).max() <= 1e-4                      # the oracle reproduces the reference golden
    sensitivity = float((ref_q - ref).abs().max() / ref.abs().max())
    assert sensitivity > 2e-2, f"case is no longer ill-conditioned ({sensitivity:.2e}): test it end to end instead"

    eng = be.model.engine
    dev, v, Bn = eng.device, arch.vision, images.shape[0]
    bufs, side = {}, v.image_size // 4
    bufs[N.TAP_STEM] = torch.zeros(Bn, side, side, v.dims[0], device=dev, dtype=torch.bfloat16)
    for i, d in enumerate(v.dims):
        bufs[N.TAP_VIS_STAGE0 + i] = torch.zeros(Bn, side, side, d, device=dev, dtype=torch.bfloat16)
        side //= 2
    bufs[N.TAP_IMAGE_FEATURES] = torch.zeros(Bn, v.num_tokens, v.out_channels, device=dev, dtype=torch.bfloat16)
    for k, b in bufs.items():
        eng.set_tap(k, b)
    try:
        with torch.no_grad():
            got = pol.forward(images.to(dev), states.to(dev), tasks, device=dev).float().cpu()
    finally:
        for k in bufs:
            eng.set_tap(k, None)
    assert np.array_equal(got.numpy(), actions)                          # taps do not change the result

    def rel(x, y):
        return float((x.float().cpu() - y).abs().max() / y.abs().max())

    seg1 = {"stem": rel(bufs[N.TAP_STEM], taps_q["stem"].permute(0, 2, 3, 1))}
    for i, attn in enumerate(v.attention):
        if attn:
            break
        seg1[f"vis_stage{i}"] = rel(bufs[N.TAP_VIS_STAGE0 + i], taps_q[f"vis_stage{i}"].permute(0, 2, 3, 1))
    cont = FastVLAOracle(arch, sd_q, hsd).forward(images, states, ids.long(), mask,
                                                  inject={"image_features": bufs[N.TAP_IMAGE_FEATURES]})
    seg2 = float((got - cont).abs().max() / cont.abs().max())
    print(f"uint8-valued bf16: oracle sensitivity to weight rounding {sensitivity:.2e}; pre-attention stages {seg1}; "
          f"post-tower actions rel {seg2:.2e}")
    assert all(e <= 6e-2 for e in seg1.values()), seg1                   # the tiny-arch bf16 stage tolerance
    assert seg2 <= 2e-2, seg2


@pytest.mark.parametrize("name", sorted(CASES))
@pytest.mark.parametrize("dtype", ["float32", "bfloat16"])
def test_policy_forward_matches_reference_golden(name, dtype):
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    spec = CASES[name]
    gold = dict(np.load(GOLD / f"tiny_{name}.npz"))
    images, states, tasks = case_inputs(name)
    pol = _policy(dtype, spec)
    dev = torch.device("cuda")
    with torch.no_grad():
        actions = pol.forward(images.to(dev), states.to(dev), tasks, device=dev).float().cpu().numpy()
        pixel = pol.processor.prepare_images(images.to(dev), dev).cpu()
    assert pixel.shape[1:] == (3, 256, 256)
    assert np.allclose(pixel[:, :, ::8, ::8].numpy(), gold["pixel_probe"], rtol=0, atol=2e-4 * max(1.0, float(np.abs(gold["pixel_probe"]).max())))
    ids, lens, _ = pol.model.backbone._prompt_ids(pol.processor.prepare_tasks(tasks, actions.shape[0]))
    assert np.array_equal(ids.numpy() * (np.arange(ids.shape[1])[None] < lens.numpy()[:, None]),
                          gold["input_ids"] * gold["attention_mask"])
    want = gold["actions"]
    if dtype == "float32":
        assert np.abs(actions - want).max() <= 1e-3, np.abs(actions - want).max()
    elif name == "prefix_bhwc_uint8_values":
        _check_ill_conditioned_bf16(pol, images, states, tasks, actions, want)
    else:
        assert np.abs(actions - want).max() / np.abs(want).max() <= 2e-2, (actions, want)


def test_select_action_single_observation():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    spec = CASES["prefix_letterbox"]
    gold = dict(np.load(GOLD / "tiny_prefix_letterbox.npz"))
    images, states, tasks = case_inputs("prefix_letterbox")
    pol = _policy("float32", spec)
    a0 = pol.select_action(images[0], states[0], tasks[0], torch.device("cuda")).cpu().numpy()
    # a single right-padded sample sees exactly the same tokens as in the batch -> same action
    assert np.abs(a0 - gold["actions"][0]).max() <= 1e-3


def test_head_training_path_and_refresh():
    """Training mode: engine backbone (no grad) + autograd head; after an optimizer step the eval path
    picks the new head weights up."""
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    spec = CASES["none_ragged"]
    images, states, tasks = case_inputs("none_ragged")
    pol = _policy("float32", spec)
    dev = torch.device("cuda")
    with torch.no_grad():
        before = pol.forward(images.to(dev), states.to(dev), tasks, device=dev).clone()
    pol.train()
    opt = torch.optim.SGD([p for p in pol.parameters() if p.requires_grad], lr=0.05)
    torch.manual_seed(0)
    out = pol.compute_loss({"images": images.to(dev), "states": states.to(dev), "tasks": tasks,
                            "actions": torch.zeros(3, TINY_HEAD["action_dim"], device=dev)})
    out["loss"].backward()
    assert all(p.grad is None for p in pol.model.backbone.parameters())
    assert pol.model.action_head.weight.grad is not None
    opt.step()
    pol.eval()
    with torch.no_grad():
        after = pol.forward(images.to(dev), states.to(dev), tasks, device=dev)
        # eager evaluation of the updated head on the engine's pooled features, for comparison
        feats = pol.model.backbone(images.to(dev), pol.processor.prepare_tasks(tasks, 3), device=dev)
        s = pol.model.state_projection(states.to(dev))
        want = pol.model.action_head(pol.model.fusion(torch.cat([feats, s], -1)))
    assert (after - before).abs().max() > 1e-4
    assert torch.allclose(after, want, atol=1e-4, rtol=1e-4)
