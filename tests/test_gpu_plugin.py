"""LeRobot plugin surface (`vla_fastvlm.lerobot_fastvla`) exercised with a minimal fake `lerobot` package:
registration, feature resolution + error texts, select_action queue semantics, forward loss dict."""
import sys
from pathlib import Path

import pytest
import torch

FAKE = Path(__file__).resolve().parent / "fake_lerobot"
if str(FAKE) not in sys.path:
    sys.path.insert(0, str(FAKE))
sys.path.insert(0, str(Path(__file__).resolve().parent / "golden"))

from cases import TINY_HEAD  # noqa: E402


def _features(n_cam=2, state=6, action=5):
    from lerobot.configs.types import FeatureType, PolicyFeature

    inp = {f"observation.images.cam{i}": PolicyFeature(FeatureType.VISUAL, (3, 96, 128)) for i in range(n_cam)}
    inp["observation.state"] = PolicyFeature(FeatureType.STATE, (state,))
    return inp, {"action": PolicyFeature(FeatureType.ACTION, (action,))}


def test_config_registration_and_validation():
    from lerobot.configs.policies import PreTrainedConfig
    from lerobot.configs.types import FeatureType, PolicyFeature

    from vla_fastvlm.lerobot_fastvla import FastVLAConfig, FastVLAPolicy, make_fastvla_pre_post_processors  # noqa: F401

    assert PreTrainedConfig.get_choice_class("fastvla") is FastVLAConfig
    assert FastVLAPolicy.name == "fastvla" and FastVLAPolicy.config_class is FastVLAConfig
    cfg = FastVLAConfig()
    assert (cfg.chunk_size, cfg.n_action_steps, cfg.n_obs_steps) == (1, 1, 1)
    assert cfg.observation_delta_indices == [0] and cfg.action_delta_indices == [0] and cfg.reward_delta_indices is None
    assert FastVLAConfig(chunk_size=4).action_delta_indices == [0, 1, 2, 3]
    with pytest.raises(ValueError, match="n_action_steps must be <= chunk_size"):
        FastVLAConfig(n_action_steps=3, chunk_size=2)
    only_state = FastVLAConfig(input_features={"observation.state": PolicyFeature(FeatureType.STATE, (4,))})
    with pytest.raises(ValueError, match="at least one visual"):
        only_state.validate_features()
    inp, _ = _features()
    inp.pop("observation.state")
    with pytest.raises(ValueError, match="at least one state"):
        FastVLAConfig(input_features=inp).validate_features()
    opt, sch = cfg.get_optimizer_preset(), cfg.get_scheduler_preset()
    assert (opt.lr, opt.betas, opt.weight_decay, opt.grad_clip_norm) == (1e-4, (0.9, 0.95), 1e-4, 1.0)
    assert (sch.num_warmup_steps, sch.num_decay_steps, sch.decay_lr) == (500, 20_000, 2.5e-6)


@pytest.mark.gpu
def test_plugin_select_action_and_forward():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    from vla_fastvlm.lerobot_fastvla import FastVLAConfig, FastVLAPolicy, make_fastvla_pre_post_processors

    inp, outp = _features()
    cfg = FastVLAConfig(input_features=inp, output_features=outp, device="cuda", vlm_model_name="synthetic:tiny",
                        hidden_dim=TINY_HEAD["hidden_dim"], fusion_dim=TINY_HEAD["fusion_dim"], image_token_mode="prefix",
                        chunk_size=4, n_action_steps=1)
    with pytest.raises(ValueError, match="input_features to be set"):
        FastVLAPolicy(FastVLAConfig(vlm_model_name="synthetic:tiny"))
    pol = FastVLAPolicy(cfg).cuda()
    assert (pol.config.state_dim, pol.config.action_dim) == (6, 5)          # inferred from the features
    assert pol._image_keys[0] == "observation.images.cam0" and pol._state_key == "observation.state"
    keys = set(pol.state_dict().keys())
    assert {"model.state_projection.0.weight", "model.fusion.4.bias", "model.action_head.weight"} <= keys

    g = torch.Generator().manual_seed(0)
    B = 3
    batch = {"observation.images.cam0": torch.rand(B, 3, 96, 128, generator=g).cuda(),
             "observation.images.cam1": torch.rand(B, 3, 96, 128, generator=g).cuda(),
             "observation.state": torch.randn(B, 6, generator=g).cuda(), "task": ["open the drawer"]}
    a1 = pol.select_action(batch)
    assert a1.shape == (B, 5) and len(pol._action_queue) == 0 and not pol.training
    chunk = pol.predict_action_chunk(batch)
    assert chunk.shape == (B, 1, 5) and torch.allclose(chunk[:, 0], a1)
    other = dict(batch)
    other["observation.images.cam1"] = torch.zeros_like(batch["observation.images.cam1"])
    assert torch.equal(pol.select_action(other), a1)                         # only the first camera is used (F7)
    time_major = dict(batch)
    time_major["observation.images.cam0"] = torch.stack([torch.zeros_like(batch["observation.images.cam0"]),
                                                        batch["observation.images.cam0"]], 1)
    assert torch.allclose(pol.select_action(time_major), a1)                 # last observation step
    no_task = dict(batch)
    no_task.pop("task")
    assert pol.select_action(no_task).shape == (B, 5)                        # task None -> ""

    pol.train()
    batch["action"] = torch.randn(B, 4, 5, generator=g).cuda()               # chunk of 4 loaded; loss uses step 0
    loss, info = pol.forward(batch)
    assert loss.requires_grad and set(info) == {"loss", "mse"} and info["loss"] == pytest.approx(loss.item())
    pol.eval()
    with torch.no_grad():
        pred = pol._predict_actions(batch)
    pre, post = make_fastvla_pre_post_processors(cfg, dataset_stats=None)
    assert post(pred).device.type == "cpu"


@pytest.mark.gpu
@pytest.mark.parametrize("dtype,tol", [("float32", 2e-5), ("bfloat16", 2e-2)])
def test_fused_io_normalization_and_uint8_host_ingest(dtype, tol):
    """Row f1: raw uint8 HWC camera frames + raw state in HOST memory through the fused pipelines (statistics applied
    inside the head kernel, frames scaled inside the ingest kernel, staged copy on a side stream) give the actions the
    reference-shaped pipelines give on float [0,1] CHW device tensors (reference lerobot_fastvla/processor_fastvla.py:30-48)."""
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    from vla_fastvlm.lerobot_fastvla import FastVLAConfig, FastVLAPolicy, make_fastvla_pre_post_processors

    inp, outp = _features(n_cam=1)
    common = dict(input_features=inp, output_features=outp, device="cuda", vlm_model_name="synthetic:tiny",
                  hidden_dim=TINY_HEAD["hidden_dim"], fusion_dim=TINY_HEAD["fusion_dim"], image_token_mode="prefix",
                  compute_dtype=dtype)
    plain_cfg = FastVLAConfig(**common)
    fused_cfg = FastVLAConfig(fuse_io_normalization=True, image_input_scale=1.0 / 255.0, **common)
    plain = FastVLAPolicy(plain_cfg).cuda()
    fused = FastVLAPolicy(fused_cfg).cuda()
    fused.load_state_dict(plain.state_dict(), strict=True)
    g = torch.Generator().manual_seed(7)
    stats = {"observation.state": {"mean": torch.randn(6, generator=g), "std": torch.rand(6, generator=g) + 0.5},
             "action": {"mean": torch.randn(5, generator=g), "std": torch.rand(5, generator=g) + 0.5}}
    B = 3
    frames = torch.randint(0, 256, (B, 96, 128, 3), generator=g, dtype=torch.uint8).pin_memory()   # camera: HWC uint8
    state = (torch.randn(B, 6, generator=g) * 2 + 1).pin_memory()
    pre_p, post_p = make_fastvla_pre_post_processors(plain_cfg, dataset_stats=stats)
    pre_f, post_f = make_fastvla_pre_post_processors(fused_cfg, dataset_stats=stats)
    want = post_p(plain.select_action(pre_p({"observation.images.cam0": frames.permute(0, 3, 1, 2).float() / 255.0,
                                              "observation.state": state.clone(), "task": ["push the block"]})))
    got = post_f(fused.select_action(pre_f({"observation.images.cam0": frames, "observation.state": state,
                                             "task": ["push the block"]})))
    assert got.device.type == "cpu" and want.device.type == "cpu" and got.shape == (B, 5)
    assert fused._stager is not None and fused._stager.h2d_bytes == frames.numel() + state.numel() * 4
    err = float((got - want).abs().max() / want.abs().max())
    assert err <= tol, (err, got, want)
    # a second call with new observations reuses the staging slots and the pushed statistics
    frames2 = torch.randint(0, 256, (B, 96, 128, 3), generator=g, dtype=torch.uint8).pin_memory()
    fused.reset(); plain.reset()
    want2 = post_p(plain.select_action(pre_p({"observation.images.cam0": frames2.permute(0, 3, 1, 2).float() / 255.0,
                                               "observation.state": state.clone(), "task": ["push the block"]})))
    got2 = post_f(fused.select_action(pre_f({"observation.images.cam0": frames2, "observation.state": state,
                                              "task": ["push the block"]})))
    assert float((got2 - want2).abs().max() / want2.abs().max()) <= tol and not torch.allclose(got2, got)
    # training through the fused configuration: the loss is the MSE in NORMALISED action space, as with the plain pipelines
    fused.train(); plain.train()
    act = torch.randn(B, 1, 5, generator=g)
    torch.manual_seed(0)
    loss_p, _ = plain.forward(pre_p({"observation.images.cam0": frames.permute(0, 3, 1, 2).float() / 255.0,
                                     "observation.state": state.clone(), "task": ["push the block"], "action": act}))
    torch.manual_seed(0)
    loss_f, _ = fused.forward(pre_f({"observation.images.cam0": frames, "observation.state": state,
                                     "task": ["push the block"], "action": act}))
    assert abs(float(loss_p) - float(loss_f)) <= (1e-4 if dtype == "float32" else 5e-2) * max(1.0, float(loss_p))


@pytest.mark.gpu
def test_training_step_updates_head_and_engine():
    """The data-parallel training step (world size 1 here): frozen backbone in the engine, head through autograd,
    gradients accumulated into the flat all-reduce buffer, AdamW on the head only — the loss falls, the backbone's
    tensors never get gradients, and eval-mode inference (fused head kernel) sees the updated head."""
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    from vla_fastvlm.lerobot_fastvla import FastVLAConfig, FastVLAPolicy
    from vla_fastvlm.training import HeadGradAllReduce, train_step

    inp, outp = _features(n_cam=1)
    cfg = FastVLAConfig(input_features=inp, output_features=outp, device="cuda", vlm_model_name="synthetic:tiny",
                        hidden_dim=TINY_HEAD["hidden_dim"], fusion_dim=TINY_HEAD["fusion_dim"], image_token_mode="prefix",
                        dropout=0.0)
    pol = FastVLAPolicy(cfg).cuda()
    g = torch.Generator().manual_seed(3)
    B = 4
    batch = {"observation.images.cam0": torch.rand(B, 3, 96, 128, generator=g).cuda(),
             "observation.state": torch.randn(B, 6, generator=g).cuda(), "task": ["push the block"] * B,
             "action": torch.randn(B, 1, 5, generator=g).cuda()}
    before = pol.select_action(batch).clone()
    pol.reset()
    trainable = [p for p in pol.get_optim_params() if p.requires_grad]
    red = HeadGradAllReduce(trainable)
    assert red.numel == sum(p.numel() for p in trainable) and red.world_size == 1
    opt = torch.optim.AdamW(trainable, lr=3e-3, weight_decay=0.0)
    losses = [train_step(pol, batch, opt, red)[0] for _ in range(12)]
    assert losses[-1] < 0.7 * losses[0], losses
    after = pol.select_action(batch)
    assert not torch.allclose(after, before)                                  # the engine picked up the new head
    pol.train()
    with torch.no_grad():
        ref = pol._predict_actions(batch)                                     # autograd-path head on engine features
    assert torch.allclose(after, ref, atol=2e-2, rtol=2e-2)


@pytest.mark.gpu
@pytest.mark.parametrize("drop_p", [0.0, 0.1])
def test_native_head_forward_backward_matches_autograd(drop_p):
    """fvla_head_forward_backward (csrc/head_train.cu) against torch autograd on the reference's head modules in
    train mode: same loss, same twelve gradients, written into the flat all-reduce buffer (row a13)."""
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    from torch import nn
    from torch.nn import functional as F

    from vla_fastvlm.training import HeadGradAllReduce, NativeHeadStep

    torch.manual_seed(0)
    H, S, A, Hd, Fd, B = 136, 6, 5, 72, 80, 37

    class Head(nn.Module):  # the head of FastVLMWithExpert (fastvla/fastvlm_with_expert.py:23-38), same attribute names
        def __init__(self):
            super().__init__()
            self.state_projection = nn.Sequential(nn.LayerNorm(S), nn.Linear(S, Hd), nn.SiLU())
            self.fusion = nn.Sequential(nn.Linear(H + Hd, Fd), nn.LayerNorm(Fd), nn.SiLU(), nn.Dropout(drop_p),
                                        nn.Linear(Fd, Fd), nn.SiLU())
            self.action_head = nn.Linear(Fd, A)

    head = Head().cuda()
    for p in head.parameters():  # non-trivial norm weights / biases
        if p.ndim == 1:
            p.data.uniform_(-0.5, 1.5)
    params = list(head.parameters())
    pooled, states, target = torch.randn(B, H).cuda(), torch.randn(B, S).cuda() * 2, torch.randn(B, A).cuda()
    keep = (torch.rand(B, Fd, device="cuda") >= drop_p).to(torch.uint8)
    # autograd with the SAME mask (Dropout replaced by the explicit mask, as nn.Dropout scales by 1/(1-p))
    s = head.state_projection(states)
    f = head.fusion[2](head.fusion[1](head.fusion[0](torch.cat([pooled, s], -1))))
    if drop_p > 0:
        f = f * keep.float() / (1.0 - drop_p)
    pred = head.action_head(head.fusion[5](head.fusion[4](f)))
    loss = F.mse_loss(pred, target)
    want = torch.autograd.grad(loss, params)

    red = HeadGradAllReduce(params)
    step = NativeHeadStep(head, red)
    red.flat.fill_(123.0)  # the kernels WRITE the gradients: no dependence on what was there
    got_loss = step(pooled, states, target, train=True, keep_mask=keep)
    torch.cuda.synchronize()
    assert abs(float(got_loss) - float(loss)) <= 1e-5 * max(1.0, abs(float(loss)))
    off = 0
    for p, w in zip(params, want):
        g = red.flat[off:off + p.numel()].view_as(p)
        off += p.numel()
        assert p.grad.data_ptr() == g.data_ptr()                       # .grad IS the slice of the flat buffer
        scale = float(w.abs().max()) + 1e-12
        assert float((g - w).abs().max()) <= 2e-5 * max(scale, 1e-3), (tuple(p.shape), float((g - w).abs().max()), scale)
    assert off == red.numel


@pytest.mark.gpu
def test_native_training_step_through_the_plugin():
    """train_step_native: engine backbone -> native head fwd/bwd -> flat buffer -> (world-1) all-reduce -> AdamW; the
    loss falls and matches the autograd step's trajectory at dropout 0."""
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    from vla_fastvlm.lerobot_fastvla import FastVLAConfig, FastVLAPolicy
    from vla_fastvlm.training import HeadGradAllReduce, NativeHeadStep, train_step, train_step_native

    inp, outp = _features(n_cam=1)
    cfg = FastVLAConfig(input_features=inp, output_features=outp, device="cuda", vlm_model_name="synthetic:tiny",
                        hidden_dim=TINY_HEAD["hidden_dim"], fusion_dim=TINY_HEAD["fusion_dim"], image_token_mode="prefix",
                        dropout=0.0)
    g = torch.Generator().manual_seed(3)
    B = 4
    batch = {"observation.images.cam0": torch.rand(B, 3, 96, 128, generator=g).cuda(),
             "observation.state": torch.randn(B, 6, generator=g).cuda(), "task": ["push the block"] * B,
             "action": torch.randn(B, 1, 5, generator=g).cuda()}
    losses = {}
    for mode in ("autograd", "native"):
        torch.manual_seed(11)
        pol = FastVLAPolicy(cfg).cuda()
        trainable = [p for p in pol.get_optim_params() if p.requires_grad]
        red = HeadGradAllReduce(trainable)
        opt = torch.optim.AdamW(trainable, lr=3e-3, weight_decay=0.0)
        if mode == "native":
            step = NativeHeadStep(pol.model, red)
            losses[mode] = [train_step_native(pol, batch, opt, red, step)[0] for _ in range(8)]
        else:
            losses[mode] = [train_step(pol, batch, opt, red)[0] for _ in range(8)]
    assert losses["native"][-1] < 0.7 * losses["native"][0], losses
    for a, b in zip(losses["native"], losses["autograd"]):
        assert abs(a - b) <= 2e-3 * max(1.0, abs(b)), losses
