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
