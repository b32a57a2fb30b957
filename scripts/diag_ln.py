"""Diagnostic: LayerNormChannel attention-norm variant on the tiny tower, errors per tap (bf16)."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
for p in (ROOT, ROOT / "vla-from-fastvlm_b200", ROOT / "tests"):
    sys.path.insert(0, str(p))
import torch
from helpers import TINY_HEAD, make_engine, make_inputs, rel_err, tiny_weights
from oracle.fastvla_oracle import FastVLAOracle
from vla_fastvlm import _native as N
from vla_fastvlm.model.synthetic import synthetic_backbone_state_dict

for norm in ("layernorm", "batchnorm"):
    arch, sd0, hsd = tiny_weights(0)
    sd = synthetic_backbone_state_dict(arch, 0, attn_norm=norm) if norm == "layernorm" else sd0
    for seed in (3, 4, 5):
        images, states, ids, mask = make_inputs(3, 120, 160, 9, arch.text.vocab, TINY_HEAD["state_dim"], seed=seed)
        taps = {}
        ref = FastVLAOracle(arch, sd, hsd).forward(images, states, ids, mask, taps=taps)
        eng = make_engine(arch, sd, hsd, torch.bfloat16)
        dev, v = eng.device, arch.vision
        bufs = {}
        side = v.image_size // 4
        for i, d in enumerate(v.dims):
            bufs[i] = torch.zeros(3, side, side, d, device=dev, dtype=torch.bfloat16)
            eng.set_tap(N.TAP_VIS_STAGE0 + i, bufs[i])
            side //= 2
        feats = torch.zeros(3, v.num_tokens, v.out_channels, device=dev, dtype=torch.bfloat16)
        eng.set_tap(N.TAP_IMAGE_FEATURES, feats)
        out = eng.forward(images.to(dev), ids, mask.sum(1), states=states.to(dev)).float().cpu()
        errs = {f"s{i}": round(rel_err(bufs[i], taps[f"vis_stage{i}"].permute(0, 2, 3, 1)), 4) for i in range(5)}
        errs["feats"] = round(rel_err(feats, taps["image_features"]), 4)
        errs["act"] = round(rel_err(out, ref), 4)
        print(norm, seed, errs, flush=True)
