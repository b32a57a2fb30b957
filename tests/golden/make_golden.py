"""Generate the golden fixtures in tests/golden/ by running the REAL reference classes
(`/root/reference/src/vla_fastvlm`: FastVLMBackbone adapter + FastVLMWithExpert head + FastVLAPolicy)
around the oracle's restated `LlavaQwen2ForCausalLM` (the remote-code VLM cannot be fetched offline).

What this pins: everything that lives in the reference tree — image canonicalisation, tokenizer call
protocol, the VLM call protocol, pooling, the action head, batch/time-step handling — is executed by
the reference's own code; only the VLM body is the restatement (see oracle/fastvla_oracle.py header).

Runs in its own process because the reference package and the product package share the name
`vla_fastvlm`; product modules needed here (arch presets, synthetic weights, byte tokenizer) are
loaded by file path.

    python tests/golden/make_golden.py [--out tests/golden]
"""
from __future__ import annotations

import argparse
import importlib.util
import sys
from pathlib import Path
from types import SimpleNamespace

import numpy as np
import torch
from torch import nn

ROOT = Path(__file__).resolve().parents[2]
REF_SRC = Path("/root/reference/src")
PKG = ROOT / "vla-from-fastvlm_b200" / "vla_fastvlm"


def _load(name: str, path: Path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=str(ROOT / "tests" / "golden"))
    args = ap.parse_args()
    out_dir = Path(args.out)
    out_dir.mkdir(parents=True, exist_ok=True)
    if not REF_SRC.is_dir():
        raise SystemExit("reference sources not found at /root/reference/src")

    torch.manual_seed(0)
    torch.set_num_threads(4)
    sys.path.insert(0, str(ROOT))
    arch_mod = _load("fvla_arch", PKG / "model" / "arch.py")
    syn_mod = _load("fvla_synthetic", PKG / "model" / "synthetic.py")
    tok_mod = _load("fvla_tokenizer", PKG / "model" / "tokenizer.py")
    from oracle.fastvla_oracle import IMAGE_TOKEN_INDEX, FastVLAOracle

    arch = arch_mod.PRESETS["tiny"]
    head = dict(state_dim=6, action_dim=5, hidden_dim=64, fusion_dim=64)
    sd = syn_mod.synthetic_backbone_state_dict(arch, 0)
    hsd = syn_mod.synthetic_head_state_dict(arch.text.hidden, head["state_dim"], head["action_dim"],
                                            head["hidden_dim"], head["fusion_dim"], 1)
    oracle = FastVLAOracle(arch, sd, hsd)

    class OracleVLM(nn.Module):
        """Stands where AutoModelForCausalLM.from_pretrained(...) would return LlavaQwen2ForCausalLM."""

        def __init__(self):
            super().__init__()
            self.anchor = nn.Parameter(torch.zeros(1))
            self.config = SimpleNamespace(hidden_size=arch.text.hidden, mm_vision_tower=arch.mm_vision_tower,
                                          output_hidden_states=False)
            self.calls = []

        def forward(self, input_ids=None, attention_mask=None, images=None, output_hidden_states=None,
                    return_dict=None):
            self.calls.append(tuple(images.shape))
            hidden, _ = oracle.vlm_hidden(images, input_ids, attention_mask)
            return SimpleNamespace(hidden_states=(hidden,))  # CausalLMOutputWithPast has no last_hidden_state

    class PrefixTokenizer(tok_mod.SimpleByteTokenizer):
        """Byte tokenizer that puts the LLaVA image placeholder first (north_star 'prefix' mode)."""

        prefix = False

        def __call__(self, texts, **kw):
            out = super().__call__(texts, **kw)
            if self.prefix:
                b = out["input_ids"].shape[0]
                out["input_ids"] = torch.cat([torch.full((b, 1), IMAGE_TOKEN_INDEX, dtype=torch.long), out["input_ids"]], 1)
                out["attention_mask"] = torch.cat([torch.ones(b, 1, dtype=torch.long), out["attention_mask"]], 1)
            return out

    # ---- import the reference and point its Auto* loaders at the stand-ins ----
    sys.path.insert(0, str(REF_SRC))
    import vla_fastvlm.model.fastvlm_adapter as ref_adapter  # noqa: E402  (the REFERENCE package)
    from vla_fastvlm.fastvla.configuration_fastvla import FastVLAConfig as RefConfig  # noqa: E402
    from vla_fastvlm.fastvla.modeling_fastvla import FastVLAPolicy as RefPolicy  # noqa: E402

    assert str(Path(ref_adapter.__file__)).startswith(str(REF_SRC)), ref_adapter.__file__
    vlm = OracleVLM()
    tokenizer = PrefixTokenizer(arch.text.vocab)

    def _raise(*a, **k):
        raise OSError("offline")

    ref_adapter.AutoModelForCausalLM = SimpleNamespace(from_pretrained=lambda *a, **k: vlm)
    ref_adapter.AutoTokenizer = SimpleNamespace(from_pretrained=lambda *a, **k: tokenizer)
    ref_adapter.AutoProcessor = SimpleNamespace(from_pretrained=_raise)
    ref_adapter.AutoImageProcessor = SimpleNamespace(from_pretrained=_raise)

    def build_policy(pool="last_token"):
        cfg = RefConfig(vlm_model_name="tiny", **head)
        pol = RefPolicy(cfg)
        pol.model.backbone.config.image_feature_pool = pool
        missing = pol.model.load_state_dict(hsd, strict=False)
        assert not [k for k in missing.unexpected_keys], missing
        assert all(k.startswith("backbone.") for k in missing.missing_keys), missing
        pol.eval()
        assert pol.model.backbone.expected_size == arch.vision.image_size
        return pol

    sys.path.insert(0, str(ROOT / "tests" / "golden"))
    from cases import CASES, case_inputs, checksum  # noqa: E402

    cases = {}
    for name, spec in CASES.items():
        images, states, tasks = case_inputs(name)
        tokenizer.prefix = spec["prefix"]
        pol = build_policy(spec["pool"])
        with torch.no_grad():
            actions = pol.forward(images, states, tasks, device=torch.device("cpu"))
            pixel = pol.model.backbone._prepare_images_tensor(
                images[:, -1] if images.ndim == 5 else images, torch.device("cpu"))
            norm_tasks = pol.processor.prepare_tasks(tasks, batch_size=pixel.shape[0])
            tok = pol.model.backbone._prep_text(norm_tasks, torch.device("cpu"))
            pooled = pol.model.backbone(pixel, norm_tasks, device=torch.device("cpu"))
        cases[name] = dict(
            images_checksum=np.array(checksum(images)), states_checksum=np.array(checksum(states)),
            input_ids=tok["input_ids"].numpy(), attention_mask=tok["attention_mask"].numpy(),
            pixel_probe=pixel[:, :, ::8, ::8].contiguous().numpy(), pixel_checksum=np.array(checksum(pixel)),
            pooled=pooled.numpy(), actions=actions.numpy())
        print(f"{name:28s} images {tuple(images.shape)} T={tok['input_ids'].shape[1]} actions[0]={actions[0, :3].tolist()}")

    for name, c in cases.items():
        np.savez_compressed(out_dir / f"tiny_{name}.npz", **c)
    # weight fingerprint: detects drift of the seeded init between torch versions / machines
    fp = float(sum(v.double().abs().sum() for v in sd.values()) + sum(v.double().abs().sum() for v in hsd.values()))
    (out_dir / "tiny_weights_fingerprint.txt").write_text(f"{fp:.6f}\n")
    print("weights fingerprint", fp)


if __name__ == "__main__":
    main()
