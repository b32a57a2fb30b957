"""The oracle, run stand-alone, reproduces the golden outputs that tests/golden/make_golden.py recorded
from the REAL reference classes (adapter + head + policy around the restated VLM)."""
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent / "golden"))
from cases import CASES, TINY_HEAD, case_inputs, checksum  # noqa: E402
from helpers import tiny_weights  # noqa: E402

GOLD = Path(__file__).resolve().parent / "golden"


def load_case(name):
    return dict(np.load(GOLD / f"tiny_{name}.npz", allow_pickle=False))


def oracle_run(name):
    from oracle.fastvla_oracle import IMAGE_TOKEN_INDEX, FastVLAOracle, prepare_images
    from vla_fastvlm.model.tokenizer import SimpleByteTokenizer
    from vla_fastvlm.shared import as_prompt_list, pick_step

    spec = CASES[name]
    arch, sd, hsd = tiny_weights(0)
    images, states, tasks = case_inputs(name)
    images_b = pick_step(images, 4, -1)
    states_b = pick_step(states, 2, -1)
    prompts = as_prompt_list(tasks, images_b.shape[0], True)
    enc = SimpleByteTokenizer(arch.text.vocab)(prompts, padding="longest", truncation=True, max_length=64)
    ids, mask = enc["input_ids"], enc["attention_mask"]
    if spec["prefix"]:
        b = ids.shape[0]
        ids = torch.cat([torch.full((b, 1), IMAGE_TOKEN_INDEX, dtype=torch.long), ids], 1)
        mask = torch.cat([torch.ones(b, 1, dtype=torch.long), mask], 1)
    oracle = FastVLAOracle(arch, sd, hsd, pool_mode=spec["pool"])
    taps = {}
    actions = oracle.forward(images_b, states_b, ids, mask, taps=taps)
    return dict(images=images, states=states, ids=ids, mask=mask, actions=actions, taps=taps)


def test_weight_fingerprint_is_stable():
    arch, sd, hsd = tiny_weights(0)
    fp = float(sum(v.double().abs().sum() for v in sd.values()) + sum(v.double().abs().sum() for v in hsd.values()))
    want = float((GOLD / "tiny_weights_fingerprint.txt").read_text())
    assert abs(fp - want) < 1e-3, "seeded synthetic init drifted: regenerate tests/golden with make_golden.py"


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_reproduces_reference_outputs(name):
    gold = load_case(name)
    r = oracle_run(name)
    assert abs(checksum(r["images"]) - float(gold["images_checksum"])) < 1e-6 * max(1.0, abs(float(gold["images_checksum"])))
    assert np.array_equal(r["ids"].numpy(), gold["input_ids"])
    assert np.array_equal(r["mask"].numpy(), gold["attention_mask"])
    pixel = r["taps"]["preprocess"]
    assert np.allclose(pixel[:, :, ::8, ::8].numpy(), gold["pixel_probe"], rtol=0, atol=1e-6)
    assert abs(checksum(pixel) - float(gold["pixel_checksum"])) <= 1e-6 * max(1.0, abs(float(gold["pixel_checksum"])))
    assert np.allclose(r["taps"]["pooled"].numpy(), gold["pooled"], rtol=1e-5, atol=1e-5)
    assert np.allclose(r["actions"].numpy(), gold["actions"], rtol=1e-5, atol=1e-5)
