"""Programmatic dependent launch must not change results: the same forwards with the launch attribute on (default)
and off (FVLA_DISABLE_PDL=1, read once per process) are bit-identical — eager launches and the CUDA-graph replay of
small batches, fp32 and bf16.  A kernel that touched its inputs before griddepcontrol.wait would show up here as a
mismatch between the two processes or between the repeated calls of one."""
import os
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]

_CHILD = r"""
import sys, numpy as np, torch
for p in (r"{root}", r"{root}/vla-from-fastvlm_b200", r"{root}/tests", r"{root}/tests/golden"):
    sys.path.insert(0, p)
from helpers import TINY_HEAD, make_engine, make_inputs, tiny_weights
arch, sd, hsd = tiny_weights(0)
eng = make_engine(arch, sd, hsd, getattr(torch, "{dtype}"))
outs = []
for B in (1, 3, 12):
    images, states, ids, mask = make_inputs(B, 120, 160, 9, arch.text.vocab, TINY_HEAD["state_dim"], seed=3,
                                            image_mode="prefix")
    rep = []
    for _ in range(4):   # eager, graph capture (B <= 8), replays
        rep.append(eng.forward(images.to(eng.device), ids, mask.sum(1), states=states.to(eng.device)).float().cpu())
    assert all(torch.equal(rep[0], r) for r in rep[1:]), "run-to-run difference at B=%d" % B
    outs.append(rep[0].numpy().reshape(-1))
np.save(r"{out}", np.concatenate(outs))
"""


def _run(tmp_path, tag, dtype, disable):
    out = tmp_path / f"{tag}.npy"
    env = dict(os.environ)
    env.pop("FVLA_DISABLE_PDL", None)
    if disable:
        env["FVLA_DISABLE_PDL"] = "1"
    code = _CHILD.format(root=str(ROOT), dtype=dtype, out=str(out))
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    return np.load(out)


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", ["bfloat16", "float32"])
def test_pdl_on_off_bit_identical(tmp_path, dtype):
    import torch

    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    on = _run(tmp_path, "on", dtype, False)
    off = _run(tmp_path, "off", dtype, True)
    assert np.isfinite(on).all() and on.size > 0
    assert np.array_equal(on, off)
