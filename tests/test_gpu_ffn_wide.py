"""The opt-in C = 384 fused ConvFFN (csrc/ffn_wide_sm100.cu, FVLA_ENABLE_FFN_WIDE=1; the switch is read once per process,
so the check runs in a child): fc1 -> GELU -> fc2 -> +residual with the hidden tensor on chip against fp64 math, ragged
and multi-tile M, and the default path untouched when the switch is off."""
import os
import re
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]


@pytest.mark.gpu
def test_ffn_wide_matches_fp64_math():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    env = dict(os.environ, FVLA_ENABLE_FFN_WIDE="1")
    env.pop("FVLA_FFN_WIDE_DEBUG", None)
    r = subprocess.run([sys.executable, str(ROOT / "scripts" / "one_ffn_wide.py"), "37", "300", "4096", "20000"], env=env,
                       capture_output=True, text=True, timeout=280)
    assert r.returncode == 0, r.stderr[-2000:]
    errs = re.findall(r"M=(\d+): max rel err ([0-9.e+-]+)\s+finite=(\w+)", r.stdout)
    assert [int(m) for m, _, _ in errs] == [37, 300, 4096, 20000], r.stdout
    for m, e, fin in errs:
        assert fin == "True" and float(e) < 8e-3, (m, e, fin)   # bf16 output rounding: 2^-8 of the row maximum


@pytest.mark.gpu
def test_ffn_wide_is_opt_in(native):
    import torch

    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    if os.environ.get("FVLA_ENABLE_FFN_WIDE"):
        pytest.skip("switch set for this process")
    x = torch.zeros(256, 384, device="cuda", dtype=torch.bfloat16)
    w1 = torch.zeros(1536, 384, device="cuda", dtype=torch.bfloat16)
    w2 = torch.zeros(384, 1536, device="cuda", dtype=torch.bfloat16)
    out = native.op_ffn_fused(x, w1, torch.zeros(1536, device="cuda"), w2, torch.ones(384, device="cuda"), x)
    torch.cuda.synchronize()
    assert torch.equal(out.float(), torch.ones_like(out).float())   # the kernel itself is always callable at op level
