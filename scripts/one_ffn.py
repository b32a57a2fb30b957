"""One fused-ConvFFN launch (for ncu): python scripts/one_ffn.py M C"""
import math
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1] / "vla-from-fastvlm_b200"))
import torch  # noqa: E402

from vla_fastvlm import _native as N  # noqa: E402

M, Cc = int(sys.argv[1]), int(sys.argv[2])
hidden = 4 * Cc
x = torch.randn(M, Cc, device="cuda").bfloat16()
w1 = (torch.randn(hidden, Cc, device="cuda") / math.sqrt(Cc)).bfloat16()
b1 = torch.randn(hidden, device="cuda") * 0.5
w2 = (torch.randn(Cc, hidden, device="cuda") / math.sqrt(hidden)).bfloat16()
b2 = torch.randn(Cc, device="cuda")
res = torch.randn(M, Cc, device="cuda").bfloat16()
for _ in range(3):
    out = N.op_ffn_fused(x, w1, b1, w2, b2, res)
torch.cuda.synchronize()
print("ok", float(out.float().abs().mean()))
