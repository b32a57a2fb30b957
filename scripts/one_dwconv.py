"""One depthwise-conv launch (for ncu): python scripts/one_dwconv.py B HW C k stride mult"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1] / "vla-from-fastvlm_b200"))
import torch  # noqa: E402

from vla_fastvlm import _native as N  # noqa: E402

B, HW, C, k, stride, mult = (int(x) for x in sys.argv[1:7])
x = torch.randn(B, HW, HW, C, device="cuda").bfloat16()
w = torch.randn(k * k, C * mult, device="cuda") / k
b = torch.randn(C * mult, device="cuda")
for _ in range(3):
    out = N.op_dwconv(x, w, b, k, stride, mult, 0)
torch.cuda.synchronize()
print("ok", float(out.float().abs().mean()))
