"""Launch one depthwise-conv shape a few times (ncu target)."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1] / "vla-from-fastvlm_b200"))
import torch  # noqa: E402

from vla_fastvlm import _native as N  # noqa: E402

B, HW, C, k = [int(v) for v in sys.argv[1:5]]
x = torch.randn(B, HW, HW, C, device="cuda").bfloat16()
w = torch.randn(k * k, C, device="cuda")
b = torch.randn(C, device="cuda")
for _ in range(4):
    out = N.op_dwconv(x, w, b, k, 1, 1, 0)
torch.cuda.synchronize()
print("ok", float(out.float().abs().mean()))
