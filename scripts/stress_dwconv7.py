"""Race / hazard stress for the warp-specialised 7x7 pipeline: many launches on fresh random data, the pipeline kernel
(default) against the round-1 kernels (FVLA_DISABLE_DWCONV7_R4 / FVLA_DISABLE_DWCONV7_S2_MMA read once per process, so
the comparison output comes from a float64 conv2d on the bf16-rounded taps).  python scripts/stress_dwconv7.py [iters]"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1] / "vla-from-fastvlm_b200"))
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from vla_fastvlm import _native as N  # noqa: E402

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 40
worst = 0.0
g = torch.Generator(device="cuda").manual_seed(0)
for it in range(iters):
    for (B, H, W, C, stride, mult) in [(8, 128, 128, 192, 1, 1), (16, 64, 64, 384, 1, 1), (4, 256, 256, 96, 2, 2),
                                       (5, 64, 96, 48, 2, 2), (3, 32, 32, 768, 1, 1)]:
        x = torch.randn(B, H, W, C, device="cuda", generator=g).bfloat16()
        w = (torch.randn(C * mult, 1, 7, 7, device="cuda", generator=g) / 7)
        b = torch.randn(C * mult, device="cuda", generator=g)
        wp = w.reshape(C * mult, 49).t().contiguous()
        out = N.op_dwconv(x, wp, b, 7, stride, mult, 0).double()
        ref = F.conv2d(x.double().permute(0, 3, 1, 2), w.bfloat16().double(), b.double(), stride=stride, padding=3,
                       groups=C).permute(0, 2, 3, 1)
        err = float(((out - ref).abs() / (ref.abs() * 2.0 ** -8 + 2e-3)).max())   # one bf16 rounding + fp32 summation slack
        worst = max(worst, err)
        assert err <= 1.0, (it, B, H, W, C, stride, err)
print(f"ok: {iters} iterations x 5 shapes, worst error / bound = {worst:.3f}")
