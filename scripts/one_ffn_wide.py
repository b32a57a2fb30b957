"""C = 384 fused ConvFFN (ffn_wide_sm100.cu) against fp64 math, then timed against the fc1 + fc2 GEMM pair:
  FVLA_ENABLE_FFN_WIDE=1 python scripts/one_ffn_wide.py [M ...]"""
import math
import os
import sys
from pathlib import Path

os.environ.setdefault("FVLA_ENABLE_FFN_WIDE", "1")
sys.path.insert(0, str(Path(__file__).resolve().parents[1] / "vla-from-fastvlm_b200"))
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from vla_fastvlm import _native as N  # noqa: E402

Cc, hidden = 384, 1536
Ms = [int(a) for a in sys.argv[1:]] or [256, 300, 4096]
for M in Ms:
    g = torch.Generator().manual_seed(M)
    x = torch.randn(M, Cc, generator=g).cuda().bfloat16()
    w1 = (torch.randn(hidden, Cc, generator=g) / math.sqrt(Cc)).cuda().bfloat16()
    b1 = (torch.randn(hidden, generator=g) * 0.5).cuda()
    w2 = (torch.randn(Cc, hidden, generator=g) / math.sqrt(hidden)).cuda().bfloat16()
    b2 = torch.randn(Cc, generator=g).cuda()
    res = torch.randn(M, Cc, generator=g).cuda().bfloat16()
    out = N.op_ffn_fused(x, w1, b1, w2, b2, res)
    torch.cuda.synchronize()
    if M <= 40000:
        h = F.gelu(x.double() @ w1.double().t() + b1.double()).half().double()
        ref = (h @ w2.double().t() + b2.double() + res.double()).float()
        err = (out.float() - ref).abs().max().item() / ref.abs().max().item()
        print(f"M={M}: max rel err {err:.3e}  finite={bool(torch.isfinite(out.float()).all())}", flush=True)
    if M >= 4096:
        w1h = (w1.float() * 0.5).bfloat16()
        w2h = w2.half()
        hid = torch.empty(M, hidden, device="cuda", dtype=torch.float16)
        o2 = res.clone()
        def pair():
            N.op_gemm(x, w1h, bias=b1 * 0.5, act=5, out=hid)
            N.op_gemm(hid, w2h, bias=b2, resid=o2, out=o2)
        def fused():
            N.op_ffn_fused(x, w1, b1, w2, b2, res, out=out)
        for name, fn in (("gemm pair", pair), ("fused", fused)):
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                fn()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 10
            print(f"M={M} {name}: {ms * 1e3:.1f} us  {4.0 * M * hidden * Cc / ms / 1e9:.0f} TFLOP/s", flush=True)
