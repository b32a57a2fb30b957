"""One bf16 GEMM launch (for ncu): python scripts/one_gemm.py M N K [gelu|gelu16|res|res32] [block_n]
res32 = the decoder's in-place FP32 residual-stream epilogue."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1] / "vla-from-fastvlm_b200"))
import torch  # noqa: E402

from vla_fastvlm import _native as N  # noqa: E402

M, Nn, K = (int(x) for x in sys.argv[1:4])
mode = sys.argv[4] if len(sys.argv) > 4 else ""
act = {"gelu": 1, "gelu16": 5}.get(mode, 0)
use_res = mode == "res"
a = torch.randn(M, K, device="cuda").bfloat16()
w = (torch.randn(Nn, K, device="cuda") / K ** 0.5).bfloat16()
bias = torch.randn(Nn, device="cuda")
out = torch.empty(M, Nn, device="cuda", dtype=torch.float16 if act == 5 else torch.bfloat16)
res = torch.randn(M, Nn, device="cuda").bfloat16() if use_res else None
bn = int(sys.argv[5]) if len(sys.argv) > 5 else 0
if mode == "res32":
    out = torch.randn(M, Nn, device="cuda")
    res, bias = out, None
import time
for _ in range(3):
    N.op_gemm(a, w, bias=bias, act=act, resid=res, out=out, block_n=bn)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    N.op_gemm(a, w, bias=bias, act=act, resid=res, out=out, block_n=bn)
e1.record()
torch.cuda.synchronize()
print(f"avg {e0.elapsed_time(e1) / 20 * 1e3:.1f} us  {2.0 * M * Nn * K / (e0.elapsed_time(e1) / 20) / 1e9:.0f} TFLOP/s")
torch.cuda.synchronize()
print("ok", float(out.float().abs().mean()))
