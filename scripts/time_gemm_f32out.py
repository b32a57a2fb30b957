"""A/B of the fp32 stream-update epilogue: TMA reduce-add into D vs a plain store (no residual), o-proj / down-proj shapes."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1] / "vla-from-fastvlm_b200"))
import torch  # noqa: E402

from vla_fastvlm import _native as N  # noqa: E402

lib, sp = N.load(), N.stream_ptr()
for (M, Nn, K) in [(17408, 896, 896), (17408, 896, 4864), (17408, 1152, 896)]:
    a = torch.randn(M, K, device="cuda").bfloat16()
    w = (torch.randn(Nn, K, device="cuda") / K ** 0.5).bfloat16()
    d = torch.zeros(M, Nn, device="cuda", dtype=torch.float32)
    for name, resid, flags in (("reduce-add", d, 4), ("plain store", None, 4), ("bf16 out", None, 0)):
        out = d if flags else torch.empty(M, Nn, device="cuda", dtype=torch.bfloat16)

        def run():
            N.check(lib.fvla_op_gemm(N.FVLA_BF16, N.ptr(a), K, N.ptr(w), K, N.ptr(out), Nn, M, Nn, K, None, None,
                                     N.ptr(resid) if resid is not None else None, Nn, 0, flags, 0, sp), "gemm")
        for _ in range(3):
            run()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            run()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        print(f"M={M} N={Nn} K={K} {name:12s}: {ms * 1e3:7.1f} us  {2.0 * M * Nn * K / ms / 1e9:7.1f} TFLOP/s")
