"""Time + check the fused ConvFFN kernel at the FastViTHD stage shapes: python scripts/time_ffn.py [C ...]
(set FVLA_FFN_NO_TEAMS=1 for the lock-step epilogue).  CUDA events, 20 launches after 3 warm-ups, inputs > L2."""
import math
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1] / "vla-from-fastvlm_b200"))
import torch  # noqa: E402

from vla_fastvlm import _native as N  # noqa: E402

shapes = {96: 32 * 256 * 256, 192: 32 * 128 * 128, 384: 32 * 64 * 64}
for Cc in [int(a) for a in sys.argv[1:]] or [96, 192]:
    M, hidden = shapes[Cc], 4 * Cc
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn(M, Cc, device="cuda", generator=g).bfloat16()
    w1 = (torch.randn(hidden, Cc, device="cuda", generator=g) / math.sqrt(Cc)).bfloat16()
    b1 = torch.randn(hidden, device="cuda", generator=g) * 0.5
    w2 = (torch.randn(Cc, hidden, device="cuda", generator=g) / math.sqrt(hidden)).bfloat16()
    b2 = torch.randn(Cc, device="cuda", generator=g)
    res = torch.randn(M, Cc, device="cuda", generator=g).bfloat16()
    w1h, b1h = (w1.float() * 0.5).bfloat16().contiguous(), (b1 * 0.5).contiguous()   # packed as the engine packs them
    w2c, out = w2.half().contiguous(), torch.empty_like(x)
    lib, sp = N.load(), N.stream_ptr()

    def run():
        N.check(lib.fvla_op_ffn_fused(N.ptr(x), N.ptr(w1h), N.ptr(b1h), N.ptr(w2c), N.ptr(b2), N.ptr(res), N.ptr(out),
                                      M, Cc, hidden, sp), "fvla_op_ffn_fused")

    for _ in range(3):
        run()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    n = 4096  # fp32 reference on the first and last 4096 rows

    def ref_rows(sl):
        return res[sl].float() + torch.nn.functional.gelu(x[sl].float() @ w1.float().t() + b1) @ w2.float().t() + b2

    ref = ref_rows(slice(0, n))
    err = float((out[:n].float() - ref).abs().max() / ref.abs().max())
    tail = float((out[-n:].float() - ref_rows(slice(M - n, M))).abs().max() / ref.abs().max())
    print(f"C={Cc} M={M}: {ms:.4f} ms  {4.0 * M * hidden * Cc / ms / 1e9:.1f} TFLOP/s  rel err head {err:.2e} tail {tail:.2e}")
