"""Micro-benchmarks of single kernels at FastViTHD / Qwen2 shapes (CUDA events, L2-cold via rotation)."""
import argparse
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1] / "vla-from-fastvlm_b200"))
import torch  # noqa: E402

from vla_fastvlm import _native as N  # noqa: E402


def timeit(fn, iters=10, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def bench_dwconv(B):
    print(f"# dwconv bf16, batch {B}")
    for (C, HW, k, stride, mult) in [(96, 256, 7, 1, 1), (192, 128, 7, 1, 1), (384, 64, 7, 1, 1), (768, 32, 7, 1, 1),
                                     (96, 256, 3, 1, 1), (192, 128, 3, 1, 1), (384, 64, 3, 1, 1),
                                     (96, 256, 7, 2, 2), (96, 512, 3, 2, 1)]:
        x = torch.randn(B, HW, HW, C, device="cuda").bfloat16()
        w = torch.randn(k * k, C * mult, device="cuda")
        b = torch.randn(C * mult, device="cuda")
        ms = timeit(lambda: N.op_dwconv(x, w, b, k, stride, mult, 0))
        Ho = (HW - 1) // stride + 1
        flops = 2.0 * k * k * B * Ho * Ho * C * mult
        byts = 2.0 * (x.numel() + B * Ho * Ho * C * mult)
        print(f"C{C:5d} HW{HW:4d} k{k} s{stride} m{mult}: {ms:8.3f} ms  {flops / ms / 1e9:7.1f} TFLOP/s  {byts / ms / 1e6:8.1f} GB/s")


def bench_gemm(B):
    print(f"# gemm bf16, batch {B}")
    shapes = [(B * 65536, 384, 96, 1, 0), (B * 65536, 96, 384, 0, 1), (B * 16384, 768, 192, 1, 0), (B * 16384, 192, 768, 0, 1),
              (B * 4096, 1536, 384, 1, 0), (B * 4096, 384, 1536, 0, 1), (B * 1024, 3072, 768, 1, 0), (B * 1024, 768, 3072, 0, 1),
              (B * 1024, 2304, 768, 0, 0), (B * 272 * 8, 1152, 896, 0, 0), (B * 272 * 8, 896, 896, 0, 1),
              (B * 272 * 8, 9728, 896, 2, 0), (B * 272 * 8, 896, 4864, 0, 1), (8192, 8192, 8192, 0, 0)]
    for (M, Nn, K, mode, res) in shapes:
        a = torch.randn(M, K, device="cuda").bfloat16()
        w = (torch.randn(Nn, K, device="cuda") / K ** 0.5).bfloat16()
        bias = torch.randn(Nn, device="cuda") if mode != 2 else None
        n_out = Nn // 2 if mode == 2 else Nn
        out = torch.empty(M, n_out, device="cuda", dtype=torch.bfloat16)
        r = torch.randn(M, Nn, device="cuda").bfloat16() if res else None
        ms = timeit(lambda: N.op_gemm(a, w, bias=bias, resid=r, act=1 if mode == 1 else 0, swiglu=(mode == 2), out=out))
        flops = 2.0 * M * Nn * K
        byts = 2.0 * (M * K + Nn * K + M * n_out + (M * Nn if res else 0))
        print(f"M{M:8d} N{Nn:5d} K{K:5d} {'gelu' if mode == 1 else ('swiglu' if mode == 2 else '    ')} {'+res' if res else '    '}: "
              f"{ms:8.3f} ms {flops / ms / 1e9:7.1f} TFLOP/s {byts / ms / 1e6:8.1f} GB/s")


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--what", default="dwconv,gemm")
    ap.add_argument("--batch", type=int, default=8)
    a = ap.parse_args()
    if "dwconv" in a.what:
        bench_dwconv(a.batch)
    if "gemm" in a.what:
        bench_gemm(a.batch)


def bench_gemm_epilogues():
    print("# epilogue cost at fixed shape (bf16)")
    for (M, Nn, K) in [(524288, 768, 192), (131072, 1536, 384), (131072, 384, 1536)]:
        a = torch.randn(M, K, device="cuda").bfloat16()
        w = (torch.randn(Nn, K, device="cuda") / K ** 0.5).bfloat16()
        bias = torch.randn(Nn, device="cuda")
        out = torch.empty(M, Nn, device="cuda", dtype=torch.bfloat16)
        r = torch.randn(M, Nn, device="cuda").bfloat16()
        for name, kw in [("plain", {}), ("bias", dict(bias=bias)), ("bias+gelu", dict(bias=bias, act=1)),
                         ("bias+silu", dict(bias=bias, act=2)), ("bias+res", dict(bias=bias, resid=r))]:
            for bn in (0, 128, 256):
                ms = timeit(lambda: N.op_gemm(a, w, out=out, block_n=bn, **kw))
                print(f"M{M} N{Nn} K{K} {name:10s} bn={bn:3d}: {ms:7.3f} ms {2.0 * M * Nn * K / ms / 1e9:7.1f} TFLOP/s "
                      f"{2.0 * (M * K + M * Nn) / ms / 1e6:7.1f} GB/s")


if __name__ == "__main__" and "epi" in sys.argv:
    bench_gemm_epilogues()
