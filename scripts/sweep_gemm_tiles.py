"""Tile-width / split-K sweep of the tcgen05 GEMM at the engine's small-M shapes (b = 1 prefill and one-image tower).
Weights rotate over enough copies to exceed L2, as in a forward where every layer has its own.

  python scripts/sweep_gemm_tiles.py [--large] > gpurun_out/sweep_gemm_tiles.txt
"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1] / "vla-from-fastvlm_b200"))
import torch  # noqa: E402

from vla_fastvlm import _native as N  # noqa: E402

L2 = 160e6


def run(M, Nn, K, mode, bn, split=0, iters=48):
    if M > 4096:
        iters = 8
    f16 = mode in ("res16",)
    dt = torch.float16 if f16 else torch.bfloat16
    copies = max(2, min(64, int(L2 / (Nn * K * 2)) + 1))
    ws = [(torch.randn(Nn, K, device="cuda") / K ** 0.5).to(dt) for _ in range(copies)]
    a = torch.randn(M, K, device="cuda").to(dt)
    kw = dict(block_n=bn)
    if mode == "swiglu":
        out = torch.empty(M, Nn // 2, device="cuda", dtype=torch.bfloat16); kw.update(swiglu=True)
    elif mode == "res32":
        out = torch.zeros(M, Nn, device="cuda"); kw.update(resid=out, split_k=split)
    elif mode == "gelu16":
        out = torch.empty(M, Nn, device="cuda", dtype=torch.float16); kw.update(act=5, bias=torch.zeros(Nn, device="cuda"))
    elif mode in ("res", "res16"):
        out = torch.zeros(M, Nn, device="cuda", dtype=torch.bfloat16); kw.update(resid=out, bias=torch.zeros(Nn, device="cuda"))
    elif mode == "gelu":
        out = torch.empty(M, Nn, device="cuda", dtype=torch.bfloat16); kw.update(act=1, bias=torch.zeros(Nn, device="cuda"))
    else:
        out = torch.empty(M, Nn, device="cuda", dtype=torch.bfloat16)
    for i in range(4):
        N.op_gemm(a, ws[i % copies], out=out, **kw)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(iters):
            N.op_gemm(a, ws[i % copies], out=out, **kw)
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (5 * iters) * 1e3


SHAPES = [
    ("llm.qkv", 272, 1152, 896, "", [64, 128, 192, 256], [0]),
    ("llm.o", 272, 896, 896, "res32", [64, 128, 256], [1, 2, 4]),
    ("llm.gate_up", 272, 9728, 896, "swiglu", [128, 256], [0]),
    ("llm.down", 272, 896, 4864, "res32", [64, 128, 256], [1, 2, 4, 8]),
    ("vis.s2.fc1", 4096, 1536, 384, "gelu16", [64, 128, 192, 256], [0]),
    ("vis.s2.fc2", 4096, 384, 1536, "res16", [64, 128, 192], [0]),
    ("vis.s3.fc1", 1024, 3072, 768, "gelu16", [64, 128, 192, 256], [0]),
    ("vis.s3.fc2", 1024, 768, 3072, "res16", [64, 128, 192, 256], [0]),
    ("vis.s3.qkv", 1024, 2304, 768, "", [64, 128, 192, 256], [0]),
    ("vis.s3.proj", 1024, 768, 768, "res", [64, 128, 192, 256], [0]),
    ("vis.s4.fc1", 256, 6144, 1536, "gelu16", [64, 128, 192, 256], [0]),
    ("vis.s4.fc2", 256, 1536, 6144, "res16", [64, 128, 192, 256], [0]),
    ("vis.s4.qkv", 256, 4608, 1536, "", [64, 128, 192, 256], [0]),
    ("vis.s4.proj", 256, 1536, 1536, "res", [64, 128, 192, 256], [0]),
    ("proj.0", 256, 896, 3072, "gelu", [64, 128, 192, 256], [0]),
]
LARGE = [
    ("llm.qkv", 17408, 1152, 896, "", [128, 192, 256], [0]),
    ("llm.o", 17408, 896, 896, "res32", [128, 192, 256], [0]),
    ("llm.gate_up", 17408, 9728, 896, "swiglu", [128, 256], [0]),
    ("llm.down", 17408, 896, 4864, "res32", [128, 192, 256], [0]),
    ("vis.s2.fc1", 131072, 1536, 384, "gelu16", [128, 192, 256], [0]),
    ("vis.s2.fc2", 131072, 384, 1536, "res16", [128, 192], [0]),
    ("vis.s3.fc1", 32768, 3072, 768, "gelu16", [128, 192, 256], [0]),
    ("vis.s3.fc2", 32768, 768, 3072, "res16", [128, 192, 256], [0]),
    ("vis.s3.qkv", 32768, 2304, 768, "", [128, 192, 256], [0]),
    ("vis.s3.proj", 32768, 768, 768, "res", [128, 192, 256], [0]),
    ("vis.s4.fc1", 8192, 6144, 1536, "gelu16", [128, 192, 256], [0]),
    ("vis.s4.fc2", 8192, 1536, 6144, "res16", [128, 192, 256], [0]),
    ("vis.s4.qkv", 8192, 4608, 1536, "", [128, 192, 256], [0]),
    ("vis.s4.proj", 8192, 1536, 1536, "res", [128, 192, 256], [0]),
    ("proj.0", 16384, 896, 3072, "gelu", [128, 192, 256], [0]),
    ("pe.pw", 524288, 192, 192, "gelu", [64, 128, 192], [0]),
]
if "--large" in sys.argv:   # the batch-64 shapes (vision chunks of 32 images)
    SHAPES = LARGE
for name, M, Nn, K, mode, bns, splits in SHAPES:
    auto = run(M, Nn, K, mode, 0, 0)
    res = []
    for bn in bns:
        for sp in splits:
            try:
                res.append((run(M, Nn, K, mode, bn, sp), bn, sp))
            except Exception as e:  # unsupported tile for this epilogue
                res.append((float("inf"), bn, sp))
    res.sort()
    print(f"{name:12s} M{M} N{Nn} K{K} {mode:7s} auto {auto:6.2f} us | " +
          "  ".join(f"bn{bn}/s{sp}:{t:6.2f}" for t, bn, sp in res[:6]), flush=True)
