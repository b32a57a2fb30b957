"""Does the ConvFFN hidden tensor stay in L2 when fc1 -> fc2 run on small row chunks with a reused hidden buffer?
python scripts/l2_chain.py"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1] / "vla-from-fastvlm_b200"))
import torch  # noqa: E402

from vla_fastvlm import _native as N  # noqa: E402

C, Hd, Mtot = 384, 1536, 131072
x = torch.randn(Mtot, C, device="cuda").bfloat16()
res = torch.randn(Mtot, C, device="cuda").bfloat16()
w1 = (torch.randn(Hd, C, device="cuda") / C ** 0.5).bfloat16()
b1 = torch.randn(Hd, device="cuda")
w2 = (torch.randn(C, Hd, device="cuda") / Hd ** 0.5).half()
b2 = torch.randn(C, device="cuda")
out = torch.empty_like(x)


def run(chunk):
    hid = torch.empty(chunk, Hd, device="cuda", dtype=torch.float16)
    def step():
        for m0 in range(0, Mtot, chunk):
            rows = min(chunk, Mtot - m0)
            N.op_gemm(x[m0:m0 + rows], w1, bias=b1, act=5, out=hid[:rows])
            N.op_gemm(hid[:rows], w2, bias=b2, resid=res[m0:m0 + rows], out=out[m0:m0 + rows])
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    # replay from a CUDA graph: the python/ctypes launch cost (~15 us per op) would otherwise dominate the small slices
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        step()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr, stream=st):
            step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3):
        gr.replay()
    e0.record()
    for _ in range(10):
        gr.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 10


for chunk in (131072, 65536, 37888, 32768, 18944, 16384, 9472, 8192):
    print(f"chunk {chunk:7d} rows (hidden {chunk * Hd * 2 / 1e6:6.1f} MB): {run(chunk):.3f} ms per fc1+fc2 over {Mtot} rows")
