"""One attention launch (for ncu / timing): python scripts/one_attn.py B N heads_q heads_kv head_dim causal"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1] / "vla-from-fastvlm_b200"))
import torch  # noqa: E402

from vla_fastvlm import _native as N  # noqa: E402

B, n, hq, hkv, hd, causal = (int(x) for x in sys.argv[1:7])
qkv = torch.randn(B * n, (hq + 2 * hkv) * hd, device="cuda").bfloat16()
for _ in range(3):
    out = N.op_attention(qkv, B, n, hq, hkv, hd, hd ** -0.5, bool(causal))
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    out = N.op_attention(qkv, B, n, hq, hkv, hd, hd ** -0.5, bool(causal))
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
fl = 4.0 * B * hq * n * n * hd * (0.5 if causal else 1.0)
print(f"attn B{B} N{n} hq{hq} hkv{hkv} hd{hd} causal{causal}: {ms:.3f} ms {fl / ms / 1e9:.1f} TFLOP/s")
