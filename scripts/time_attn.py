"""Time the attention kernels at the FastViTHD MHSA shapes (bf16, head_dim 32, non-causal) and check them against an
fp32 softmax reference: python scripts/time_attn.py   (FVLA_DISABLE_TC_ATTN=1 selects the mma.sync kernel)."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1] / "vla-from-fastvlm_b200"))
import torch  # noqa: E402

from vla_fastvlm import _native as N  # noqa: E402

for B, Ntok, heads in [(32, 1024, 24), (32, 256, 48)]:
    hd = 32
    g = torch.Generator(device="cuda").manual_seed(0)
    qkv = (torch.randn(B * Ntok, 3 * heads * hd, device="cuda", generator=g) * 1.5).bfloat16()
    for _ in range(3):
        out = N.op_attention(qkv, B, Ntok, heads, heads, hd, hd ** -0.5, False)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        out = N.op_attention(qkv, B, Ntok, heads, heads, hd, hd ** -0.5, False)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    x = qkv[: 2 * Ntok].float().view(2, Ntok, 3, heads, hd).permute(2, 0, 3, 1, 4)
    ref = torch.softmax((x[0] * hd ** -0.5) @ x[1].transpose(-1, -2), -1) @ x[2]
    ref = ref.permute(0, 2, 1, 3).reshape(2 * Ntok, heads * hd)
    err = float((out[: 2 * Ntok].float() - ref).abs().max() / ref.abs().max())
    fl = 4.0 * B * heads * Ntok * Ntok * hd
    print(f"B={B} N={Ntok} heads={heads}: {ms:.4f} ms  {fl / ms / 1e9:.1f} TFLOP/s  "
          f"{B * heads * Ntok * Ntok / ms / 1e6:.1f} Gscores/s  rel err {err:.2e}")

# Qwen2-0.5B prefill of the batch-64 step: causal, 14 query heads on 2 kv heads, head_dim 64, T' = 272
B, T, hq, hkv, hd = 64, 272, 14, 2, 64
g = torch.Generator(device="cuda").manual_seed(1)
qkv = torch.randn(B * T, (hq + 2 * hkv) * hd, device="cuda", generator=g).bfloat16()
for _ in range(3):
    out = N.op_attention(qkv, B, T, hq, hkv, hd, hd ** -0.5, True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    out = N.op_attention(qkv, B, T, hq, hkv, hd, hd ** -0.5, True)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
x = qkv[:T].float()
q = x[:, : hq * hd].view(T, hq, hd).permute(1, 0, 2)
k = x[:, hq * hd: (hq + hkv) * hd].view(T, hkv, hd).permute(1, 0, 2).repeat_interleave(hq // hkv, 0)
v = x[:, (hq + hkv) * hd:].view(T, hkv, hd).permute(1, 0, 2).repeat_interleave(hq // hkv, 0)
sc = (q * hd ** -0.5) @ k.transpose(-1, -2) + torch.full((T, T), float("-inf"), device="cuda").triu(1)
ref = (torch.softmax(sc, -1) @ v).permute(1, 0, 2).reshape(T, hq * hd)
err = float((out[:T].float() - ref).abs().max() / ref.abs().max())
print(f"causal GQA B={B} T={T}: {ms:.4f} ms  {B * hq * T * T / 2 / ms / 1e6:.1f} Gscores/s  rel err {err:.2e}")
