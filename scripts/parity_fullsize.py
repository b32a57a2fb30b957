"""Full-size parity: FastVLA-0.5B (random-init, seeded) engine output vs the CPU fp32 oracle on the
same inputs, with per-stage relative errors.  ~1 min on the GPU box (oracle ~1-2 s/sample on CPU)."""
import argparse
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "vla-from-fastvlm_b200"))
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402

from oracle.fastvla_oracle import IMAGE_TOKEN_INDEX, FastVLAOracle  # noqa: E402
from vla_fastvlm import _native as N  # noqa: E402
from vla_fastvlm.model.arch import PRESETS  # noqa: E402
from vla_fastvlm.model.engine import BACKBONE_KEY_PREFIX, NativeEngine  # noqa: E402
from vla_fastvlm.model.synthetic import synthetic_backbone_state_dict, synthetic_head_state_dict  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--model", default="fastvlm-0.5b")
ap.add_argument("--batch", type=int, default=2)
ap.add_argument("--dtype", default="bf16")
ap.add_argument("--hw", type=int, nargs=2, default=[480, 480])
ap.add_argument("--bf16-weights", action="store_true",
                help="round every weight matrix / conv kernel to bf16 first (a bf16-stored checkpoint): isolates "
                     "arithmetic error from weight quantisation")
a = ap.parse_args()
arch = PRESETS[a.model]
dtype = torch.bfloat16 if a.dtype == "bf16" else torch.float32
S_DIM = A_DIM = 14
sd = synthetic_backbone_state_dict(arch, 0)
hsd = synthetic_head_state_dict(arch.text.hidden, S_DIM, A_DIM, 1024, 1024, 1)
if a.bf16_weights:
    sd = {k: (v.bfloat16().float() if v.is_floating_point() and v.ndim >= 2 else v) for k, v in sd.items()}
g = torch.Generator().manual_seed(1)
B, T = a.batch, 17
images = torch.rand(B, 3, a.hw[0], a.hw[1], generator=g)
states = torch.randn(B, S_DIM, generator=g)
ids = torch.randint(0, 151643, (B, T), generator=g)
ids[:, 0] = IMAGE_TOKEN_INDEX
mask = torch.ones(B, T, dtype=torch.long)
if B > 1:
    mask[1, 11:] = 0
torch.set_num_threads(max(1, torch.get_num_threads()))
oracle = FastVLAOracle(arch, sd, hsd)
taps = {}
ref = oracle.forward(images, states, ids, mask, taps=taps)

eng = NativeEngine(arch, dtype=dtype, state_dim=S_DIM, action_dim=A_DIM)
eng.load_state_dict(sd, prefix=BACKBONE_KEY_PREFIX)
eng.load_state_dict(hsd)
eng.finalize()
dev = eng.device
v = arch.vision
bufs = {}


def reg(stage, shape, dt=dtype):
    bufs[stage] = torch.zeros(shape, device=dev, dtype=dt)
    eng.set_tap(stage, bufs[stage])


side = v.image_size // 4
reg(N.TAP_STEM, (B, side, side, v.dims[0]))
for i, d in enumerate(v.dims):
    reg(N.TAP_VIS_STAGE0 + i, (B, side, side, d))
    side //= 2
reg(N.TAP_IMAGE_FEATURES, (B, v.num_tokens, v.out_channels))
reg(N.TAP_PROJECTOR, (B, v.num_tokens, arch.text.hidden))
Tm = taps["embeds"].shape[1]
for l in (0, arch.text.layers // 2, arch.text.layers - 1):
    reg(N.TAP_LAYER0 + l, (B, Tm, arch.text.hidden))
reg(N.TAP_POOLED, (B, arch.text.hidden), torch.float32)
out = eng.forward(images.to(dev), ids, mask.sum(1), states=states.to(dev)).float().cpu()


def rel(x, y):
    x, y = x.float().cpu(), y.float()
    return float((x - y).abs().max() / (y.abs().max() + 1e-12)), float((x - y).norm() / (y.norm() + 1e-12))


print(f"# {a.model} {a.dtype} B={B}: stage, max-abs/max, ||err||/||ref||")
print("stem", rel(bufs[N.TAP_STEM], taps["stem"].permute(0, 2, 3, 1)))
for i in range(len(v.dims)):
    print(f"vis_stage{i}", rel(bufs[N.TAP_VIS_STAGE0 + i], taps[f"vis_stage{i}"].permute(0, 2, 3, 1)))
print("image_features", rel(bufs[N.TAP_IMAGE_FEATURES], taps["image_features"]))
print("projector", rel(bufs[N.TAP_PROJECTOR], taps["projector"]))
valid = torch.zeros(B, Tm, 1)
for b in range(B):
    valid[b, : int(mask[b].sum()) + v.num_tokens - 1] = 1
for l in (0, arch.text.layers // 2, arch.text.layers - 1):
    print(f"layer{l}", rel(bufs[N.TAP_LAYER0 + l].float().cpu() * valid, taps[f"layer{l}"] * valid))
print("pooled", rel(bufs[N.TAP_POOLED], taps["pooled"]))
print("actions", rel(out, ref), "max-abs", float((out - ref).abs().max()))
print("ref actions[0,:5]", ref[0, :5].tolist())
print("got actions[0,:5]", out[0, :5].tolist())
