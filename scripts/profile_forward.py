"""Per-kernel time/roofline table of one FastVLA forward on the current GPU (CUDA events per launch).

  python scripts/profile_forward.py --model fastvlm-0.5b --batch 64 --steps 3
"""
from __future__ import annotations

import argparse
import json
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "vla-from-fastvlm_b200"))

import torch  # noqa: E402

from vla_fastvlm.model.arch import PRESETS  # noqa: E402
from vla_fastvlm.model.engine import BACKBONE_KEY_PREFIX, NativeEngine  # noqa: E402
from vla_fastvlm.model.synthetic import synthetic_backbone_state_dict, synthetic_head_state_dict  # noqa: E402


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="fastvlm-0.5b")
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--dtype", default="bf16")
    ap.add_argument("--hw", type=int, nargs=2, default=[480, 480])
    ap.add_argument("--text", type=int, default=16)
    ap.add_argument("--chunk", type=int, default=0)
    ap.add_argument("--state-dim", type=int, default=4)
    ap.add_argument("--action-dim", type=int, default=4)
    ap.add_argument("--json", default="")
    args = ap.parse_args()

    arch = PRESETS[args.model]
    dtype = torch.bfloat16 if args.dtype == "bf16" else torch.float32
    t0 = time.time()
    sd = synthetic_backbone_state_dict(arch, 0)
    hsd = synthetic_head_state_dict(arch.text.hidden, args.state_dim, args.action_dim, 1024, 1024, 1)
    eng = NativeEngine(arch, dtype=dtype, state_dim=args.state_dim, action_dim=args.action_dim,
                       vision_chunk=args.chunk)
    eng.load_state_dict(sd, prefix=BACKBONE_KEY_PREFIX)
    eng.load_state_dict(hsd)
    del sd
    eng.finalize()
    print(f"# weights ready in {time.time() - t0:.1f}s, {eng.weight_bytes / 1e9:.2f} GB on device", flush=True)

    B, T = args.batch, args.text + 1
    g = torch.Generator().manual_seed(1)
    images = torch.rand(B, 3, args.hw[0], args.hw[1], generator=g).cuda()
    states = torch.randn(B, args.state_dim, generator=g).cuda()
    ids = torch.randint(0, arch.text.vocab, (B, T), generator=g)
    ids[:, 0] = -200
    lens = torch.full((B,), T)
    for _ in range(args.warmup):
        out = eng.forward(images, ids, lens, states=states)
    torch.cuda.synchronize()
    print(f"# workspace {eng.workspace_bytes / 1e9:.2f} GB, launches/forward {eng.last_launch_count}, "
          f"T'={eng.merged_len}, algorithmic {eng.last_forward_flops / 1e12:.2f} TFLOP/forward", flush=True)

    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        out = eng.forward(images, ids, lens, states=states)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    tf = eng.last_forward_flops / ms / 1e9
    print(f"# forward {ms:.2f} ms/step  {B / ms * 1e3:.1f} samples/s  {tf:.1f} TFLOP/s algorithmic", flush=True)
    assert torch.isfinite(out).all()

    eng.set_profile(True)
    for _ in range(args.steps):
        eng.forward(images, ids, lens, states=states)
    rows = eng.profile_report()
    eng.set_profile(False)
    tot = sum(r["total_ms"] for r in rows)
    rows.sort(key=lambda r: -r["total_ms"])
    print(f"# profiled sum {tot / args.steps:.2f} ms/step over {len(rows)} distinct kernels/shapes")
    print(f"{'label':58s} {'n':>5s} {'ms/step':>9s} {'%':>6s} {'TFLOP/s':>9s} {'GB/s':>9s}")
    for r in rows:
        ms_step = r["total_ms"] / args.steps
        tfs = r["flops"] / r["total_ms"] / 1e9 if r["total_ms"] > 0 else 0
        gbs = r["bytes"] / r["total_ms"] / 1e6 if r["total_ms"] > 0 else 0
        print(f"{r['label']:58s} {r['count'] // args.steps:5d} {ms_step:9.3f} {100 * r['total_ms'] / tot:6.2f} "
              f"{tfs:9.1f} {gbs:9.1f}")
    if args.json:
        with open(args.json, "w") as f:
            json.dump(dict(ms_per_step=ms, batch=B, rows=rows, steps=args.steps), f, indent=1)


if __name__ == "__main__":
    main()
