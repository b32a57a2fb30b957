#!/usr/bin/env python
"""Batch sweep through the LeRobot plugin: ALOHA-shaped multi-camera batch dicts.

    python scripts/bench_sweep.py --model fastvlm-0.5b --batches 1,2,4,8,16,32,64,128,256,512

The batch dict carries three VISUAL keys (B,3,480,640), a 14-dim state and a task string per sample, on the device,
as LeRobot's pre-processor hands it to the policy.  Parity mode: like the reference, the policy consumes the FIRST
camera only (lerobot_fastvla/modeling_fastvla.py:53-67).  Every call is a full `select_action` (queue of length 1,
so each call refills it): letterbox 480x640 -> 768x1024 + 256 padded rows, encoder, projector, prefill, head.
Timing: CUDA events around `--iters` back-to-back calls after warm-up; one JSON line per batch size.  Without a real
`lerobot` install the minimal stand-in under tests/fake_lerobot provides the base classes (as in the tests).
"""
from __future__ import annotations

import argparse
import contextlib
import json
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "vla-from-fastvlm_b200"))
sys.path.insert(0, str(ROOT))
try:
    import lerobot  # noqa: F401
except ImportError:
    sys.path.insert(0, str(ROOT / "tests" / "fake_lerobot"))

import torch  # noqa: E402

STATE_DIM = ACTION_DIM = 14
IMG_HW = (480, 640)
CAMS = ("observation.images.cam_high", "observation.images.cam_left_wrist", "observation.images.cam_right_wrist")


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="fastvlm-0.5b")
    ap.add_argument("--batches", default="1,2,4,8,16,32,64,128,256,512")
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--dtype", default="bfloat16", choices=["bfloat16", "float32"],
                    help="float32 = the plugin's default parity mode (SIMT fp32 kernels, ~30x slower)")
    ap.add_argument("--warmup", type=int, default=3)
    args = ap.parse_args()
    if not torch.cuda.is_available():
        raise SystemExit("bench_sweep.py needs a CUDA device: the FastVLA B200 path has no CPU fallback")
    from lerobot.configs.types import FeatureType, PolicyFeature

    from vla_fastvlm.lerobot_fastvla import FastVLAConfig, FastVLAPolicy

    inp = {k: PolicyFeature(FeatureType.VISUAL, (3, *IMG_HW)) for k in CAMS}
    inp["observation.state"] = PolicyFeature(FeatureType.STATE, (STATE_DIM,))
    outp = {"action": PolicyFeature(FeatureType.ACTION, (ACTION_DIM,))}
    cfg = FastVLAConfig(input_features=inp, output_features=outp, device="cuda", vlm_model_name=f"synthetic:{args.model}",
                        image_token_mode="prefix", chunk_size=50, n_action_steps=1,
                        compute_dtype=args.dtype)
    with contextlib.redirect_stdout(sys.stderr):
        policy = FastVLAPolicy(cfg).cuda().eval()
    dev = torch.device("cuda", 0)
    g = torch.Generator().manual_seed(5)
    for b in [int(x) for x in args.batches.split(",")]:
        batch = {k: torch.rand(b, 3, *IMG_HW, generator=g).to(dev) for k in CAMS}
        batch["observation.state"] = torch.randn(b, STATE_DIM, generator=g).to(dev)
        batch["task"] = ["insert the peg into the socket"] * b
        for _ in range(args.warmup):
            out = policy.select_action(batch)
        torch.cuda.synchronize()
        assert out.shape == (b, ACTION_DIM) and torch.isfinite(out).all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(args.iters):
            policy.select_action(batch)
        e1.record()
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) * 1e3 / args.iters
        ms = e0.elapsed_time(e1) / args.iters
        print(json.dumps({"metric": "obs->action chunks/sec", "batch": b, "ms_per_call": ms, "wall_ms_per_call": wall,
                          "value": b / (ms / 1e3), "unit": "samples/s", "model": args.model, "cameras_in_batch": len(CAMS),
                          "cameras_used": 1, "image": list(IMG_HW), "dtype": args.dtype}), flush=True)
        del batch
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
