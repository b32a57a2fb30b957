#!/usr/bin/env python
"""Headline benchmark: FastVLA-0.5B batched policy forward (obs images + prompt + state -> action).

    python bench.py --gpus 1 --steps 20 --warmup 5
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus 8 --steps 20 --warmup 5
    python bench.py --impl reference --steps 3 --warmup 1      # the reference's CPU path (oracle port)

One "step" = one pass of the hot path over one batch of synthetic MetaWorld-MT50-shaped observations
(BASELINE.json configs[1]): 64 frames 480x480 float [0,1], 4-dim state, one prompt per sample from a
pool of 50 (8-16 tokens) with the image placeholder in front (T' = 256 + T_text), bf16, random-init
FastVLM-0.5B + FastVLA head.  Prints ONE JSON line (see the task contract); rank 0 only.
"""
from __future__ import annotations

import argparse
import contextlib
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT / "vla-from-fastvlm_b200"))
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402

METRIC = "obs->action chunks/sec"
UNIT = "samples/s"
STATE_DIM = ACTION_DIM = 4          # MetaWorld MT50 (LeRobot): 4-dim state, 4-dim action
IMG_HW = (480, 480)


# ------------------------------------------------------------------------------------------------
# workload
# ------------------------------------------------------------------------------------------------
def make_prompt_pool(n: int = 50, seed: int = 3):
    """50 synthetic prompts whose byte-tokenised length is 8..16 (ragged -> right padding + per-sample pool_idx)."""
    g = torch.Generator().manual_seed(seed)
    words = ["pick", "push", "open", "close", "reach", "press", "pull", "slide", "turn", "place", "door", "drawer",
             "button", "peg", "cup", "box", "lever", "block", "red", "blue"]
    pool = []
    while len(pool) < n:
        k = int(torch.randint(1, 4, (1,), generator=g))
        s = " ".join(words[int(torch.randint(0, len(words), (1,), generator=g))] for _ in range(k))
        s = s[: int(torch.randint(7, 16, (1,), generator=g))]
        if 7 <= len(s) <= 15 and s not in pool:  # +1 for the trailing newline => 8..16 tokens
            pool.append(s)
    return pool


def make_batch(batch: int, seed: int):
    g = torch.Generator().manual_seed(seed)
    images = torch.rand(batch, 3, IMG_HW[0], IMG_HW[1], generator=g)
    states = torch.randn(batch, STATE_DIM, generator=g)
    pool = make_prompt_pool()
    tasks = [pool[int(torch.randint(0, len(pool), (1,), generator=g))] for _ in range(batch)]
    return images, states, tasks


def shard_range(total: int, rank: int, world: int):
    """Contiguous split of `total` independent samples over ranks (strong-scaling helper; host-side only)."""
    base, rem = divmod(total, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def max_over_ranks(value: float, device) -> float:
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int) -> None:
        self.index = index
        self.samples = []
        self._stop = threading.Event()
        self._thread = threading.Thread(target=self._run, daemon=True)

    # NVML throttle-reason bits (nvml.h: nvmlClocksEventReason*)
    _BITS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4))

    def _run_nvml(self) -> bool:
        """Sample through NVML every 20 ms (a 10-step timed region lasts < 1 s; nvidia-smi manages 2-3 samples)."""
        try:
            import pynvml

            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            mx = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
            reasons_fn = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
                pynvml.nvmlDeviceGetCurrentClocksThrottleReasons
            reasons_fn(h)
        except Exception:
            return False
        while not self._stop.is_set():
            try:
                sm = pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
                mask = int(reasons_fn(h))
                pw = pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0
                self.samples.append([str(sm), str(mx), f"{pw:.1f}"] +
                                    ["Active" if mask & bit else "Not Active" for _, bit in self._BITS])
            except Exception:
                pass
            self._stop.wait(0.02)
        return True

    def _run(self) -> None:
        if self._run_nvml():
            return
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.FIELDS}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                parts = [p.strip() for p in out.strip().split(",")]
                if len(parts) >= 7:
                    self.samples.append(parts)
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self._thread.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        self._thread.join(timeout=6)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = [float(s[0]) for s in self.samples if s[0].replace(".", "").isdigit()]
        mx = [float(s[1]) for s in self.samples if s[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[3 + i].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------
# CPU baseline (the reference's own CPU path = oracle port; the remote-code VLM cannot be installed offline)
# ------------------------------------------------------------------------------------------------
def cpu_reference_run(model: str, steps: int, warmup: int, seed: int = 0):
    """Times the fp32 CPU oracle on ONE sample per step (bounded sample of the batch-64 workload)."""
    from oracle.fastvla_oracle import IMAGE_TOKEN_INDEX, FastVLAOracle
    from vla_fastvlm.model.arch import PRESETS
    from vla_fastvlm.model.synthetic import synthetic_backbone_state_dict, synthetic_head_state_dict
    from vla_fastvlm.model.tokenizer import SimpleByteTokenizer

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    arch = PRESETS[model]
    sd = synthetic_backbone_state_dict(arch, seed)
    hsd = synthetic_head_state_dict(arch.text.hidden, STATE_DIM, ACTION_DIM, 1024, 1024, seed + 1)
    oracle = FastVLAOracle(arch, sd, hsd)
    images, states, tasks = make_batch(max(1, steps + warmup), seed=7)
    tok = SimpleByteTokenizer(arch.text.vocab)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        enc = tok([tasks[i] + "\n"], padding="longest", truncation=True, max_length=64)
        ids = torch.cat([torch.full((1, 1), IMAGE_TOKEN_INDEX, dtype=torch.long), enc["input_ids"]], 1)
        mask = torch.cat([torch.ones(1, 1, dtype=torch.long), enc["attention_mask"]], 1)
        out = oracle.forward(images[i:i + 1], states[i:i + 1], ids, mask)
        assert torch.isfinite(out).all()
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    per = statistics.mean(times)
    return {"value": 1.0 / per, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{steps} single-sample fp32 forwards of the same workload (1 x 480x480 frame + prompt + state), "
                      f"{torch.get_num_threads()} torch threads, mean {per:.2f} s/sample"}, per


# ------------------------------------------------------------------------------------------------
def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--model", default="fastvlm-0.5b")
    ap.add_argument("--batch", type=int, default=64, help="per-GPU batch (weak scaling)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-samples", type=int, default=16, help="oracle forwards timed for cpu_baseline (~0.7 s each)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    config = {"workload": f"FastVLA-{args.model.split('-')[-1]} bf16 batched select_action, batch {args.batch}/GPU synthetic "
                          "MetaWorld-MT50-shaped obs (480x480 frame -> 1024^2, 4-dim state, 8-16 token prompt, image tokens "
                          "prefixed: T'=256+T_text)",
              "per_gpu_batch": args.batch, "global_batch": args.batch * world, "image": list(IMG_HW),
              "model": args.model, "parallelism": f"dp{world} (batch sharded, no collective)",
              "l2_policy": "inputs+weights per step (177 MB images + 1.25 GB weights) exceed the 126 MB L2"}

    # ------------------------------------------------------------ reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return
        base, per = cpu_reference_run(args.model, max(1, args.steps), max(0, args.warmup))
        line = {"metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": per * 1e3, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config, "impl": "reference",
                "cpu_baseline": base,
                "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line), flush=True)
        return

    # ------------------------------------------------------------ B200 arm
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device: the FastVLA B200 path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=dev)

    from vla_fastvlm.fastvla import FastVLAConfig, FastVLAPolicy

    cfg = FastVLAConfig(vlm_model_name=f"synthetic:{args.model}", state_dim=STATE_DIM, action_dim=ACTION_DIM,
                        compute_dtype="bfloat16", image_token_mode="prefix")
    with contextlib.redirect_stdout(sys.stderr):  # the adapter logs its image size like the reference; stdout = the JSON line only
        policy = FastVLAPolicy(cfg).to(dev).eval()
    engine = policy.model.backbone.model.engine  # builds + uploads weights
    images_h, states_h, tasks = make_batch(args.batch, seed=100 + rank)
    tasks_n = policy.processor.prepare_tasks(tasks, batch_size=args.batch)
    images_pin, states_pin = images_h.pin_memory(), states_h.pin_memory()
    images_d, states_d = images_pin.to(dev), states_pin.to(dev)
    ids, lens, pool_idx = policy.model.backbone._prompt_ids(tasks_n)

    def step_device():
        return engine.forward(images_d, ids, lens, states=states_d, pool_idx=pool_idx)

    # End to end: every step copies THAT step's observations from pinned host memory and reads its actions back.
    # The copy of step i+1 is issued on a side stream while step i computes (double-buffered device staging), the
    # way a serving loop overlaps ingest with inference; the first copy of a run is exposed.
    copy_stream = torch.cuda.Stream(device=dev)
    stage_img = [torch.empty_like(images_d) for _ in range(2)]
    stage_st = [torch.empty_like(states_d) for _ in range(2)]
    copied = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]

    def h2d(slot):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[slot])  # the forward that last read this slot has finished
            stage_img[slot].copy_(images_pin, non_blocking=True)
            stage_st[slot].copy_(states_pin, non_blocking=True)
            copied[slot].record(copy_stream)

    out_pin = [torch.empty((args.batch, ACTION_DIM), dtype=torch.float32).pin_memory() for _ in range(2)]
    out_ready = [torch.cuda.Event() for _ in range(2)]

    def run_e2e(n):
        """n steps; the actions of step i are read back (pinned, async) and consumed on the host while step i+1 is
        already queued, so neither direction of PCIe nor the host read-back stalls the GPU between steps."""
        res = None
        cur = torch.cuda.current_stream(dev)
        for c in consumed:
            c.record(cur)
        h2d(0)
        for i in range(n):
            slot = i & 1
            if i + 1 < n:
                h2d(slot ^ 1)
            cur.wait_event(copied[slot])
            act = policy.forward(stage_img[slot], stage_st[slot], tasks, device=dev)  # public API, one engine call
            consumed[slot].record(cur)
            out_pin[slot].copy_(act, non_blocking=True)  # device -> host read of this step's result
            out_ready[slot].record(cur)
            if i > 0:
                out_ready[slot ^ 1].synchronize()
                res = out_pin[slot ^ 1].clone()
        out_ready[(n - 1) & 1].synchronize()
        res = out_pin[(n - 1) & 1].clone()
        return res

    def barrier():
        if world > 1:
            import torch.distributed as dist

            dist.barrier()
        torch.cuda.synchronize()

    with torch.no_grad():
        for _ in range(max(3, args.warmup)):
            out = step_device()
        torch.cuda.synchronize()
        assert torch.isfinite(out).all()
        launches_per_step = engine.last_launch_count
        flops_per_step = engine.last_forward_flops

        # ---- timed region 1: inputs resident in HBM ----
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        with ClockSampler(local_rank) as clk:
            e0.record()
            for _ in range(args.steps):
                step_device()
            e1.record()
            barrier()
        ms = max_over_ranks(e0.elapsed_time(e1), dev)
        clocks = clk.summary()

        # ---- timed region 2: end to end through the public API (pinned host inputs, result read back) ----
        run_e2e(2)
        barrier()
        t0 = time.perf_counter()
        res = run_e2e(args.steps)
        barrier()
        e2e_s = max_over_ranks(time.perf_counter() - t0, dev)
        h2d = images_pin.numel() * 4 + states_pin.numel() * 4 + ids.numel() * 4
        d2h = res.numel() * 4

        # ---- per-kernel roofline pass (CUDA events around every launch, same stream) ----
        roof = None
        breakdown = None
        if rank == 0:
            engine.set_profile(True)
            for _ in range(min(args.steps, 3)):
                step_device()
            rows = engine.profile_report()
            engine.set_profile(False)
            n_prof = min(args.steps, 3)
            tot = sum(r["total_ms"] for r in rows)
            fam = {}
            for r in rows:
                key = ("gemm_tcgen05" if ("gemm" in r["label"] or "ffn_fused" in r["label"]) else
                       "dwconv" if "dwconv" in r["label"] else
                       "attention" if "attention" in r["label"] else "other")
                f = fam.setdefault(key, dict(ms=0.0, flops=0.0, bytes=0.0, launches=0))
                f["ms"] += r["total_ms"]; f["flops"] += r["flops"]; f["bytes"] += r["bytes"]; f["launches"] += r["count"]
            peaks = {}
            pk = ROOT / "MEASURED_PEAKS.json"
            if pk.is_file():
                peaks = json.loads(pk.read_text())
            peak_tf = float(peaks.get("bf16_tflops_sustained", 1400.0))
            peak_src = "MEASURED_PEAKS.json bf16_tflops_sustained" if peaks else "fallback 1.4 PFLOP/s sustained"
            gm = fam.get("gemm_tcgen05", dict(ms=1.0, flops=0.0, launches=1))
            achieved = gm["flops"] / gm["ms"] / 1e9
            # DRAM bytes per launch of the same kernels from the committed ncu pass (profiles/, same command line);
            # null until that file exists
            traffic = None
            tj = ROOT / "profiles" / "r01_ncu_traffic.json"
            if tj.is_file():
                tr = json.loads(tj.read_text()).get("tcgen05_gemm_family", {})
                if tr.get("launches"):
                    traffic = tr["dram_bytes"] / tr["launches"]
            roof = {"bound": "tensor",
                    "kernel": "gemm_bf16_tcgen05_kernel + ffn_fused_kernel (all tcgen05 launches of a step)",
                    "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved / peak_tf,
                    "traffic": traffic, "algorithmic_bytes_per_launch": gm["bytes"] / max(1, gm["launches"]),
                    "peak_source": peak_src,
                    "avg_launch_ms": gm["ms"] / max(1, gm["launches"]), "share_of_step": gm["ms"] / tot}
            hbm = float(peaks.get("hbm_gbs", 6650.0))
            breakdown = {k: {"ms_per_step": v["ms"] / n_prof, "share": v["ms"] / tot,
                             "tflops": v["flops"] / v["ms"] / 1e9 if v["ms"] else 0.0,
                             "gbs": v["bytes"] / v["ms"] / 1e6 if v["ms"] else 0.0,
                             "hbm_frac": (v["bytes"] / v["ms"] / 1e6) / hbm if v["ms"] else 0.0}
                         for k, v in fam.items()}

        # ---- p50 latency of a single observation (the metric's second half) ----
        lat = None
        if rank == 0:
            one_i, one_s, one_t = images_pin[:1], states_pin[:1], tasks[:1]
            ts = []
            for i in range(25):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                policy.forward(one_i.to(dev, non_blocking=True), one_s.to(dev, non_blocking=True), one_t, device=dev).cpu()
                if i >= 5:
                    ts.append((time.perf_counter() - t0) * 1e3)
            lat = {"p50_ms": statistics.median(ts), "batch": 1, "path": "policy.forward, host in / host out"}

    total = args.batch * world * args.steps
    value = total / (ms / 1e3)
    cpu_base = None
    if rank == 0 and not args.no_cpu_baseline and world == 1:
        cpu_base, _ = cpu_reference_run(args.model, args.cpu_samples, 1)
    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(3, args.warmup), "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": config,
                "clocks": clocks,
                "e2e": {"value": total / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
                "gpu_launches": int(launches_per_step * args.steps),
                "algorithmic_tflop_per_step": flops_per_step / 1e12,
                "achieved_tflops_whole_step": flops_per_step * world / (ms / args.steps) / 1e9,
                "roofline": roof, "kernel_families": breakdown, "latency": lat, "cpu_baseline": cpu_base,
                "impl": "b200"}
        print(json.dumps(line), flush=True)
    if world > 1:
        import torch.distributed as dist

        dist.destroy_process_group()


if __name__ == "__main__":
    main()
