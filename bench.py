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


# ------------------------------------------------------------------------------------------------
# LeRobot plugin set-up (the metric's entry point is `FastVLAPolicy.select_action(batch)`)
# ------------------------------------------------------------------------------------------------
def _import_lerobot_plugin():
    """`lerobot` (0.4.x) is not installed in this image and cannot be fetched; the plugin subclasses only its base
    classes (PreTrainedPolicy / PreTrainedConfig / processor steps), which tests/fake_lerobot provides with the same
    module paths.  A real installation, when importable, wins."""
    try:
        import lerobot  # noqa: F401
    except ImportError:
        sys.path.insert(0, str(ROOT / "tests" / "fake_lerobot"))
    from vla_fastvlm import lerobot_fastvla

    return lerobot_fastvla


def dataset_stats():
    """Synthetic MEAN_STD statistics for the state and action features (what `make_fastvla_pre_post_processors`
    receives from a LeRobot dataset)."""
    g = torch.Generator().manual_seed(11)
    return {"observation.state": {"mean": torch.randn(STATE_DIM, generator=g) * 0.2,
                                  "std": torch.rand(STATE_DIM, generator=g) * 0.5 + 0.75},
            "action": {"mean": torch.randn(ACTION_DIM, generator=g) * 0.2,
                       "std": torch.rand(ACTION_DIM, generator=g) * 0.5 + 0.75}}


def make_lerobot_policy(model: str, dev, **overrides):
    plug = _import_lerobot_plugin()
    from lerobot.configs.types import FeatureType, PolicyFeature

    img_key, state_key = "observation.image", "observation.state"
    inp = {img_key: PolicyFeature(FeatureType.VISUAL, (3, IMG_HW[0], IMG_HW[1])),
           state_key: PolicyFeature(FeatureType.STATE, (STATE_DIM,))}
    outp = {"action": PolicyFeature(FeatureType.ACTION, (ACTION_DIM,))}
    kw = dict(input_features=inp, output_features=outp, device=str(dev), vlm_model_name=f"synthetic:{model}",
              compute_dtype="bfloat16", image_token_mode="prefix", fuse_io_normalization=True,
              image_input_scale=1.0 / 255.0)
    kw.update(overrides)
    cfg = plug.FastVLAConfig(**kw)
    with contextlib.redirect_stdout(sys.stderr):  # the adapter logs its image size like the reference; stdout = JSON only
        policy = plug.FastVLAPolicy(cfg).to(dev).eval()
    pre, post = plug.make_fastvla_pre_post_processors(cfg, dataset_stats=dataset_stats())
    return policy, pre, post, (img_key, state_key)


def make_camera_batch(batch: int, seed: int):
    """The same observations as `make_batch`, as a camera delivers them: uint8 HWC frames (the [0,1] floats of
    make_batch quantised to 8 bits) and the raw state, in pinned host memory."""
    images, states, tasks = make_batch(batch, seed)
    frames = (images.permute(0, 2, 3, 1) * 255.0).round().clamp(0, 255).to(torch.uint8).contiguous()
    pin = torch.cuda.is_available()
    return (frames.pin_memory() if pin else frames), (states.pin_memory() if pin else states), tasks


def default_config_throughput(model: str, dev, frames, states, tasks):
    """samples/s of the plugin's DEFAULT configuration — `compute_dtype='float32'` (the reference loads fp32) and the
    reference-literal `image_token_mode='none'`, where the prompt carries no image placeholder, the tower's output
    cannot reach the action and the engine skips it — plus fp32 with the image tokens spliced in."""
    out = {}
    img = frames.permute(0, 3, 1, 2).float().div(255.0).to(dev)
    st = states.to(dev)
    for name, kw in (("fp32_none", dict(compute_dtype="float32", image_token_mode="none")),
                     ("fp32_prefix", dict(compute_dtype="float32", image_token_mode="prefix"))):
        policy, pre, post, (ik, sk) = make_lerobot_policy(model, dev, fuse_io_normalization=False,
                                                          image_input_scale=1.0, **kw)
        batch = {ik: img, sk: st, "task": tasks}
        with torch.no_grad():
            post(policy.select_action(pre(batch)))
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            post(policy.select_action(pre(batch)))
            torch.cuda.synchronize()
        out[name] = {"value": len(tasks) / (time.perf_counter() - t0), "unit": UNIT, "batch": len(tasks)}
        del policy
        torch.cuda.empty_cache()
    return out



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
def cpu_reference_run(model: str, samples: int, warmup: int, seed: int = 0, batch: int = 8):
    """Times the fp32 CPU oracle (the reference's arithmetic restated, all host cores) on `samples` observations of the
    batch-64 workload, in batches of 8 (a bounded sample: ~30-60 s of CPU work).  Prompts are tokenised outside the
    timed region on both arms (the GPU arm caches token ids per prompt)."""
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
    batch = max(1, min(batch, samples))
    steps = max(1, samples // batch)
    images, states, tasks = make_batch(batch * (steps + warmup), seed=7)
    tok = SimpleByteTokenizer(arch.text.vocab)
    times = []
    for i in range(warmup + steps):
        sl = slice(i * batch, (i + 1) * batch)
        enc = tok([t + "\n" for t in tasks[sl]], padding="longest", truncation=True, max_length=64)
        ids = torch.cat([torch.full((batch, 1), IMAGE_TOKEN_INDEX, dtype=torch.long), enc["input_ids"]], 1)
        mask = torch.cat([torch.ones(batch, 1, dtype=torch.long), enc["attention_mask"]], 1)
        t0 = time.perf_counter()
        out = oracle.forward(images[sl], states[sl], ids, mask)
        dt = time.perf_counter() - t0
        assert torch.isfinite(out).all()
        if i >= warmup:
            times.append(dt)
    per = sum(times) / (len(times) * batch)
    return {"value": 1.0 / per, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{len(times) * batch} observations of the same workload (480x480 frame + prompt + state) as "
                      f"{len(times)} fp32 forwards of batch {batch} through the CPU oracle port, "
                      f"{torch.get_num_threads()} torch threads, {per:.2f} s/sample; fp32 CPU vs the bf16 GPU arm: a "
                      "stated baseline, not like for like"}, per * batch


# ------------------------------------------------------------------------------------------------
def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])  # reference: a step = 8 CPU samples
    ap.add_argument("--model", default="fastvlm-0.5b")
    ap.add_argument("--batch", type=int, default=64, help="per-GPU batch (weak scaling) / global batch (strong scaling)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: --batch observations per GPU (the driver's contract); strong: ONE batch of --batch "
                         "observations sharded over the ranks (BASELINE configs[1]: 'batch-sharded at 2/4/8')")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-default-config", action="store_true", help="skip the fp32 default-configuration figures")
    ap.add_argument("--cpu-samples", type=int, default=64, help="samples timed for cpu_baseline (batches of 8, ~0.5 s/sample)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    global_batch = args.batch * world
    if args.scaling == "strong":   # one global batch, contiguous shards; per-rank sizes differ by at most one
        global_batch = args.batch
        lo, hi = shard_range(args.batch, rank, world)
        if hi - lo < 1:
            raise SystemExit(f"--scaling strong: batch {args.batch} leaves rank {rank} of {world} without work")
        args.batch = hi - lo
    config = {"workload": f"FastVLA-{args.model.split('-')[-1]} bf16 batched select_action, batch {args.batch}/GPU synthetic "
                          "MetaWorld-MT50-shaped obs (480x480 frame -> 1024^2, 4-dim state, 8-16 token prompt, image tokens "
                          "prefixed: T'=256+T_text)",
              "per_gpu_batch": args.batch, "global_batch": global_batch, "image": list(IMG_HW),
              "model": args.model,
              "parallelism": f"dp{world} (" + ("one global batch sharded over the ranks" if args.scaling == "strong"
                                                else "batch sharded") + ", no collective)",
              "l2_policy": "inputs+weights per step (177 MB images + 1.25 GB weights) exceed the 126 MB L2"}

    # ------------------------------------------------------------ reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return
        base, per = cpu_reference_run(args.model, 8 * max(1, args.steps), max(0, args.warmup))
        line = {"metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": per * 1e3, "higher_is_better": True, "scaling": args.scaling,
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

    policy, pre, post, feat_keys = make_lerobot_policy(args.model, dev)
    engine = policy.model.backbone.model.engine  # builds + uploads weights
    img_key, state_key = feat_keys
    frames_h, states_h, tasks = make_camera_batch(args.batch, seed=100 + rank)  # uint8 HWC frames + raw state, pinned host
    batch_host = {img_key: frames_h, state_key: states_h, "task": tasks}
    frames_d, states_d = frames_h.to(dev), states_h.to(dev)
    prompts = [t + "\n" for t in tasks]
    ids, lens, pool_idx = policy.model.backbone._prompt_ids(prompts)
    policy.select_action(pre(batch_host))  # pushes the dataset statistics into the head kernel, builds the stager
    policy.reset()

    def step_device():
        return engine.forward(frames_d, ids, lens, states=states_d, pool_idx=pool_idx, nhwc=True,
                              img_scale=1.0 / 255.0)

    def step_e2e():
        """The call a LeRobot rollout makes: pre-processor -> FastVLAPolicy.select_action -> post-processor, observation
        in pinned HOST memory (uint8 HWC camera frames, raw state), action back on the host (blocking)."""
        return post(policy.select_action(pre(batch_host)))

    out_pin = [torch.empty((args.batch, ACTION_DIM), dtype=torch.float32).pin_memory() for _ in range(2)]
    out_ready = [torch.cuda.Event() for _ in range(2)]

    def run_pipelined(n):
        """Same call, but the actions of step i are read back (pinned, async) one step late, so the staged H2D copy of
        step i+1 (side stream inside the policy) and the read-back overlap the forward of step i — a serving loop."""
        cur = torch.cuda.current_stream(dev)
        res = None
        for i in range(n):
            slot = i & 1
            act = policy.select_action(pre(batch_host))
            out_pin[slot].copy_(act, non_blocking=True)
            out_ready[slot].record(cur)
            if i > 0:
                out_ready[slot ^ 1].synchronize()
                res = out_pin[slot ^ 1].clone()
        out_ready[(n - 1) & 1].synchronize()
        return out_pin[(n - 1) & 1].clone() if res is not None or n == 1 else res

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

        # ---- timed region 2: end to end through the LeRobot plugin (host observation in, host action out) ----
        for _ in range(2):
            res = step_e2e()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            res = step_e2e()
        barrier()
        e2e_s = max_over_ranks(time.perf_counter() - t0, dev)
        assert res.device.type == "cpu" and torch.isfinite(res).all()
        h2d = policy._stager.h2d_bytes + ids.numel() * 4 + lens.numel() * 8
        d2h = res.numel() * 4
        run_pipelined(2)
        barrier()
        t0 = time.perf_counter()
        run_pipelined(args.steps)
        barrier()
        pipe_s = max_over_ranks(time.perf_counter() - t0, dev)

        # ---- per-kernel roofline pass (CUDA events around every launch, same stream) ----
        roof = None
        breakdown = None
        kernels = None
        if rank == 0:
            engine.set_profile(True)
            for _ in range(min(args.steps, 3)):
                step_device()
            rows = engine.profile_report()
            engine.set_profile(False)
            n_prof = min(args.steps, 3)
            tot = sum(r["total_ms"] for r in rows)

            def family(label):
                return ("gemm_tcgen05" if ("gemm" in label or "ffn_fused" in label) else
                        "dwconv" if "dwconv" in label else "attention" if "attention" in label else "other")

            def kernel_of(label):
                return ("ffn_fused_kernel" if "ffn_fused" in label else
                        "gemm_bf16_tcgen05_kernel" if "gemm" in label else
                        "dwconv7_mma_kernel" if ("dwconv k7 s1" in label and label.endswith("HW16")) else
                        "dwconv7_mma_r4_kernel" if ("dwconv k7 s1" in label or "dwconv k7 s2" in label) else   # both strides
                        "dwconv3_tma_kernel" if "dwconv k3 s1 m1" in label else
                        "attention_vis" if "vis.attention" in label else
                        "attention_llm" if "llm.attention" in label else
                        "stem_fused_kernel" if "stem_fused" in label else None)

            def acc(table, key, r):
                f = table.setdefault(key, dict(ms=0.0, flops=0.0, bytes=0.0, launches=0))
                f["ms"] += r["total_ms"]; f["flops"] += r["flops"]; f["bytes"] += r["bytes"]; f["launches"] += r["count"]

            fam, ker = {}, {}
            for r in rows:
                acc(fam, family(r["label"]), r)
                if kernel_of(r["label"]) is not None:
                    acc(ker, kernel_of(r["label"]), r)
            peaks = {}
            pk = ROOT / "MEASURED_PEAKS.json"
            if pk.is_file():
                peaks = json.loads(pk.read_text())
            peak_tf = float(peaks.get("bf16_tflops_sustained", 1400.0))
            hbm = float(peaks.get("hbm_gbs", 6650.0))
            peak_src = ("MEASURED_PEAKS.json (bf16_tflops_sustained, hbm_gbs)" if peaks
                        else "fallback 1.4 PFLOP/s sustained, 6.65 TB/s")
            # DRAM bytes per launch from the committed ncu pass of this command line (profiles/); null until it exists
            traffic = {}
            for tj in (ROOT / "profiles" / "r02_ncu_traffic.json", ROOT / "profiles" / "r01_ncu_traffic.json"):
                if tj.is_file():
                    traffic = json.loads(tj.read_text()).get("kernels", {})
                    if traffic:
                        break

            def roof_of(name, v):
                tensor = name in ("gemm_bf16_tcgen05_kernel", "ffn_fused_kernel", "attention_vis", "attention_llm")
                achieved = (v["flops"] / v["ms"] / 1e9) if tensor else (v["bytes"] / v["ms"] / 1e6)
                peak = peak_tf if tensor else hbm
                tr = traffic.get(name)
                return {"bound": "tensor" if tensor else "hbm", "kernel": name, "achieved": achieved, "peak": peak,
                        "unit": "TFLOP/s" if tensor else "GB/s", "frac": achieved / peak,
                        "traffic": (tr["dram_bytes"] / tr["launches"]) if tr and tr.get("launches") else None,
                        "algorithmic_bytes_per_launch": v["bytes"] / max(1, v["launches"]),
                        "launches_per_step": v["launches"] // n_prof,
                        "avg_launch_ms": v["ms"] / max(1, v["launches"]), "share_of_step": v["ms"] / tot,
                        "peak_source": peak_src}

            kernels = {k: roof_of(k, v) for k, v in sorted(ker.items(), key=lambda kv: -kv[1]["ms"])}
            # the dominant kernel of the step (largest share) is the headline roofline entry
            roof = next(iter(kernels.values())) if kernels else None
            breakdown = {k: {"ms_per_step": v["ms"] / n_prof, "share": v["ms"] / tot,
                             "tflops": v["flops"] / v["ms"] / 1e9 if v["ms"] else 0.0,
                             "gbs": v["bytes"] / v["ms"] / 1e6 if v["ms"] else 0.0,
                             "hbm_frac": (v["bytes"] / v["ms"] / 1e6) / hbm if v["ms"] else 0.0}
                         for k, v in fam.items()}

        # ---- p50 latency of a single observation through the same plugin call (the metric's second half) ----
        lat = None
        if rank == 0:
            one = {img_key: frames_h[:1].clone().pin_memory(), state_key: states_h[:1].clone().pin_memory(),
                   "task": tasks[:1]}
            ts = []
            for i in range(40):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                post(policy.select_action(pre(one)))
                if i >= 8:
                    ts.append((time.perf_counter() - t0) * 1e3)
            lat = {"p50_ms": statistics.median(ts), "p90_ms": sorted(ts)[int(0.9 * len(ts))], "batch": 1,
                   "path": "pre-processor -> FastVLAPolicy.select_action (LeRobot plugin) -> post-processor; uint8 HWC "
                           "frame + state in pinned host memory, action on the host"}

        # ---- the plugin's DEFAULT configuration (fp32 parity mode, reference-literal image_token_mode='none') ----
        default_cfg = None
        if rank == 0 and not args.no_default_config:
            default_cfg = default_config_throughput(args.model, dev, frames_h[:16], states_h[:16], tasks[:16])

    total = global_batch * args.steps
    value = total / (ms / 1e3)
    cpu_base = None
    if rank == 0 and not args.no_cpu_baseline and world == 1:
        cpu_base, _ = cpu_reference_run(args.model, args.cpu_samples, 1)
    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(3, args.warmup), "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": args.scaling, "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": config,
                "clocks": clocks,
                "e2e": {"value": total / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "path": "pre-processor -> lerobot FastVLAPolicy.select_action(batch) -> post-processor, "
                                "synchronous (action on the host before the next observation is submitted)",
                        "pipelined_value": total / pipe_s},
                "gpu_launches": int(launches_per_step * args.steps),
                "algorithmic_tflop_per_step": flops_per_step / 1e12,
                "achieved_tflops_whole_step": flops_per_step * world / (ms / args.steps) / 1e9,
                "roofline_frac_whole_step": flops_per_step / (ms / args.steps) / 1e9 /
                                            float(json.loads((ROOT / "MEASURED_PEAKS.json").read_text()).get(
                                                "bf16_tflops_sustained", 1400.0)
                                                  if (ROOT / "MEASURED_PEAKS.json").is_file() else 1400.0),
                "roofline": roof, "roofline_kernels": kernels, "kernel_families": breakdown, "latency": lat,
                "default_config": default_cfg, "cpu_baseline": cpu_base, "impl": "b200"}
        print(json.dumps(line), flush=True)
    if world > 1:
        import torch.distributed as dist

        dist.destroy_process_group()


if __name__ == "__main__":
    main()
