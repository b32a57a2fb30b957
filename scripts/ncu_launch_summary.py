"""Summarise an ncu launch list (--csv, metrics gpu__time_duration.sum [+ dram__bytes_read.sum, dram__bytes_write.sum])
per kernel: launches, total device time, share, DRAM bytes.  Writes a text table and (optionally) a JSON with the
traffic of the tcgen05 GEMM family that bench.py reports as roofline.traffic.

  python scripts/ncu_launch_summary.py gpurun_out/launches.csv profiles/r01_ncu_launch_summary.txt [profiles/r01_ncu_traffic.json]
"""
import csv
import json
import re
import sys
from collections import defaultdict


def short(name: str) -> str:
    m = re.search(r"(\w+)(<[^(]*>)?\(", name)
    base = m.group(1) if m else name
    t = re.search(r"<([^>]*)>", name.split("(")[0].split("::")[-1] if "(" in name else name)
    targs = re.findall(r"\(int\)(\d+)|\(bool\)(\d)", name.split(">(")[0]) if "<" in name else []
    flat = ",".join(a or b for a, b in targs)
    return f"{base}<{flat}>" if flat else base


def main() -> None:
    src, out_txt = sys.argv[1], sys.argv[2]
    out_json = sys.argv[3] if len(sys.argv) > 3 else None
    rows = [r for r in csv.reader(open(src, errors="replace")) if len(r) > 10]
    hdr = rows[0]
    ci = {h: i for i, h in enumerate(hdr)}
    per = defaultdict(lambda: dict(n=0, ns=0.0, rd=0.0, wr=0.0))
    ids = set()
    for r in rows[1:]:
        if r[ci["ID"]] == "ID":
            continue
        name, metric, unit, val = r[ci["Kernel Name"]], r[ci["Metric Name"]], r[ci["Metric Unit"]], r[ci["Metric Value"]]
        v = float(val.replace(",", ""))
        k = per[short(name)]
        scale = {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
        if metric == "gpu__time_duration.sum":
            k["ns"] += v * scale
            if (name, r[ci["ID"]]) not in ids:
                ids.add((name, r[ci["ID"]]))
                k["n"] += 1
        elif metric == "dram__bytes_read.sum":
            k["rd"] += v * scale
        elif metric == "dram__bytes_write.sum":
            k["wr"] += v * scale
    tot = sum(k["ns"] for k in per.values())
    lines = [f"# {src}: {sum(k['n'] for k in per.values())} launches, {tot / 1e6:.2f} ms device time (cold-cache, serialised: compare shares)",
             f"{'kernel':60s} {'n':>6s} {'ms':>9s} {'share':>7s} {'dram rd MB':>11s} {'dram wr MB':>11s}"]
    for name, k in sorted(per.items(), key=lambda kv: -kv[1]["ns"]):
        lines.append(f"{name[:60]:60s} {k['n']:6d} {k['ns'] / 1e6:9.3f} {100 * k['ns'] / tot:6.2f}% {k['rd'] / 1e6:11.1f} {k['wr'] / 1e6:11.1f}")
    open(out_txt, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines[:16]))
    if out_json:
        fam = dict(launches=0, dram_bytes=0.0, ms=0.0)
        for name, k in per.items():
            if "tcgen05" in name or "ffn_fused" in name:
                fam["launches"] += k["n"]; fam["dram_bytes"] += k["rd"] + k["wr"]; fam["ms"] += k["ns"] / 1e6
        # per kernel (template instances merged): what bench.py reports as roofline_kernels[*].traffic
        kernels = {}
        alias = {"gemm_bf16_tcgen05_kernel": "gemm_bf16_tcgen05_kernel", "ffn_fused": "ffn_fused_kernel",
                 "dwconv7_mma_r4_kernel": "dwconv7_mma_r4_kernel", "dwconv7_mma_kernel": "dwconv7_mma_kernel",
                 "dwconv3_tma_kernel": "dwconv3_tma_kernel",
                 "dwconv7_s2m2": "dwconv7_s2m2_kernel", "stem_fused_kernel": "stem_fused_kernel",
                 "attn_tc_causal_kernel": "attention_llm", "attn_tc_kernel": "attention_vis",
                 "flash_attn_v2_kernel": "attention_vis", "flash_attn_gqa_kernel": "attention_llm"}
        for name, k in per.items():
            for pat, key in alias.items():
                if pat in name:
                    d = kernels.setdefault(key, dict(launches=0, dram_bytes=0.0, ms=0.0))
                    d["launches"] += k["n"]; d["dram_bytes"] += k["rd"] + k["wr"]; d["ms"] += k["ns"] / 1e6
                    break
        json.dump({"source": src, "tcgen05_gemm_family": fam, "kernels": kernels, "total_ms": tot / 1e6},
                  open(out_json, "w"), indent=1)


if __name__ == "__main__":
    main()
