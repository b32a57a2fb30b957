"""Per-instruction stall breakdown of an ncu --set full report: python scripts/ncu_stalls.py <rep> [n_lines] [lo hi]
(prints the top-n sampled SASS lines with their dominant stall reasons, L1 shared wavefronts and conflicts)"""
import csv, subprocess, sys
rep = sys.argv[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hdr = rows[1]
col = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
data = []
for idx, r in enumerate(rows[2:]):
    try:
        data.append((int(r[col["# Samples"]]), idx, r))
    except Exception:
        pass
tot = sum(d[0] for d in data)
print("total samples", tot)
if len(sys.argv) > 4:   # dump a range of instructions in program order
    lo, hi = int(sys.argv[3]), int(sys.argv[4])
    for v, idx, r in data[lo:hi]:
        st = sorted(((int(r[col[c]] or 0), c[6:]) for c in stall_cols), reverse=True)[:2]
        print(f"{idx:5d} {v:5d} ex={r[col['Instructions Executed']]:>8} wf={r[col['L1 Wavefronts Shared']]:>8} "
              f"{st[0][1]}:{st[0][0]} {st[1][1]}:{st[1][0]} | {r[col['Source']].strip()[:90]}")
    sys.exit(0)
for v, idx, r in sorted(data, reverse=True)[:n]:
    st = sorted(((int(r[col[c]] or 0), c[6:]) for c in stall_cols), reverse=True)[:3]
    print(f"{idx:5d} {v:5d} {100 * v / tot:4.1f}% ex={r[col['Instructions Executed']]:>8} wf={r[col['L1 Wavefronts Shared']]:>8} "
          f"xs={r[col['L1 Wavefronts Shared Excessive']]:>7} " + " ".join(f"{c}:{x}" for x, c in st) + f" | {r[col['Source']].strip()[:80]}")
