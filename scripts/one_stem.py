"""Time the vision stem inside a forward (profile rows only): python scripts/one_stem.py"""
import subprocess, sys
out = subprocess.run([sys.executable, "scripts/profile_forward.py", "--batch", "64", "--steps", "4"], capture_output=True, text=True).stdout
for l in out.splitlines():
    if l.startswith("# forward") or "stem" in l or "preprocess" in l:
        print(l)
