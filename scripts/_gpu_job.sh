mkdir -p gpurun_out
for b in 1 64; do
  python scripts/profile_forward.py --batch $b --steps 8 --warmup 3 2>&1 | head -3 | tail -1
done
timeout 600 python -m pytest tests -m gpu -x -q -k "engine or head or plugin or golden or bench_parity" 2>&1 | tail -3
