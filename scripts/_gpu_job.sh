set -x
mkdir -p gpurun_out
for b in 1 64; do
  python scripts/profile_forward.py --batch $b --steps 5 --warmup 3 2>&1 | head -4 > gpurun_out/pdl_on_b$b.txt
  FVLA_DISABLE_PDL=1 python scripts/profile_forward.py --batch $b --steps 5 --warmup 3 2>&1 | head -4 > gpurun_out/pdl_off_b$b.txt
done
head -4 gpurun_out/pdl_*.txt
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/t_gpu.log
cat gpurun_out/t_gpu.log
