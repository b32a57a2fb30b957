# one GPU box call: the GPU test suite, one bench line, the per-launch tables (outputs under gpurun_out/)
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/t_gpu.log; cat gpurun_out/t_gpu.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; tail -c 300 gpurun_out/r02_bench_n1.json
python scripts/profile_forward.py --batch 64 --steps 3 --warmup 2 > gpurun_out/r02_step_final_b64.txt 2>&1; head -4 gpurun_out/r02_step_final_b64.txt
python scripts/profile_forward.py --batch 1 --steps 5 --warmup 3 > gpurun_out/r02_b1_profile.txt 2>&1; head -4 gpurun_out/r02_b1_profile.txt
