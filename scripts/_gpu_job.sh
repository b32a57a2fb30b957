set -x
python scripts/one_dwconv.py 32 128 192 3 1 1 > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:dwconv3_tma -s 1 -c 1 -f -o gpurun_out/r02_dw3 python scripts/one_dwconv.py 32 128 192 3 1 1 > gpurun_out/ncu_dw3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:stem_fused -s 2 -c 1 -f -o gpurun_out/r02_stem python scripts/profile_forward.py --batch 32 --steps 1 --warmup 1 > gpurun_out/ncu_stem.log 2>&1
python scripts/one_dwconv.py 32 256 96 7 2 2 > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:dwconv7_mma_r4 -s 1 -c 1 -f -o gpurun_out/r02_dw7_s2 python scripts/one_dwconv.py 32 256 96 7 2 2 > gpurun_out/ncu_dw7s2.log 2>&1
python bench.py --model fastvlm-7b --batch 32 --steps 10 --warmup 3 --no-cpu-baseline --no-default-config > gpurun_out/r02_bench_7b_b32_n1.json 2> gpurun_out/r02_bench_7b.err
python bench.py --model fastvlm-1.5b --batch 32 --steps 10 --warmup 3 --no-cpu-baseline --no-default-config > gpurun_out/r02_bench_1p5b_b32_n1.json 2> gpurun_out/r02_bench_1p5b.err
python scripts/bench_sweep.py > gpurun_out/r02_sweep_aloha_bf16.jsonl 2> gpurun_out/r02_sweep.err
tail -3 gpurun_out/r02_sweep_aloha_bf16.jsonl | cut -c1-300
