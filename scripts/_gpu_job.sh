set -x
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err
tail -c 300 gpurun_out/r02_bench_n1.err
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-default-config > gpurun_out/plain_b.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 2600 -c 1100 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-default-config > gpurun_out/ncu_b.log 2>&1
tail -2 gpurun_out/ncu_b.log
python scripts/one_gemm.py 131072 1536 384 gelu16 > gpurun_out/plain_g1.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_tcgen05 -s 4 -c 1 -f -o gpurun_out/r02_gemm_fc1_s2 python scripts/one_gemm.py 131072 1536 384 gelu16 > gpurun_out/ncu_g1.log 2>&1
python scripts/one_gemm.py 131072 384 1536 res > gpurun_out/plain_g2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_tcgen05 -s 4 -c 1 -f -o gpurun_out/r02_gemm_fc2_s2 python scripts/one_gemm.py 131072 384 1536 res > gpurun_out/ncu_g2.log 2>&1
cat gpurun_out/plain_g1.log gpurun_out/plain_g2.log
