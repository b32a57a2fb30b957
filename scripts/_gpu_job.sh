set -x
timeout 400 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err
tail -c 300 gpurun_out/r02_bench_n1.err
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-default-config > gpurun_out/plain_b.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 2600 -c 1100 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-default-config > gpurun_out/ncu_b.log 2>&1
tail -2 gpurun_out/ncu_b.log
python __graft_entry__.py smoke 2>&1 | tail -3
