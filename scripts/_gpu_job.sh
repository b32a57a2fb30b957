set -x
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench3.json 2> gpurun_out/r02_bench3.err
tail -c 300 gpurun_out/r02_bench3.err
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-default-config > gpurun_out/plain_b.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 2600 -c 1100 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-default-config > gpurun_out/ncu_b.log 2>&1
tail -2 gpurun_out/ncu_b.log
python scripts/time_attn.py > gpurun_out/plain_attn.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:attn_tc -s 4 -c 2 -o gpurun_out/r02_attn_tc python scripts/time_attn.py > gpurun_out/ncu_attn.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:attn_tc_causal -s 2 -c 1 -o gpurun_out/r02_attn_tc_causal python scripts/time_attn.py >> gpurun_out/ncu_attn.log 2>&1
tail -3 gpurun_out/ncu_attn.log
ls -la gpurun_out/*.ncu-rep | tail -3
