set -x
mkdir -p gpurun_out
for N in 4 2; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 10 --warmup 3 --scaling strong --no-default-config > gpurun_out/r02_bench_strong_n$N.json 2> gpurun_out/r02_bench_strong_n$N.err
python -c "
import json;d=json.load(open('gpurun_out/r02_bench_strong_n$N.json'));print(d['n_gpus'],d['scaling'],round(d['value'],1),round(d['ms_per_step'],2),d['config']['per_gpu_batch'],d['config']['global_batch'],round(d['e2e']['value'],1))" || tail -5 gpurun_out/r02_bench_strong_n$N.err
done
