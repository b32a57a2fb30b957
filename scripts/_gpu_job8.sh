set -x
N=$1
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 --no-default-config > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err
grep '^{' gpurun_out/r02_bench_n$N.json | tail -c 300
