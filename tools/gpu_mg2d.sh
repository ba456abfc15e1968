mkdir -p gpurun_out
N=${1:-8}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2_bench_n${N}_2048.json 2> gpurun_out/r2_bench_n${N}_2048.err; echo "rc=$?" >> gpurun_out/r2_bench_n${N}_2048.err
