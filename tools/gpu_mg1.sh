mkdir -p gpurun_out
N=${1:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r2_bench_n${N}_2048.json 2> gpurun_out/r2_bench_n${N}_2048.err; echo "rc=$?" >> gpurun_out/r2_bench_n${N}_2048.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --workload 3d --grid3 64 --steps 3 --warmup 3 > gpurun_out/r2_bench3d_n${N}_64.json 2> gpurun_out/r2_bench3d_n${N}_64.err; echo "rc=$?" >> gpurun_out/r2_bench3d_n${N}_64.err
timeout 600 python -m pytest tests/test_parallel.py -q -m gpu > gpurun_out/r2_t_par_n${N}.log 2>&1; echo "rc=$?" >> gpurun_out/r2_t_par_n${N}.log
