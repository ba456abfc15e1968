# 3D complex Helmholtz on N GPUs of one box:  bash tools/gpu_mg3d.sh N "grid sizes"
mkdir -p gpurun_out
N=${1:-8}
for G in ${2:-96}; do
  EXTRA=""
  if [ "$G" -ge 112 ]; then EXTRA="--no-e2e"; fi
  timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 295$((RANDOM % 90 + 10)) bench.py --gpus $N --workload 3d --grid3 $G --steps 2 --warmup 3 $EXTRA > gpurun_out/r2_bench3d_n${N}_${G}.json 2> gpurun_out/r2_bench3d_n${N}_${G}.err; echo "rc=$?" >> gpurun_out/r2_bench3d_n${N}_${G}.err
  nvidia-smi --query-gpu=index,memory.used,memory.total --format=csv,noheader > gpurun_out/r2_mem_n${N}_${G}.txt 2>&1
done
