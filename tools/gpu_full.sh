# full single-GPU validation: every GPU test, the default bench line (c64 block and CPU baseline included), the 3D line
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -q -m gpu -x > gpurun_out/r2_t_full.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2_t_full.log
timeout 1500 python bench.py > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err; echo "bench rc=$?" >> gpurun_out/r2_bench_final.err
timeout 1500 python bench.py --workload 3d --grid3 96 --no-cpu-baseline --steps 2 --warmup 3 > gpurun_out/r2_bench3d_96_final.json 2> gpurun_out/r2_bench3d_96_final.err; echo "bench3d rc=$?" >> gpurun_out/r2_bench3d_96_final.err
