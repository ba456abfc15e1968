mkdir -p gpurun_out
timeout 1800 python -m pytest tests/test_gpu_hss.py tests/test_gpu_compress.py tests/test_gpu_parity.py tests/test_ordering.py -q -m gpu -x > gpurun_out/r2_t_all.log 2>&1; echo "all rc=$?" >> gpurun_out/r2_t_all.log
HS_PLAN_TIMING=1 timeout 600 python tools/hss_run.py 2048 poisson 128 1e-5 32 > gpurun_out/r2_hss_timing2.log 2>&1
timeout 900 python bench.py --steps 5 --warmup 3 --no-c64 --no-cpu-baseline > gpurun_out/r2_bench9.json 2> gpurun_out/r2_bench9.err; echo "bench rc=$?" >> gpurun_out/r2_bench9.err
timeout 900 python tools/hss_run.py 48 helmholtz3d 256 1e-3 32 > gpurun_out/r2_hss3d_48.log 2>&1; echo "rc=$?" >> gpurun_out/r2_hss3d_48.log
timeout 900 python tools/hss_run.py 48 helmholtz3d 256 1e-3 32 0 > gpurun_out/r2_lr3d_48.log 2>&1; echo "rc=$?" >> gpurun_out/r2_lr3d_48.log
