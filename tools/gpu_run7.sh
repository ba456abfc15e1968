mkdir -p gpurun_out
HS_PLAN_TIMING=1 timeout 600 python tools/e2e_timing.py 2048 > gpurun_out/r2_e2e_timing.log 2>&1
HS_PLAN_TIMING=1 timeout 600 python tools/hss_run.py 2048 poisson 128 1e-5 32 > gpurun_out/r2_hss_timing.log 2>&1
timeout 1800 python -m pytest tests/test_gpu_hss.py tests/test_gpu_compress.py tests/test_gpu_parity.py -q -m gpu > gpurun_out/r2_t_new.log 2>&1; echo "rc=$?" >> gpurun_out/r2_t_new.log
