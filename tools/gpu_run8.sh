mkdir -p gpurun_out
timeout 1800 python -m pytest tests -q -m gpu -x > gpurun_out/r2_t_all.log 2>&1; echo "all rc=$?" >> gpurun_out/r2_t_all.log
HS_PLAN_TIMING=1 timeout 600 python tools/hss_run.py 2048 poisson 128 1e-5 32 > gpurun_out/r2_hss_timing.log 2>&1
timeout 900 python bench.py --steps 5 --warmup 3 --no-c64 > gpurun_out/r2_bench5.json 2> gpurun_out/r2_bench5.err; echo "bench rc=$?" >> gpurun_out/r2_bench5.err
