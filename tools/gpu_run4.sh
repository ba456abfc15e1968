mkdir -p gpurun_out
timeout 2400 python -m pytest tests -q -m gpu -x > gpurun_out/r2_t_all.log 2>&1; echo "all rc=$?" >> gpurun_out/r2_t_all.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench2.json 2> gpurun_out/r2_bench2.err; echo "bench rc=$?" >> gpurun_out/r2_bench2.err
HS_PROFILE=1 timeout 600 python tools/hss_run.py 2048 poisson 128 1e-5 32 > gpurun_out/r2_hssprof_2048.log 2>&1
