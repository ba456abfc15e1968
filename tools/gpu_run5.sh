mkdir -p gpurun_out
timeout 2400 python -m pytest tests -q -m gpu -x > gpurun_out/r2_t_all.log 2>&1; echo "all rc=$?" >> gpurun_out/r2_t_all.log
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-c64 > gpurun_out/r2_bench3.json 2> gpurun_out/r2_bench3.err; echo "bench rc=$?" >> gpurun_out/r2_bench3.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_hss1024_launches.csv python tools/hss_run.py 1024 poisson 128 1e-5 32 > gpurun_out/r2_ncu_hss.log 2>&1
python tools/summarize_launches.py gpurun_out/r2_hss1024_launches.csv > gpurun_out/r2_hss1024_launches.txt 2>&1
