mkdir -p gpurun_out
timeout 2400 python -m pytest tests -q -m gpu -x > gpurun_out/r2_t_all.log 2>&1; echo "all rc=$?" >> gpurun_out/r2_t_all.log
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-c64 > gpurun_out/r2_bench4.json 2> gpurun_out/r2_bench4.err; echo "bench rc=$?" >> gpurun_out/r2_bench4.err
timeout 600 python tools/hss_run.py 2048 poisson 128 1e-5 32 > gpurun_out/r2_hssrun_2048b.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_hss2048_launches.csv python tools/hss_run.py 2048 poisson 128 1e-5 32 > gpurun_out/r2_ncu_hss.log 2>&1
python tools/summarize_launches.py gpurun_out/r2_hss2048_launches.csv > gpurun_out/r2_hss2048_launches.txt 2>&1
rm -f gpurun_out/r2_hss2048_launches.csv gpurun_out/r2_hss1024_launches.csv
timeout 900 python bench.py --workload 3d --grid3 64 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench3d_64.json 2> gpurun_out/r2_bench3d_64.err; echo "rc=$?" >> gpurun_out/r2_bench3d_64.err
timeout 1200 python bench.py --workload 3d --grid3 96 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2_bench3d_96.json 2> gpurun_out/r2_bench3d_96.err; echo "rc=$?" >> gpurun_out/r2_bench3d_96.err
