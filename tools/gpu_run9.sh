mkdir -p gpurun_out
timeout 1800 python -m pytest tests -q -m gpu -x --deselect tests/test_parallel.py > gpurun_out/r2_t_all.log 2>&1; echo "all rc=$?" >> gpurun_out/r2_t_all.log
timeout 900 python bench.py --steps 5 --warmup 3 --no-c64 --no-cpu-baseline > gpurun_out/r2_bench6.json 2> gpurun_out/r2_bench6.err; echo "bench rc=$?" >> gpurun_out/r2_bench6.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_l.csv python tools/profile_run.py 2048 > gpurun_out/r2_ncu_a.log 2>&1
python tools/summarize_launches.py gpurun_out/r2_l.csv > gpurun_out/r02_v1_launches_2048.txt 2>&1
timeout 1200 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed --clock-control none --csv --log-file gpurun_out/r2_t.csv python tools/profile_run.py 2048 > gpurun_out/r2_ncu_b.log 2>&1
python tools/traffic_summary.py gpurun_out/r2_t.csv gpurun_out/r02_traffic_2048.json "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_tensor_cycles_active... --clock-control none python tools/profile_run.py 2048" > gpurun_out/r02_traffic_2048.txt 2>&1
rm -f gpurun_out/r2_l.csv gpurun_out/r2_t.csv
