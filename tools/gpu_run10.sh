mkdir -p gpurun_out
timeout 1800 python -m pytest tests/test_gpu_parity.py tests/test_gpu_compress.py tests/test_gpu_hss.py -q -m gpu -x > gpurun_out/r2_t_all.log 2>&1; echo "all rc=$?" >> gpurun_out/r2_t_all.log
timeout 900 python bench.py --steps 5 --warmup 3 --no-c64 --no-cpu-baseline --no-compressed > gpurun_out/r2_bench7.json 2> gpurun_out/r2_bench7.err; echo "bench rc=$?" >> gpurun_out/r2_bench7.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_l.csv python tools/profile_run.py 2048 > gpurun_out/r2_ncu_a.log 2>&1
python tools/summarize_launches.py gpurun_out/r2_l.csv > gpurun_out/r02_v2_launches_2048.txt 2>&1
rm -f gpurun_out/r2_l.csv
