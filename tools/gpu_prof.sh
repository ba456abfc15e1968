mkdir -p gpurun_out
python tools/profile_run.py 2048 > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_l.csv python tools/profile_run.py 2048 > gpurun_out/r2_ncu_a.log 2>&1
python tools/summarize_launches.py gpurun_out/r2_l.csv > gpurun_out/r02_launches_2048.txt 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed --clock-control none --csv --log-file gpurun_out/r2_t.csv python tools/profile_run.py 2048 > gpurun_out/r2_ncu_b.log 2>&1
python tools/traffic_summary.py gpurun_out/r2_t.csv gpurun_out/r02_traffic_2048.json "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed --clock-control none python tools/profile_run.py 2048" > gpurun_out/r02_traffic_2048.txt 2>&1
ncu --set full --clock-control none --import-source on -f -k regex:k_gemm -s 300 -c 3 -o gpurun_out/r02_gemm python tools/profile_run.py 2048 > gpurun_out/r2_ncu_c.log 2>&1
ncu --set full --clock-control none --import-source on -f -k regex:k_sv_tri_fwd -s 40 -c 2 -o gpurun_out/r02_svtri python tools/profile_run.py 2048 > gpurun_out/r2_ncu_d.log 2>&1
ncu --set full --clock-control none --import-source on -f -k regex:k_front_rows -c 1 -o gpurun_out/r02_rows python tools/profile_run.py 2048 > gpurun_out/r2_ncu_e.log 2>&1
ncu --set full --clock-control none --import-source on -f -k regex:k_trtri_diag -c 1 -o gpurun_out/r02_trtri python tools/profile_run.py 2048 > gpurun_out/r2_ncu_f.log 2>&1
ls -la gpurun_out/*.ncu-rep
